"""Per-instruction stall samples of one kernel from an ncu report (source page): total samples, the loop's share, and the
instructions that collect the most samples with their dominant stall reason.

    python scripts/ncu_source_hot.py gpurun_out/x.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
S, N, X = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(int(r[N]) for r in body)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples %d over %d instructions" % (tot, len(body)))
agg = {}
for r in body:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
mx = max(int(r[X]) for r in body)
print("\nidx  samples  share  executed  instruction  (dominant stall)")
for idx, r in enumerate(body):
    n = int(r[N])
    if n >= sorted((int(q[N]) for q in body), reverse=True)[min(top, len(body)) - 1] and n > 0:
        dom = max(stall, key=lambda i: int(r[i]))
        print("%4d %7d %5.1f%% %9s  %-60s %s" % (idx, n, 100.0 * n / tot, r[X], r[S].strip()[:60], h[dom][6:]))
