"""Turn ncu output into the summaries kept under profiles/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/x.csv "title" > profiles/x_summary.md
    python scripts/ncu_summary.py full gpurun_out/x.ncu-rep "title" > profiles/x_ncu.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores"]


def launches(path, title):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("wr::", "")
        t = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("%d launches, %.1f us of kernel time in total.  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the\n"
          "CUDA-event phase times of the un-profiled run, not absolutes.\n" % (len(rows), tot))
    print("| kernel | launches | avg us | sum us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, n, t / n, t, 100 * t / tot))


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    print("# %s\n" % title)
    for v in rows[2:]:
        d = dict(zip(h, v)); un = dict(zip(h, u))
        print("kernel: `%s`\n" % d.get("Kernel Name", "?"))
        print("| metric | value | unit |\n|---|---:|---|")
        for m in WANT:
            if m in d:
                print("| %s | %s | %s |" % (m, d[m], un[m]))
        for m in h:
            if "issue_stalled" in m and m.endswith("per_issue_active.ratio") and "not_issued" not in m:
                try:
                    if float(d[m]) >= 0.05:
                        print("| %s | %s | %s |" % (m, d[m], un[m]))
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
