"""Per-phase kernel times (ms per iteration) over the life of one C2 search, for an update mode: shows where early
(wandering colony, ~1.5 M distinct deposit slots) and converged (~800 hot slots) iterations spend their time.

    python scripts/phase_timeline.py [update_mode] [K]
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wl = bench.build_workload_gpu()
acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192, update_mode=mode, K=K)
acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
with contextlib.redirect_stdout(io.StringIO()):
    acs.initFromGridMap()
acs.setEndpoints(wl["start"], wl["goal"]); acs.begin(1.0)
acs.setTiming(True)
done, prev, prec = 0, None, 0
for target in (1, 2, 3, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50, 60, 80, 120):
    acs.iterate(target - done)
    ms = acs.kernelMs(); c = acs.counters()
    d = {k: (ms[k] - (prev[k] if prev else 0.0)) / (target - done) for k in ms}
    print("mode %d K %d iter %3d..%3d: %s records/iter %d %s" % (mode, K, done + 1, target, {k: round(v, 4) for k, v in d.items()},
                                                               (c["deposit_records"] - prec) // (target - done), (acs.updateStats(), acs.fieldStats())), flush=True)
    prev, prec, done = ms, c["deposit_records"], target
