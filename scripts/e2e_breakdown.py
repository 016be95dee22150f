"""Host-side timing of every API call of one end-to-end bench step (setup cost analysis)."""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402

wl = bench.build_workload_gpu()
for rep in range(8):
    t = [time.perf_counter()]
    a = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192, update_mode=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    a.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION); t.append(time.perf_counter())
    with contextlib.redirect_stdout(io.StringIO()):
        a.initFromGridMap()
    t.append(time.perf_counter())
    a.setEndpoints(wl["start"], wl["goal"]); a.begin(1.0); t.append(time.perf_counter())
    a.iterate(5); a.sync(); t.append(time.perf_counter())
    ids, dirs, L = a.bestPath(); t.append(time.perf_counter())
    c = a.counters()
    del a
    t.append(time.perf_counter())
    names = ["grid_from_occupancy", "acs_create", "endpoints+begin", "iterate(5)+sync", "best", "destroy"]
    print("rep %d: " % rep + "  ".join("%s %.2f ms" % (n, 1e3 * (t[i + 1] - t[i])) for i, n in enumerate(names)) +
          "  | total %.2f ms, %d ant-steps" % (1e3 * (t[-1] - t[0]), c["ant_steps"]), flush=True)
