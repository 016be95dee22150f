"""Update-path kernels timed alone on the C2 grid: fused update vs all-TMA ring vs float4 evaporation vs
D2D copy, with no deposit records (fresh search), early (spread deposits) and converged (hot slots)."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402

wl = bench.build_workload_gpu()
acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192)
acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
with contextlib.redirect_stdout(io.StringIO()):
    acs.initFromGridMap()
acs.setEndpoints(wl["start"], wl["goal"]); acs.begin(1.0)
nbytes = 256 ** 3 * 6 * 8
names = {0: "update_fused", 1: "evaporate_float4", 2: "d2d_copy"}
for label, iters in (("no records", 0), ("after 3 iterations", 3), ("after 60 iterations", 57)):
    if iters:
        acs.iterate(iters)
    acs.sync()
    print("---", label, acs.counters()["deposit_records"], "records so far", flush=True)
    for which in (0, 3, 1, 2):
        ms = acs.benchKernel(which, 20)
        print("%-18s %.4f ms  %.0f GB/s" % (names[which], ms, nbytes / ms / 1e6), flush=True)
acs.setTiming(True); acs.iterate(10); print(acs.kernelMs())
