"""Short C2 run for ncu captures: `iters` ACS iterations, then the fused update kernel alone a few times.

    ncu --set full --clock-control none --import-source on -k regex:k_update_fused -s 40 -c 2 -o gpurun_out/upd python scripts/profile_run.py 40
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K = int(sys.argv[3]) if len(sys.argv) > 3 else 6
wl = bench.build_workload_gpu()
acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192, update_mode=mode, K=K)
acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
with contextlib.redirect_stdout(io.StringIO()):
    acs.initFromGridMap()
acs.setEndpoints(wl["start"], wl["goal"])
acs.begin(1.0)
acs.iterate(iters)
acs.sync()
print("fused update alone: %.4f ms" % acs.benchKernel(0 if mode != 3 else 3, 3))
print(acs.counters())
