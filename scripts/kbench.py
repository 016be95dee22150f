"""Update-path kernels timed alone on the C2 grid (fused TMA vs float4 evaporation vs D2D copy)."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402

wl = bench.build_workload()
for mode in (0, 1):
    acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192, update_mode=mode)
    acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
    with contextlib.redirect_stdout(io.StringIO()):
        acs.initFromGridMap()
    acs.setEndpoints(wl["start"], wl["goal"]); acs.begin(1.0); acs.iterate(3); acs.sync()
    acs.setTiming(True); acs.iterate(10); print('mode', mode, acs.kernelMs(), flush=True)
    nbytes = 256 ** 3 * 6 * 8
    for which in (0, 1, 2, 0, 1, 2):
        ms = acs.benchKernel(which, 20)
        print('which', which, 'ms %.4f' % ms, 'GB/s %.0f' % (nbytes / ms / 1e6), flush=True)
    del acs
