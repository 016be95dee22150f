"""Distribution of ant walk lengths per iteration on the C2 workload (tail analysis)."""
import contextlib
import ctypes as C
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402
from welding_robot_b200 import _lib  # noqa: E402

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = bench.build_workload_gpu()
acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=cap)
acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
with contextlib.redirect_stdout(io.StringIO()):
    acs.initFromGridMap()
acs.setEndpoints(wl["start"], wl["goal"]); acs.begin(1.0)
acs.setTiming(True)
done = 0
prev = 0.0
for target in (1, 2, 5, 10, 20, 40, 80):
    acs.iterate(target - done)
    steps = []
    for k in range(4096):
        n = C.c_int(); L = C.c_float(); o = C.c_int()
        _lib.check(_lib.lib().wr_acs_last_ant(acs._a, k, None, None, 0, C.byref(n), C.byref(L), C.byref(o)))
        steps.append(n.value - 1 if n.value > 0 else -1)
    s = np.array(steps); a = s[s >= 0]
    ms = acs.kernelMs()
    print("iter %3d: arrived %4d  mean %6.0f  p50 %5.0f p90 %5.0f p99 %5.0f max %5d | walk ms/iter over the last %d: %.3f" % (
        target, len(a), a.mean(), np.percentile(a, 50), np.percentile(a, 90), np.percentile(a, 99), a.max(), target - done,
        (ms["walk"] - prev) / (target - done)), flush=True)
    prev = ms["walk"]; done = target
print(acs.counters())
