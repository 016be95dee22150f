"""Run-to-run spread of bench.py's timed region: the same 100 converged C2 iterations timed several times in one process,
with and without the NVML clock sampler thread, as 20 x iterate(5) and as one iterate(100).

    python scripts/timed_jitter.py [repeats]
"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402
from welding_robot_b200 import _lib  # noqa: E402

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
stream = torch.cuda.current_stream()
wl = bench.build_workload_gpu()
acs = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=8192, update_mode=4)
acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
with contextlib.redirect_stdout(io.StringIO()):
    acs.initFromGridMap()
_lib.check(_lib.lib().wr_acs_set_stream(acs._a, stream.cuda_stream))
acs.setEndpoints(wl["start"], wl["goal"])
acs.begin(1.0)
acs.iterate(25)
torch.cuda.synchronize()


def timed(calls, per_call):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = acs.counters()["ant_steps"]
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(calls):
        acs.iterate(per_call)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return ms / (calls * per_call), (acs.counters()["ant_steps"] - c0) / (ms * 1e-3)


for label, calls, per, sampler_on in (("20 x iterate(5), no sampler", 20, 5, False), ("20 x iterate(5), NVML sampler", 20, 5, True),
                                      ("1 x iterate(100), no sampler", 1, 100, False), ("100 x iterate(1), no sampler", 100, 1, False)):
    out = []
    for _ in range(rep):
        s = None
        if sampler_on:
            s = bench.ClockSampler(0); s.start(); s.mark_begin()
        out.append(timed(calls, per))
        if s:
            s.mark_end(); s.stop()
    print("%-32s ms/iteration %s   ant-steps/s %s" % (label, " ".join("%.4f" % a for a, _ in out), " ".join("%.3g" % b for _, b in out)))
