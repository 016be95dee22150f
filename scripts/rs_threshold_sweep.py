"""Exploration: time of a whole C2 search (150 iterations) against the rank-set switch threshold WR_RANKSET_ON (deposit tiles of a
record-path iteration at or below which the search moves to rank sets), on 1 GPU or sharded (torchrun).  Also prints the
per-iteration trajectory (path taken, deposit tiles / row blocks, ms) of the default setting.

    python scripts/rs_threshold_sweep.py            |  python -m torch.distributed.run --nproc-per-node N scripts/rs_threshold_sweep.py
"""
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402
from welding_robot_b200 import _lib  # noqa: E402
from welding_robot_b200.dist import ShardedSearch  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.check(_lib.lib().wr_set_device(local))
torch.cuda.set_stream(torch.cuda.Stream())
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bench.select_workload("C2")
wl = bench.build_workload_gpu()
stream = torch.cuda.current_stream()


def make():
    a = wr.ACS_Rank(seed=bench.SEED, fixed_colony=bench.ANTS_PER_GPU * world, step_cap=bench.STEP_CAP, update_mode=4)
    a.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
    with contextlib.redirect_stdout(io.StringIO()):
        a.initFromGridMap()
    _lib.check(_lib.lib().wr_acs_set_stream(a._a, stream.cuda_stream))
    a.setEndpoints(wl["start"], wl["goal"])
    if world > 1:
        S = ShardedSearch(a, rank, world); S.begin(bench.PREDICT)
        return a, S.iterate
    a.begin(bench.PREDICT)
    return a, a.iterate


def sync():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


out = {"world": world, "sweep": {}, "trajectory": []}
out["traj_threshold"] = os.environ.get("WR_TRAJ_ON")
for thr in (["1000", "2000", "4000", "8000", "16000", "30000"] if len(sys.argv) < 2 else [t for t in sys.argv[1:] if t != "none"]):
    os.environ["WR_RANKSET_ON"] = thr
    for rep in range(2):
        a, step = make()
        sync()
        t0 = time.perf_counter()
        step(150)
        a.sync(); sync()
        dt = time.perf_counter() - t0
        st = a.updateStats()
        del a
    out["sweep"][thr] = {"seconds_150_iterations": dt, "rank_set_iterations": st["rankset_iterations"]}
if os.environ.get("WR_TRAJ_ON"):
    os.environ["WR_RANKSET_ON"] = os.environ["WR_TRAJ_ON"]
else:
    os.environ.pop("WR_RANKSET_ON", None)
a, step = make()
a.setTiming(True)
prev = a.kernelMs()
for it in range(80):
    sync()
    t0 = time.perf_counter()
    step(1)
    a.sync()
    dt = time.perf_counter() - t0
    st = a.updateStats()
    k = a.kernelMs()
    ph = [round(k[n] - prev[n], 3) for n in ("walk", "rank", "deposit_build", "update")]
    prev = k
    out["trajectory"].append([it, st["rankset_last"], st["deposit_tiles"], st["distinct_slots"], round(dt * 1e3, 3)] + ph)
if rank == 0:
    print(json.dumps(out))
if dist is not None:
    dist.destroy_process_group()
