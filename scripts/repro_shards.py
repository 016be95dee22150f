"""Debug helper: C2-sized colony as LocalShards (one process, one GPU) with the rank-set path forced at a chosen iteration,
to run under compute-sanitizer.  usage: python scripts/repro_shards.py <world> <ants_total> <rankset_on_tiles> <iterations>"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
world, ants, thr, iters = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
os.environ["WR_RANKSET_ON"] = thr
os.environ.setdefault("WR_PEER_TIMEOUT_MS", "60000")
import bench  # noqa: E402
import welding_robot_b200 as wr  # noqa: E402
from welding_robot_b200.dist import LocalShards  # noqa: E402

bench.select_workload("C2")
wl = bench.build_workload_gpu()


def make():
    a = wr.ACS_Rank(seed=1, fixed_colony=ants, step_cap=8192, update_mode=4)
    a.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
    with contextlib.redirect_stdout(io.StringIO()):
        a.initFromGridMap()
    a.setEndpoints(wl["start"], wl["goal"])
    return a


shards = [make() for _ in range(world)]
S = LocalShards(shards); S.begin(1.0)
for it in range(iters):
    S.iterate(1); S.sync()
    st = shards[0].updateStats()
    print(it, st, flush=True)
one = make(); one.begin(1.0); one.iterate(iters); one.sync()
t1 = one.pheromone()
for a in shards:
    assert np.array_equal(t1.view(np.uint32), a.pheromone().view(np.uint32))
print("ok")
