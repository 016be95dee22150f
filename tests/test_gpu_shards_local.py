"""GPU, one device: the peer-memory protocol of a sharded colony (trails read through peer pointers, replicated or
owner-computes update with final-value lists — welding_robot_b200/dist.py) with all shards living in ONE process,
against the un-sharded search.
Covers the device side of the multi-GPU protocol on a 1-GPU box; tests/test_gpu_multi.py runs it over NCCL + CUDA IPC."""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,colony,sliced", [(2, 1001, True), (3, 640, True), (2, 333, False)])
def test_peer_protocol_shards_equal_unsharded(world, colony, sliced):
    import torch
    import welding_robot_b200 as wr
    from welding_robot_b200 import _lib
    from welding_robot_b200.dist import LocalShards
    tris = np.load(os.path.join(GOLDEN, "meshes.npz"))["simplified_piece"]

    def make():
        a = wr.ACS_Rank(seed=21, fixed_colony=colony, step_cap=600)     # odd colony: ragged last chunk
        with contextlib.redirect_stdout(io.StringIO()):
            a.creatGridMap(tris, 0.012, 4)
            a.initFromGridMap()
        _lib.check(_lib.lib().wr_acs_set_stream(a._a, torch.cuda.current_stream().cuda_stream))
        free = np.flatnonzero(a.isfree())
        a.setEndpoints(int(free[11]), int(free[-11]))
        return a

    single = make(); single.begin(1.0)
    shards = [make() for _ in range(world)]
    S = LocalShards(shards, sliced=sliced); S.begin(1.0)
    for its in (1, 1, 6):
        single.iterate(its); S.iterate(its)
        torch.cuda.synchronize()
        t1 = single.pheromone()
        b1 = single.bestPath()
        for a in shards:
            assert np.array_equal(t1.view(np.uint32), a.pheromone().view(np.uint32)), "sharded pheromone field differs from the un-sharded field"
            b2 = a.bestPath()
            assert np.array_equal(b1[0], b2[0]) and np.array_equal(b1[1], b2[1]) and np.float32(b1[2]) == np.float32(b2[2])
    c1 = single.counters()
    assert sum(a.counters()["ant_steps"] for a in shards) == c1["ant_steps"]
    assert c1["deposit_records"] > 0 and all(a.counters()["deposit_records"] == c1["deposit_records"] for a in shards)
