"""GPU, 2 ranks, NCCL: an ant-sharded search must equal the un-sharded search bit for bit
(pheromone field, best path, ranks) on every rank — Philox is keyed by the global ant index and the
deposit merge is order-preserving."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def worker(rank, world, port, q):
    import contextlib
    import io
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import welding_robot_b200 as wr
        from welding_robot_b200 import _lib
        from welding_robot_b200.dist import ShardedSearch
        _lib.check(_lib.lib().wr_set_device(rank))
        tris = np.load(os.path.join(GOLDEN, "meshes.npz"))["simplified_piece"]

        def make():
            a = wr.ACS_Rank(seed=21, fixed_colony=1001, step_cap=600)     # odd colony: ragged last chunk
            with contextlib.redirect_stdout(io.StringIO()):
                a.creatGridMap(tris, 0.012, 4)
                a.initFromGridMap()
            free = np.flatnonzero(a.isfree())
            a.setEndpoints(int(free[11]), int(free[-11]))
            return a

        single = make(); single.begin(1.0); single.iterate(8)
        t1, b1, c1 = single.pheromone(), single.bestPath(), single.counters()
        chunk = (1001 + world - 1) // world
        # deposit paths: sorted records only (owner-computes update), rank sets only (per-rank sets OR-merged over peer
        # memory), the adaptive choice (thresholds set so that both paths and both transitions occur), and the fixed-capacity
        # rank-set table overflowing into the exact serial pass; rendezvous through torch.distributed or through the
        # library's own NCCL bootstrap (wr_comm_unique_id / wr_acs_comm_init)
        for policy, env, own_comm in (("2", {}, False), ("1", {}, False), ("0", {"WR_RANKSET_ON": "100000", "WR_RANKSET_OFF": "2500"}, False),
                                      ("1", {"WR_RANKSET_LOG2": "9"}, False), ("1", {}, True)):
            os.environ["WR_RANKSET_POLICY"] = policy
            os.environ.update(env)
            sharded = make()
            for k in ["WR_RANKSET_POLICY"] + list(env):
                del os.environ[k]
            if own_comm:
                uid = np.zeros(128, np.uint8)
                if rank == 0:
                    _lib.check(_lib.lib().wr_comm_unique_id(uid.ctypes.data))
                box = [uid.tobytes()]
                dist.broadcast_object_list(box, src=0)
                _lib.check(_lib.lib().wr_acs_comm_init(sharded._a, box[0], rank, world))
                sharded.begin(1.0)                      # exchanges the slab handles itself (ncclAllGather)
                sharded.iterate(3); sharded.iterate(5)
            else:
                S = ShardedSearch(sharded, rank, world)
                S.begin(1.0); S.iterate(3); S.iterate(5)
            sharded.sync()
            t2 = sharded.pheromone()
            assert np.array_equal(t1.view(np.uint32), t2.view(np.uint32)), "sharded pheromone field differs from the 1-GPU field (%s, %s)" % (policy, env)
            b2 = sharded.bestPath()
            assert np.array_equal(b1[0], b2[0]) and np.array_equal(b1[1], b2[1]) and np.float32(b1[2]) == np.float32(b2[2])
            for k in range(rank * chunk, min((rank + 1) * chunk, 1001), 37):
                i1, d1, L1, o1 = single.lastAnt(k); i2, d2, L2, o2 = sharded.lastAnt(k)
                assert o1 == o2 and np.array_equal(i1, i2) and (L1 == L2 or (np.isinf(L1) and np.isinf(L2)))
            c2 = sharded.counters()
            tot = torch.tensor([c2["ant_steps"], c2["ants"]], device="cuda", dtype=torch.int64)
            dist.all_reduce(tot)
            assert int(tot[0]) == c1["ant_steps"] and int(tot[1]) == c1["ants"]
            st = sharded.updateStats()
            assert (st["rankset_iterations"] > 0) == (policy != "2"), st
            if policy == "0":
                assert 0 < st["rankset_iterations"] < 8, st
            dist.barrier()
            del sharded
        q.put((rank, "ok", 0))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def run_world(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", "rank %d: %s" % (rank, info)


@pytest.mark.timeout(600)
def test_sharded_equals_single_gpu_world2():
    run_world(2)


@pytest.mark.timeout(600)
def test_sharded_equals_single_gpu_world4():
    """four ranks: three peers to pull final values from, slot slices that do not divide evenly"""
    run_world(4)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("ranks", [2])
def test_cpp_harness_sharded_equals_single_gpu(ranks, tmp_path):
    """examples/sharded_main: the C++ facade + wr_comm_unique_id / wr_acs_comm_init, one process per GPU, no Python in the
    ranks: every rank's pheromone field digest, best length and best path equal the single-GPU run's."""
    import subprocess
    import torch
    from test_oracle_golden import stl_bytes
    if torch.cuda.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    exe = os.path.join(ROOT, "examples", "sharded_main")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True)
    stl = tmp_path / "piece.stl"
    stl.write_bytes(stl_bytes(dict(np.load(os.path.join(GOLDEN, "meshes.npz")))["simplified_piece"]))
    r = subprocess.run([exe, str(stl), str(ranks), "0.012", "4", "2001", "10", str(tmp_path)], capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded == single GPU, bit for bit" in r.stdout
