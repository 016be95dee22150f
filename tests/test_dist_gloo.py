"""CPU, world_size 2, gloo: the host side of a sharded search (welding_robot_b200/dist.py).  The per-iteration exchange
runs inside libwrgpu.so over peer memory (tests/test_gpu_multi.py, tests/test_gpu_shards_local.py); what the host does
is the rendezvous — every rank hands every other rank the 64-byte IPC handle of its slab, in rank order, once per
search — and the partition arithmetic.  Both are driven here over gloo with a recording backend."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class RecordingBackend:
    """Stand-in for GpuBackend: same methods, no GPU."""

    def __init__(self):
        self.calls = []
        self.imported = None

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world
        self.calls.append(("set_shard", rank, world))

    def begin(self, predict):
        self.calls.append(("begin", predict))

    def export_handle(self):
        self.calls.append(("export",))
        return bytes([(self.rank * 37 + i) & 0xFF for i in range(64)]), 0

    def import_handles(self, allh):
        self.calls.append(("import",))
        self.imported = allh

    def iterate(self, n):
        self.calls.append(("iterate", n))


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from welding_robot_b200.dist import ShardedSearch, shard_queries
        b = RecordingBackend()
        S = ShardedSearch(None, rank, world, backend=b)
        for search in range(2):                      # the slabs are exchanged again at every begin
            S.begin(1.5)
            assert b.imported == b"".join(bytes([(r * 37 + i) & 0xFF for i in range(64)]) for r in range(world))
            S.iterate(3); S.iterate(2)
        assert b.calls == [("set_shard", rank, world)] + [("begin", 1.5), ("export",), ("import",), ("iterate", 3), ("iterate", 2)] * 2
        assert S.bytes_exchanged == 2 * 64 * world
        # independent queries (BASELINE config 5): query q -> rank q mod world; the gathered results come back in query order
        mine = shard_queries(11, rank, world)
        assert list(mine) == list(range(rank, 11, world))
        got = [None] * world
        dist.all_gather_object(got, [(int(qi), float(qi) * 0.5) for qi in mine])
        merged = sorted(x for part in got for x in part)
        assert [m[0] for m in merged] == list(range(11))
        q.put((rank, "ok", ""))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_rendezvous_and_partition_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, status, info in res:
        assert status == "ok", "rank %d: %s" % (rank, info)


def test_shard_bounds_cover_the_colony_exactly_once():
    sys.path.insert(0, ROOT)
    from welding_robot_b200.dist import shard_bounds
    for colony in (0, 1, 7, 1001, 4096, 65536):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(colony, np.int32)
            chunk = (max(colony, 1) + world - 1) // world
            for r in range(world):
                lo, hi = shard_bounds(colony, r, world)
                assert 0 <= lo <= hi <= colony and hi - lo <= chunk
                assert lo == min(r * chunk, colony)     # rank r owns the global ants [r*chunk, (r+1)*chunk): wr_gpu.h
                seen[lo:hi] += 1
            assert (seen == 1).all()
