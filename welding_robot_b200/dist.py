"""Ant-sharded search across ranks: one process per GPU.

SURVEY.md §8e: within an iteration the pheromone field is read-only, so ants are independent units with ONE exchange
step per iteration.  The grid and the pheromone field are replicated on every GPU; rank r constructs ants
[r*chunk, (r+1)*chunk) of the global colony (Philox is keyed by the global ant index).

The whole iteration loop lives in libwrgpu.so (wr_acs_iterate on a sharded handle, include/wr_gpu.h): the exchange runs
over NVLink peer memory inside the library's own kernels — barrier flags, step counts, ant trails, published rank-set
blocks and final slot values all sit in one slab per rank — with no collective and no host synchronisation per
iteration.  What is left for the host is the rendezvous: once per search every rank hands the others the CUDA IPC handle
of its slab.  This module does that with torch.distributed (any backend; 64 bytes per rank); a C++ host does the same
with wr_comm_unique_id / wr_acs_comm_init (NCCL opened by the library itself).

The deposit list of a sharded search equals the single-GPU list, so the pheromone field stays bit-identical on all
ranks and to a 1-GPU run — no floating-point reduction order is involved.  (The dense alternative — all_reduce of a
delta field, 3.2 GB at 512^3 — moves 20-200x more bytes over NVLink and is not reproducible.)

`LocalShards` drives all shards of a colony from one process on one GPU (tests/test_gpu_shards_local.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib


def shard_bounds(colony, rank, world):
    """Global ant indices [first, last) that rank `rank` constructs (chunk = ceil(colony / world), ragged last chunk)."""
    chunk = (max(colony, 1) + world - 1) // world
    return min(rank * chunk, colony), min((rank + 1) * chunk, colony)


def shard_queries(nqueries, rank, world):
    """Independent start/goal queries (BASELINE config 5) shard with no communication: query q runs on rank q mod world."""
    return np.arange(rank, nqueries, world)


def exchange_handles(mine: bytes, world, group=None, device=None):
    """all_gather of one 64-byte CUDA IPC handle per rank -> world * 64 bytes in rank order (works under nccl and gloo)."""
    assert len(mine) == 64
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, t, group=group)
    return allh.cpu().numpy().tobytes()


class GpuBackend:
    """Per-rank handle over libwrgpu.so."""

    def __init__(self, acs):
        self.acs = acs
        self.h = acs._need()

    def set_shard(self, rank, world):
        check(lib().wr_acs_set_shard(self.h, rank, world))

    def begin(self, predict):
        self.acs.begin(predict)

    def export_handle(self):
        """-> (64 bytes: CUDA IPC handle of this rank's slab, raw device pointer of the slab)."""
        h = (C.c_ubyte * 64)(); raw = C.c_void_p()
        check(lib().wr_acs_peer_export(self.h, h, C.byref(raw)))
        return bytes(h), raw.value

    def import_handles(self, all_handles: bytes):
        check(lib().wr_acs_peer_import(self.h, all_handles))

    def set_peer_pointers(self, all_raw):
        arr = (C.c_void_p * len(all_raw))(*all_raw)
        check(lib().wr_acs_peer_set_pointers(self.h, arr))

    def iterate(self, n):
        check(lib().wr_acs_iterate(self.h, n))


class ShardedSearch:
    """Drives one ant-sharded search: rendezvous through torch.distributed, iterations through wr_acs_iterate."""

    def __init__(self, acs, rank, world, group=None, backend=None):
        self.rank, self.world, self.group = rank, world, group
        self.backend = backend if backend is not None else GpuBackend(acs)
        self.backend.set_shard(rank, world)
        self.bytes_exchanged = 0     # through collectives (peer loads are not counted here)

    def begin(self, predict_path_len):
        b = self.backend
        b.begin(predict_path_len)
        if self.world > 1:
            mine, _ = b.export_handle()
            b.import_handles(exchange_handles(mine, self.world, self.group))
            self.bytes_exchanged += 64 * self.world

    def iterate(self, n=1):
        self.backend.iterate(n)


class LocalShards:
    """`world` shards of one colony that live in ONE process on ONE GPU (tests, debugging): the peer slabs are plain
    device pointers.  Every shard runs on its own stream (a barrier kernel waits for the other shards' kernels, which
    must be able to run beside it) and the shards are iterated in turn, one iteration at a time (a shard may run at
    most four iterations ahead of the device, and its iterations cannot finish before the others' are enqueued)."""

    def __init__(self, searches):
        self.world = len(searches)
        self.backends = [GpuBackend(a) for a in searches]
        for r, b in enumerate(self.backends):
            b.set_shard(r, self.world)

    def begin(self, predict_path_len):
        raws = []
        for b in self.backends:
            b.begin(predict_path_len)
            raws.append(b.export_handle()[1])
        for b in self.backends:
            b.set_peer_pointers(raws)

    def iterate(self, n=1):
        for _ in range(n):
            for b in self.backends:
                b.iterate(1)

    def sync(self):
        for b in self.backends:
            b.acs.sync()
