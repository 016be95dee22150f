"""Ant-sharded search across ranks: one process per GPU, torch.distributed for the plumbing.

SURVEY.md §8e: within an iteration the pheromone field is read-only, so ants are independent units
with ONE exchange step per iteration.  The grid and the pheromone field are replicated on every GPU;
rank r constructs ants [r*chunk, (r+1)*chunk) of the global colony (Philox is keyed by the global
ant index).  Per iteration:

  1. local ant construction (K2)                                         wr_acs_walk
  2. all_gather of per-ant step counts (4 B per ant)                     -> every rank ranks the whole colony
  3. ranking + best decision, identical on every rank                    wr_acs_rank_global
  4. all_reduce(SUM, int32) of the best-candidate buffer (only the owner of the new best ant
     holds non-zero words)                                               wr_acs_apply_best
  5. deposit records of the local ants at their GLOBAL (rank, step) positions, zeros elsewhere,
     all_reduce(SUM, int32) of keys and values                           wr_acs_build_records
  6. owner-computes update (default): the slot space is cut into one tile-aligned slice per rank; every rank
     keeps the merged records of ITS slice (stable partition), sorts and applies them while evaporating the
     whole field, and lists the final value of every slot it touched       wr_acs_finish_iteration_sliced
  7. barrier, then every rank pulls the peers' lists straight out of their HBM over NVLink (kernel-side peer
     loads through CUDA IPC pointers, no host-sized collective)             wr_acs_pull_finals
     (`sliced=False`: step 6 = slot sort + fused update of ALL records on every rank, wr_acs_finish_iteration)

Every reduced position has exactly one non-zero contributor, so integer SUM is a merge, the merged
deposit list equals the single-GPU list and the pheromone field stays bit-identical on all ranks and
to a 1-GPU run — no floating-point reduction order is involved.  (The dense alternative — all_reduce
of a delta field, 3.2 GB at 512^3 — moves 20-200x more bytes over NVLink and is not reproducible.)

The exchange logic is written against a small backend interface so that it runs under gloo on CPU
in tests (tests/test_dist_gloo.py); `GpuBackend` is the product backend over the C ABI.
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def _view(ptr, n, device):
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=device)
    return torch.as_tensor(_DevArray(ptr, n), device=device)


class GpuBackend:
    """Per-rank compute over libwrgpu.so; every method is asynchronous on the handle's stream except
    build_records (reads the record count back)."""

    def __init__(self, acs):
        self.acs = acs
        self.h = acs._need()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def set_shard(self, rank, world):
        check(lib().wr_acs_set_shard(self.h, rank, world))

    def begin(self, predict):
        self.acs.begin(predict)

    def walk(self):
        check(lib().wr_acs_walk(self.h))
        p = C.c_void_p(); first = C.c_int(); count = C.c_int()
        check(lib().wr_acs_local_steps_dev(self.h, C.byref(p), C.byref(first), C.byref(count)))
        return _view(p.value, count.value, self.device)

    def rank_global(self, all_steps):
        check(lib().wr_acs_rank_global(self.h, C.c_void_p(all_steps.data_ptr())))
        p = C.c_void_p(); n = C.c_size_t()
        check(lib().wr_acs_best_candidate_dev(self.h, C.byref(p), C.byref(n)))
        return _view(p.value, n.value, self.device)

    def apply_best(self):
        check(lib().wr_acs_apply_best(self.h))

    def build_records(self):
        k = C.c_void_p(); v = C.c_void_p(); n = C.c_int()
        check(lib().wr_acs_build_records(self.h, C.byref(k), C.byref(v), C.byref(n)))
        return _view(k.value, n.value, self.device), _view(v.value, n.value, self.device)

    def finish_iteration(self):
        check(lib().wr_acs_finish_iteration(self.h))

    # ---- owner-computes deposits ----------------------------------------------------------------------
    def export_handles(self):
        """-> (128 bytes of CUDA IPC handles, [2 raw device pointers]) of this rank's two final-value lists."""
        h = (C.c_ubyte * 128)(); raw = (C.c_void_p * 2)()
        check(lib().wr_acs_peer_export(self.h, h, raw))
        return bytes(h), [raw[0], raw[1]]

    def peer_setup(self, world, group=None):
        """Exchange the IPC handles of the final-value lists (once per begin)."""
        mine, _ = self.export_handles()
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
        allh = torch.empty(128 * world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, t, group=group)
        buf = allh.cpu().numpy().tobytes()
        check(lib().wr_acs_peer_import(self.h, buf))

    def set_peer_pointers(self, all_raw):
        arr = (C.c_void_p * len(all_raw))(*all_raw)
        check(lib().wr_acs_peer_set_pointers(self.h, arr))

    def finish_iteration_sliced(self):
        check(lib().wr_acs_finish_iteration_sliced(self.h))

    def pull_finals(self):
        check(lib().wr_acs_pull_finals(self.h))


class ShardedSearch:
    """Drives one ant-sharded search.  `backend` defaults to the GPU backend of `acs`.
    sliced: owner-computes update (steps 6-7 above); None = on when the backend supports it (WR_SHARD_SLICED=0 disables)."""

    def __init__(self, acs, rank, world, group=None, backend=None, sliced=None):
        import os
        self.rank, self.world, self.group = rank, world, group
        self.backend = backend if backend is not None else GpuBackend(acs)
        self.backend.set_shard(rank, world)
        if sliced is None:
            sliced = world > 1 and hasattr(self.backend, "finish_iteration_sliced") and os.environ.get("WR_SHARD_SLICED", "1") != "0"
        self.sliced = bool(sliced) and world > 1
        self._all = None
        self._bar = None
        self.bytes_exchanged = 0

    def begin(self, predict_path_len):
        self.backend.begin(predict_path_len)
        if self.sliced:
            self.backend.peer_setup(self.world, self.group)

    def iterate(self, n=1):
        b = self.backend
        for _ in range(n):
            local = b.walk()                                   # int32[chunk], -1 = dead / beyond the colony
            if self._all is None or self._all.numel() != local.numel() * self.world:
                self._all = torch.empty(local.numel() * self.world, dtype=local.dtype, device=local.device)
                self._bar = torch.zeros(1, dtype=torch.int32, device=local.device)
            dist.all_gather_into_tensor(self._all, local, group=self.group)
            cand = b.rank_global(self._all)                    # zeros unless this rank owns the new best ant
            dist.all_reduce(cand, op=dist.ReduceOp.SUM, group=self.group)
            b.apply_best()
            keys, vals = b.build_records()                     # this rank's records at global positions, zeros elsewhere
            if keys.numel():
                dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
            if self.sliced:
                b.finish_iteration_sliced()
                dist.all_reduce(self._bar, op=dist.ReduceOp.SUM, group=self.group)   # barrier: every rank's list is complete
                b.pull_finals()
            else:
                b.finish_iteration()
            self.bytes_exchanged += 4 * (self._all.numel() + cand.numel() + 2 * keys.numel() + 1)


class LocalShards:
    """The same protocol for `world` shards of one colony that live in ONE process on ONE GPU (tests, debugging): the
    collectives become tensor operations on the shared stream, the peer buffers plain device pointers."""

    def __init__(self, searches):
        self.world = len(searches)
        self.backends = [GpuBackend(a) for a in searches]
        stream = torch.cuda.current_stream().cuda_stream
        for r, b in enumerate(self.backends):
            check(lib().wr_acs_set_stream(b.h, C.c_void_p(stream)))
            b.set_shard(r, self.world)

    def begin(self, predict_path_len):
        raws = []
        for b in self.backends:
            b.begin(predict_path_len)
            raws += b.export_handles()[1]
        for b in self.backends:
            b.set_peer_pointers(raws)

    def iterate(self, n=1):
        bs = self.backends
        for _ in range(n):
            allsteps = torch.cat([b.walk() for b in bs])
            cands = [b.rank_global(allsteps) for b in bs]
            tot = torch.stack(cands).sum(0, dtype=torch.int32)
            for b, c in zip(bs, cands):
                c.copy_(tot)
                b.apply_best()
            recs = [b.build_records() for b in bs]
            if recs[0][0].numel():
                ks = torch.stack([k for k, _ in recs]).sum(0, dtype=torch.int32)
                vs = torch.stack([v for _, v in recs]).sum(0, dtype=torch.int32)
                for k, v in recs:
                    k.copy_(ks); v.copy_(vs)
            for b in bs:
                b.finish_iteration_sliced()
            for b in bs:
                b.pull_finals()
