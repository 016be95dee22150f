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
  6. slot sort + fused evaporation/deposit (K3), identical on every rank  wr_acs_finish_iteration

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


class ShardedSearch:
    """Drives one ant-sharded search.  `backend` defaults to the GPU backend of `acs`."""

    def __init__(self, acs, rank, world, group=None, backend=None):
        self.rank, self.world, self.group = rank, world, group
        self.backend = backend if backend is not None else GpuBackend(acs)
        self.backend.set_shard(rank, world)
        self._all = None
        self.bytes_exchanged = 0

    def begin(self, predict_path_len):
        self.backend.begin(predict_path_len)

    def iterate(self, n=1):
        b = self.backend
        for _ in range(n):
            local = b.walk()                                   # int32[chunk], -1 = dead / beyond the colony
            if self._all is None or self._all.numel() != local.numel() * self.world:
                self._all = torch.empty(local.numel() * self.world, dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(self._all, local, group=self.group)
            cand = b.rank_global(self._all)                    # zeros unless this rank owns the new best ant
            dist.all_reduce(cand, op=dist.ReduceOp.SUM, group=self.group)
            b.apply_best()
            keys, vals = b.build_records()                     # this rank's records at global positions, zeros elsewhere
            if keys.numel():
                dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
            b.finish_iteration()
            self.bytes_exchanged += 4 * (self._all.numel() + cand.numel() + 2 * keys.numel())
