"""Ant-sharded search across ranks: one process per GPU, torch.distributed for the plumbing.

SURVEY.md §8e: within an iteration the pheromone field is read-only, so ants are independent units with ONE
exchange step per iteration.  The grid and the pheromone field are replicated on every GPU; rank r constructs ants
[r*chunk, (r+1)*chunk) of the global colony (Philox is keyed by the global ant index).

PEER protocol (default on NVLink/NVSwitch boxes) — one small collective per iteration, no host synchronisation:

  0. once per search: every rank exports ONE CUDA IPC handle of the slab that holds what its peers read (ant trails
     and the list of final slot values, double-buffered by iteration parity)            wr_acs_peer_export / _import
  1. local ant construction (K2), trails written into the slab                          wr_acs_walk
  2. all_gather of per-ant step counts (4 B per ant) — also the barrier that makes the trails visible
  3. every rank: global ranking + best decision; the new best trail and the trails of ALL eligible ants are read
     straight out of their owners' HBM over NVLink (kernel-side peer loads) and the deposit records are generated
     locally in global (rank, step) order                                               wr_acs_finish_iteration_peer
  4a. sliced = False: slot sort + fused evaporation/deposit (K3) of all records, identical on every rank
  4b. sliced = True (default): owner-computes update — the slot space is cut into one tile-aligned slice per rank;
      every rank keeps (stable partition), sorts and applies only ITS slice's records (1/world of the sort and of the
      dependent add chains) while evaporating the whole field, and lists the final value of every slot it touched;
      barrier; every rank pulls the peers' lists out of their HBM and overwrites those slots  wr_acs_pull_finals

NCCL-only protocol (`peer=False`; also what the CPU/gloo test drives): steps 3-4 become
  all_reduce(SUM, int32) of the best-candidate buffer (only the owner of the new best ant holds non-zero words),
  deposit records of the local ants at their GLOBAL positions + all_reduce(SUM, int32) of keys and values (every
  position has exactly one non-zero contributor, so integer SUM is a merge), replicated sort + update.

Either way the deposit list equals the single-GPU list, so the pheromone field stays bit-identical on all ranks and
to a 1-GPU run — no floating-point reduction order is involved.  (The dense alternative — all_reduce of a delta
field, 3.2 GB at 512^3 — moves 20-200x more bytes over NVLink and is not reproducible.)

The exchange logic is written against a small backend interface so that it runs under gloo on CPU in tests
(tests/test_dist_gloo.py); `GpuBackend` is the product backend over the C ABI; `LocalShards` drives all shards of a
colony from one process on one GPU (tests/test_gpu_shards_local.py).
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

    def __init__(self, acs, bind_stream=True):
        self.acs = acs
        self.h = acs._need()
        self.device = torch.device("cuda", torch.cuda.current_device())
        # wr_acs_create gives every handle a private non-blocking stream, but the collectives of torch.distributed are
        # ordered against torch's CURRENT stream only: the handle has to run on that stream, or the all_gather could read
        # the step counts before the walk has written them (and peers could read trails that are still being written).
        if bind_stream:
            check(lib().wr_acs_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def set_shard(self, rank, world):
        check(lib().wr_acs_set_shard(self.h, rank, world))

    def begin(self, predict):
        self.acs.begin(predict)
        self._local = None

    def walk(self):
        check(lib().wr_acs_walk(self.h))
        if getattr(self, "_local", None) is None:   # the buffer of this rank's step counts does not move during a search: wrap it once
            p = C.c_void_p(); first = C.c_int(); count = C.c_int()
            check(lib().wr_acs_local_steps_dev(self.h, C.byref(p), C.byref(first), C.byref(count)))
            self._local = _view(p.value, count.value, self.device)
        return self._local

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

    # ---- NVLink peer-memory protocol ------------------------------------------------------------------
    def export_handle(self):
        """-> (64 bytes: CUDA IPC handle of this rank's slab, raw device pointer of the slab)."""
        h = (C.c_ubyte * 64)(); raw = C.c_void_p()
        check(lib().wr_acs_peer_export(self.h, h, C.byref(raw)))
        return bytes(h), raw.value

    def peer_setup(self, world, group=None):
        """Exchange the IPC handles of the slabs (once per begin)."""
        mine, _ = self.export_handle()
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
        allh = torch.empty(64 * world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, t, group=group)
        check(lib().wr_acs_peer_import(self.h, allh.cpu().numpy().tobytes()))

    def set_peer_pointers(self, all_raw):
        arr = (C.c_void_p * len(all_raw))(*all_raw)
        check(lib().wr_acs_peer_set_pointers(self.h, arr))

    def finish_iteration_peer(self, all_steps, sliced):
        if getattr(self, "_all_ptr", None) is None or self._all_src is not all_steps:
            self._all_src, self._all_ptr = all_steps, C.c_void_p(all_steps.data_ptr())
        check(lib().wr_acs_finish_iteration_peer(self.h, self._all_ptr, 1 if sliced else 0))

    def pull_finals(self):
        check(lib().wr_acs_pull_finals(self.h))


class ShardedSearch:
    """Drives one ant-sharded search.  `backend` defaults to the GPU backend of `acs`.
    peer:   NVLink peer-memory protocol; None = on when the backend supports it (WR_SHARD_PEER=0 disables).
    sliced: owner-computes update (step 4b); None = on with the peer protocol (WR_SHARD_SLICED=0 disables)."""

    def __init__(self, acs, rank, world, group=None, backend=None, peer=None, sliced=None):
        import os
        self.rank, self.world, self.group = rank, world, group
        self.backend = backend if backend is not None else GpuBackend(acs)
        self.backend.set_shard(rank, world)
        if peer is None:
            peer = hasattr(self.backend, "finish_iteration_peer") and os.environ.get("WR_SHARD_PEER", "1") != "0"
        self.peer = bool(peer) and world > 1
        if sliced is None:
            sliced = os.environ.get("WR_SHARD_SLICED", "1") != "0"
        self.sliced = bool(sliced) and self.peer
        self._all = None
        self._bar = None
        self.bytes_exchanged = 0     # through collectives (peer loads are not counted here)

    def begin(self, predict_path_len):
        self.backend.begin(predict_path_len)
        if self.peer:
            self.backend.peer_setup(self.world, self.group)

    def iterate(self, n=1):
        b = self.backend
        for _ in range(n):
            local = b.walk()                                   # int32[chunk], -1 = dead / beyond the colony
            if self._all is None or self._all.numel() != local.numel() * self.world:
                self._all = torch.empty(local.numel() * self.world, dtype=local.dtype, device=local.device)
                self._bar = torch.zeros(1, dtype=torch.int32, device=local.device)
            dist.all_gather_into_tensor(self._all, local, group=self.group)
            if self.peer:
                b.finish_iteration_peer(self._all, self.sliced)
                if self.sliced:
                    dist.all_reduce(self._bar, op=dist.ReduceOp.SUM, group=self.group)   # barrier: every rank's list is complete
                    b.pull_finals()
                self.bytes_exchanged += 4 * (self._all.numel() + (1 if self.sliced else 0))
                continue
            cand = b.rank_global(self._all)                    # zeros unless this rank owns the new best ant
            dist.all_reduce(cand, op=dist.ReduceOp.SUM, group=self.group)
            b.apply_best()
            keys, vals = b.build_records()                     # this rank's records at global positions, zeros elsewhere
            if keys.numel():
                dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
            b.finish_iteration()
            self.bytes_exchanged += 4 * (self._all.numel() + cand.numel() + 2 * keys.numel())


class LocalShards:
    """The peer protocol for `world` shards of one colony that live in ONE process on ONE GPU (tests, debugging): the
    all_gather becomes a concatenation on the shared stream, the peer slabs plain device pointers."""

    def __init__(self, searches, sliced=True):
        self.world = len(searches)
        self.sliced = sliced
        self.backends = [GpuBackend(a) for a in searches]
        stream = torch.cuda.current_stream().cuda_stream
        for r, b in enumerate(self.backends):
            check(lib().wr_acs_set_stream(b.h, C.c_void_p(stream)))
            b.set_shard(r, self.world)

    def begin(self, predict_path_len):
        raws = []
        for b in self.backends:
            b.begin(predict_path_len)
            raws.append(b.export_handle()[1])
        for b in self.backends:
            b.set_peer_pointers(raws)

    def iterate(self, n=1):
        bs = self.backends
        for _ in range(n):
            allsteps = torch.cat([b.walk() for b in bs])
            for b in bs:
                b.finish_iteration_peer(allsteps, self.sliced)
            if self.sliced:
                for b in bs:
                    b.pull_finals()
