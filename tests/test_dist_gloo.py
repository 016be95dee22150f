"""CPU, world_size 2, gloo: the ant-sharding exchange logic of welding_robot_b200/dist.py
(all_gather layout by global ant index, single-contributor integer all_reduce merges, best-owner
logic) driven by a backend built on the CPU oracle.  The merged deposit list, applied in
(slot, rank) order to rho*tau, must reproduce the oracle's pheromone field bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class OracleBackend:
    """Per-rank stand-in for GpuBackend: every rank runs the deterministic oracle iteration but only
    exposes ITS ants' results; everything global must come out of the collectives."""

    def __init__(self, O, A, colony, cap):
        self.O, self.A, self.colony, self.cap = O, A, colony, cap

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world
        self.chunk = (self.colony + world - 1) // world
        self.first = rank * self.chunk

    def begin(self, predict):
        self.A.begin(predict)

    def walk(self):
        self.tau_prev = self.A.pheromone()
        self.A.iterate(1)
        self.ants = [self.A.last_ant(k) for k in range(self.colony)]   # (ids, dirs, L, order)
        local = np.full(self.chunk, -1, np.int32)
        for k in range(self.first, min(self.first + self.chunk, self.colony)):
            ids, dirs, L, order = self.ants[k]
            local[k - self.first] = -1 if np.isinf(L) else len(dirs)
        return torch.from_numpy(local)

    def rank_global(self, all_steps):
        s = all_steps.numpy()[:self.colony].astype(np.int64)
        key = np.where(s < 0, self.cap + 1, s)
        self.order = np.argsort(key, kind="stable")               # (steps, ant index): the oracle's total order
        self.sorted_steps = key[self.order]
        colony, lam, Q = self.A.last_colony()
        self.lam, self.Q = np.float32(lam), np.float32(Q)
        elig = [(r, int(self.order[r])) for r in range(colony)
                if self.sorted_steps[r] <= self.cap and not (np.float32(r + 1) > np.float32(self.lam - np.float32(1)))]
        self.elig = elig
        self.offsets = np.concatenate([[0], np.cumsum([self.sorted_steps[r] for r, _ in elig])]).astype(np.int64)
        cand = np.zeros(2 * self.cap + 2, np.int32)
        best_ids, best_dirs, best_L = self.A.best()
        top = int(self.order[0])
        self.best_is_new = self.sorted_steps[0] <= self.cap and len(best_dirs) == self.sorted_steps[0] and np.array_equal(self.ants[top][0], best_ids)
        if self.best_is_new and self.first <= top < self.first + self.chunk:
            cand[0] = 1
            cand[1:1 + len(best_ids)] = best_ids
            cand[self.cap + 2:self.cap + 2 + len(best_dirs)] = best_dirs
        return torch.from_numpy(cand)

    def apply_best(self):
        pass

    def build_records(self):
        n = int(self.offsets[-1])
        keys = np.zeros(n, np.int32); vals = np.zeros(n, np.int32)
        best_ids, _, best_L = self.A.best()
        onbest = set(int(i) for i in best_ids)
        f = np.float32
        for (r, ant), off in zip(self.elig, self.offsets[:-1]):
            if not (self.first <= ant < self.first + self.chunk):
                continue
            ids, dirs, L, order = self.ants[ant]
            assert order == r + 1
            base = f(f(f(self.lam - f(order)) * self.Q) / f(L))
            elite = f(f(f(f(1) * self.lam) * self.Q) / f(best_L))
            for i, d in enumerate(dirs):
                onb = int(ids[i]) in onbest and int(ids[i + 1]) in onbest
                v = f(base + elite) if onb else f(base + f(0))
                keys[off + i] = np.uint32(int(ids[i]) * 6 + int(d)).astype(np.int32)
                vals[off + i] = np.float32(v).view(np.int32)
        self.keys, self.vals = torch.from_numpy(keys), torch.from_numpy(vals)
        return self.keys, self.vals

    def finish_iteration(self):
        keys = self.keys.numpy().view(np.uint32).astype(np.int64); vals = self.vals.numpy().view(np.float32)
        tau = (self.tau_prev * np.float32(0.8)).astype(np.float32)
        for j in np.argsort(keys, kind="stable"):                 # slot order, rank order inside a slot
            tau[keys[j]] = np.float32(tau[keys[j]] + vals[j])
        assert np.array_equal(tau.view(np.uint32), self.A.pheromone().view(np.uint32)), "merged deposits do not reproduce the oracle field"


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from welding_robot_b200.dist import ShardedSearch
        tris = np.load(os.path.join(GOLDEN, "meshes.npz"))["simplified_piece"]
        G = O.Grid.from_triangles(tris, 0.02, 3, O.VOX_AABB)
        free = np.flatnonzero(G.isfree())
        colony, cap = 45, 400                                     # odd colony: the last rank's chunk is ragged
        A = O.Acs(G, seed=13, fixed_colony=colony, step_cap=cap)
        A.set_endpoints(int(free[7]), int(free[-7]))
        backend = OracleBackend(O, A, colony, cap)
        S = ShardedSearch(None, rank, world, backend=backend)
        S.begin(1.0)
        S.iterate(4)
        # global views must agree across ranks
        t = torch.from_numpy(np.array([S.bytes_exchanged, int(backend.offsets[-1])], np.int64))
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        q.put((rank, "ok", S.bytes_exchanged))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", "rank %d: %s" % (rank, info)
    assert res[0][2] == res[1][2] and res[0][2] > 0
