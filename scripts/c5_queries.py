"""C5 (BASELINE.json configs[4]): independent start/goal queries on a synthetic 512^3 obstacle grid, 256 ants x 50 iterations
per query, all on one GPU through wr_acs_search_pairs (the all-pairs driver: begin + iterate + reset per query, no host sync).
Queries shard across GPUs with no communication (query q -> GPU q mod N), so one GPU's rate times N is the box's rate.

    python scripts/c5_queries.py [queries] [lazy 0|1]

Prints one JSON line: queries/s, ant-steps/s, ms per query, tiles the evaporation pass streamed.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 32
LAZY = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if not LAZY:
    os.environ["WR_LAZY_TAU"] = "0"
import welding_robot_b200 as wr  # noqa: E402

N = 512


def synthetic_boxes(n, nboxes, seed):
    """union of axis-aligned boxes (edges 4-48 cells), 10-cell free wall, coordinates = indices"""
    rng = np.random.default_rng(seed)
    occ = np.zeros((n, n, n), bool)
    for _ in range(nboxes):
        e = rng.integers(4, 49, 3)
        c = [int(rng.integers(10, n - 10 - int(e[k]))) for k in range(3)]
        occ[c[2]:c[2] + e[2], c[1]:c[1] + e[1], c[0]:c[0] + e[0]] = True
    return (~occ).astype(np.uint8).ravel()


free = synthetic_boxes(N, 4096, 4)
axis = np.arange(N, dtype=np.float32)
ids = np.flatnonzero(free)
rng = np.random.default_rng(5)
starts, goals = [], []
while len(starts) < NQ:
    s = int(ids[rng.integers(0, len(ids))])
    sz, sy, sx = s // (N * N), (s // N) % N, s % N
    d = int(rng.integers(64, 257))
    # a free cell at Manhattan distance d: random split of d over the axes
    for _ in range(64):
        a = rng.multinomial(d, [1 / 3] * 3) * rng.choice([-1, 1], 3)
        gz, gy, gx = sz + a[0], sy + a[1], sx + a[2]
        if 0 <= gz < N and 0 <= gy < N and 0 <= gx < N and free[(gz * N + gy) * N + gx]:
            starts.append(s); goals.append(int((gz * N + gy) * N + gx))
            break

g = wr.ACS_Rank(seed=5, fixed_colony=256, step_cap=4096, update_mode=4)
g.creatFromOccupancy(free, axis, axis, axis, 1.0)
with contextlib.redirect_stdout(io.StringIO()):
    g.initFromGridMap()
g.searchPairs(starts[:2], goals[:2], 300.0, 50, with_paths=False)        # warm-up: pools, heuristic buffers
g.sync()
c0 = g.counters()
t0 = time.perf_counter()
res = g.searchPairs(starts, goals, 300.0, 50, with_paths=False)
g.sync()
L = np.array([r[2] for r in res], np.float32)
dt = time.perf_counter() - t0
c1 = g.counters()
steps = c1["ant_steps"] - c0["ant_steps"]
print(json.dumps({
    "workload": "C5: %d queries on a synthetic 512^3 obstacle grid (%.1f %% occupied), 256 ants x 50 iterations each, Manhattan separation 64-256"
                % (NQ, 100.0 * (1 - free.mean())),
    "clean_tile_field": bool(LAZY), "queries_per_s": NQ / dt, "ms_per_query": 1e3 * dt / NQ, "ant_steps_per_s": steps / dt,
    "acs_iterations_per_s": 50 * NQ / dt, "found": int(np.isfinite(L).sum()), "mean_best_L": float(np.mean(L[np.isfinite(L)])) if np.isfinite(L).any() else None,
    "field_tiles": g.fieldStats()[1], "pheromone_field_bytes": N ** 3 * 6 * 4}))
