"""C4 (BASELINE.json configs[3]): ACS_GTSP seam ordering, N = 256 cities, 1024 colonies batched.
Prints colony-iterations/s and ant-steps/s of K4 next to the CPU oracle on one core (bounded sample)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import welding_robot_b200 as wr  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 8

P = np.random.default_rng(3).random((N, 3))
D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
D = np.array([[float("%.6f" % v) for v in row] for row in D])   # as a graph file would carry it

g = wr.ACS_GTSP(seed=3)
g.dis, g.city_num, g.cnt = D, N, N * (N - 1) // 2
g._create(B, 0)
g.iterate(1)                      # warm-up
t0 = time.perf_counter()
g.iterate(ITERS)
dt = time.perf_counter() - t0
ms = g.kernelMs()
best = [g.best(b)[1] for b in range(0, B, max(1, B // 8))]

# CPU oracle, one colony, same iterations (single thread like the reference)
T = O.Gtsp(D, seed=3, colony_id=0)
c0 = time.perf_counter()
T.iterate(min(ITERS + 1, 3))
cdt = (time.perf_counter() - c0) / min(ITERS + 1, 3)
tour, L = T.best()
ok = None
if ITERS + 1 <= 3:
    ok = bool(np.array_equal(g.best(0)[0], tour))

out = {
    "workload": "C4: ACS_GTSP N=%d, %d colonies, %d iterations" % (N, B, ITERS),
    "gpu_colony_iterations_per_s": B * ITERS / dt,
    "gpu_ant_steps_per_s": B * ITERS * N * N / dt,
    "gpu_seconds": dt,
    "gpu_phase_ms": ms,
    "info_matrix_bytes_streamed_per_s": B * ITERS * N * (N * N * 8) / dt,
    "cpu_oracle_colony_iterations_per_s_1core": 1.0 / cdt,
    "sample_best_L": best,
    "parity_colony0": ok,
}
print(json.dumps(out))
