"""C5 (BASELINE.json configs[4]): 1024 independent start/goal queries on a synthetic 512^3 obstacle grid, 256 ants x 50 iterations
each, through the concurrent path (wr_acs_search_batch) — bench.py's `other_configs.C5` record on its own (one GPU).

    python scripts/c5_queries.py [queries] [iterations] [ants]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ants = int(sys.argv[3]) if len(sys.argv) > 3 else 256
torch.cuda.set_device(0)
print(json.dumps(bench.sub_c5(0, 1, None, nq=nq, iters=iters, ants=ants)))
