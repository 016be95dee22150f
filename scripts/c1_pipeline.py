"""C1 (BASELINE.json configs[0]): the reference's demo-3 planning pipeline on files/cubic.stl with its default parameters —
voxel grid (0.005, wall 10), all 15 pairs of the six SURVEY weld points x 150 iterations with the adaptive colony, seam
ordering — through the Python twin of the reference's class surface, timed end to end; the unmodified reference's own
all-pairs search (oracle/_ref) is timed beside it on one host core when present.

    python scripts/c1_pipeline.py [ref 0|1]
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import welding_robot_b200 as wr  # noqa: E402

POINTS = [(1.600931, y, z) for y in (-0.259319, -0.074319, 0.085681) for z in (1.224003, 1.399003)]
tris = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))["cubic"]
tmp = tempfile.mkdtemp()
pts = os.path.join(tmp, "weld_points.in"); graph = os.path.join(tmp, "graph.in")
with open(pts, "w") as f:
    f.write("%d\n" % len(POINTS) + "".join("%.6f %.6f %.6f\n" % p for p in POINTS))


def run():
    sink = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        a = wr.ACS_Rank(seed=0x5EED, update_mode=4)
        a.creatGridMap(tris, 0.005, 10)
        t1 = time.perf_counter()
        a.searchBestPathOfPoints(0.5, pts, graph)
        a.sync()
        t2 = time.perf_counter()
        g = wr.ACS_GTSP(seed=3)
        g.readFromGraphFile(graph)
        g.computeSolution()
        g.read_all_segments(a.best_matrix)
        t3 = time.perf_counter()
    c = a.counters()
    lengths = [float(a.best_matrix[i][j].L) for i in range(6) for j in range(i + 1, 6)]
    return dict(grid_s=t1 - t0, search_s=t2 - t1, gtsp_s=t3 - t2, total_s=t3 - t0, ant_steps=c["ant_steps"], ants=c["ants"],
                iterations=c["iterations"], lengths=lengths, stitched_points=len(g.g_path_x))


run()                       # warm-up: module load, memory pools
r = run()
out = {"workload": "C1: cubic.stl @ (0.005, wall 10) = 57x91x57 nodes, 15 pairs x 150 iterations, adaptive colony (<= 35 ants), reference defaults",
       "gpu": r, "gpu_ant_steps_per_s": r["ant_steps"] / r["search_s"], "gpu_acs_iterations_per_s": r["iterations"] / r["search_s"]}
if (len(sys.argv) < 2 or sys.argv[1] != "0"):
    from oracle import oracle as O
    if O.have_ref():
        R = O.Ref()
        t0 = time.perf_counter()
        R.voxelize(tris, 0.005, 10)
        R.acs_init()
        t1 = time.perf_counter()
        calls = 0
        for i in range(6):
            for j in range(i + 1, 6):
                assert R.set_points(POINTS[i], POINTS[j])[0]
                calls += int(R.compute(0.5, 150, 0x5EED + 7 * (6 * i + j)))
                R.reset()
        t2 = time.perf_counter()
        out["reference_cpu_1core"] = {"grid_and_init_s": t1 - t0, "search_s": t2 - t1, "ant_steps": calls, "ant_steps_per_s": calls / (t2 - t1),
                                      "acs_iterations_per_s": 15 * 150 / (t2 - t1)}
        out["search_speedup_vs_reference_1core"] = (t2 - t1) / r["search_s"]
print(json.dumps(out))
