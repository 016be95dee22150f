"""GPU: the C++ facade (include/welding_robot_b200/welding_robot.hpp, the reference's class surface)
through the headless harness that mirrors main.cpp:273-283 — and the Python mirror of the same
surface — against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import C1_POINTS, ROOT
from test_oracle_golden import stl_bytes

pytestmark = pytest.mark.gpu


def test_headless_cpp_pipeline(oracle, meshes, tmp_path):
    exe = os.path.join(ROOT, "examples", "headless_main")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True)
    stl = tmp_path / "cubic.stl"
    stl.write_bytes(stl_bytes(meshes["cubic"]))
    r = subprocess.run([exe, str(stl), "0.005", "10", "0.5", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[ACS 3D] 15 Result has been written" in r.stdout and "[headless] stitched path" in r.stdout
    tok = open(tmp_path / "graph.in").read().split()
    assert tok[:2] == ["6", "15"]
    lens = [float(t) for t in tok[2:]]
    # same searches on the oracle (seed 0 = the facade default, keyed Philox, total order)
    G = oracle.Grid.from_triangles(meshes["cubic"], 0.005, 10, oracle.VOX_AABB)
    A = oracle.Acs(G, seed=0)
    pts = [l.split() for l in open(tmp_path / "weld_points.in").read().splitlines()[1:]]
    pts = [tuple(np.float32(v) for v in p) for p in pts]
    k = 0
    for i in range(6):
        for j in range(i + 1, 6):
            assert A.set_points(pts[i], pts[j])[0]
            A.begin(0.5); A.iterate(150)
            assert "%.3f" % A.best()[2] == "%.3f" % lens[k], (i, j)
            A.reset()
            k += 1
    # the grid dump reloads to the same grid through the reference's text format
    H = oracle.Grid.read_file(str(tmp_path / "grid_map.in"))
    assert H.dims == G.dims and np.array_equal(H.isfree(), G.isfree())


def test_python_facade_search_all_pairs(oracle, meshes, tmp_path, capsys):
    import welding_robot_b200 as wr
    pts = C1_POINTS[:4]
    with open(tmp_path / "points.in", "w") as f:
        f.write("%d\n" % len(pts) + "".join("%.6f %.6f %.6f\n" % p for p in pts))
    s = wr.ACS_Rank(seed=3)
    s.creatGridMap(meshes["cubic"], 0.005, 10, str(tmp_path / "grid.in"))
    s.searchBestPathOfPoints(0.5, str(tmp_path / "points.in"), str(tmp_path / "graph.in"))
    G = oracle.Grid.from_triangles(meshes["cubic"], 0.005, 10, oracle.VOX_AABB)
    A = oracle.Acs(G, seed=3)
    for i in range(4):
        for j in range(i + 1, 4):
            A.set_points(pts[i], pts[j]); A.begin(0.5); A.iterate(150)
            ids, dirs, L = A.best(); A.reset()
            ag = s.best_matrix[i][j]
            assert ag is s.best_matrix[j][i] and np.array_equal(ag.ids, ids) and np.float32(ag.L) == np.float32(L)
            assert ag.findPathNode(int(ids[1])) and ag.nodeIndex() == [int(d) for d in dirs]
    route = wr.ACS_GTSP(seed=1)
    route.readFromGraphFile(str(tmp_path / "graph.in"))
    route.computeSolution()
    route.read_all_segments(s.best_matrix)
    assert route.path_segment_nums() == 3 and len(route.g_path_x) == sum(len(s.best_matrix[r][c].ids) for r, c in route.best_path[:-1])
    # grid file written by the facade reloads to the same grid
    g2 = wr.GridMap(); g2.readGridMap(str(tmp_path / "grid.in"))
    assert np.array_equal(g2.isfree(), s.isfree()) and np.allclose(g2.coords()[1], s.coords()[1], atol=2e-6)
    out = capsys.readouterr().out
    assert "[Grid Map] Done!" in out and "route points have been checked" in out
