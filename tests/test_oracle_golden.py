"""CPU: the oracle restatement against the committed fixtures that the UNMODIFIED reference produced
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import hashlib

import numpy as np
import pytest

from conftest import C1_NAN_PAIR, C1_POINTS


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32_10."""
    L = oracle.lib()
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        L.wro_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(v) for v in o) == want


def stl_bytes(tris):
    """Serialise a (T,12) triangle list as a binary STL (80 B header, u32 count, 50 B records)."""
    out = bytearray(b"golden fixture".ljust(80, b"\0"))
    out += np.uint32(len(tris)).tobytes()
    for t in tris:
        out += t.astype(np.float32).tobytes() + b"\0\0"
    return bytes(out)


def test_stl_parse_matches_reference_reader(oracle, meshes):
    for name in ("cubic", "test", "simplified_piece"):
        t = oracle.stl_parse(stl_bytes(meshes[name]))
        assert np.array_equal(t.view(np.uint32), meshes[name].view(np.uint32))
    with pytest.raises(ValueError):
        oracle.stl_parse(b"solid ascii".ljust(84, b" "))


@pytest.mark.parametrize("mode", [0, 1])
def test_voxel_grid_kats(oracle, meshes, kat, mode):
    for k in kat["grids"]:
        if mode == 0 and k["mesh"] != "cubic" and np.prod(k["dims"]) > 60000:
            continue  # brute force is O(T*N); the AABB-restricted form covers the big ones
        G = oracle.Grid.from_triangles(meshes[k["mesh"]], k["precision"], k["wall"], mode)
        assert list(G.dims) == k["dims"]
        free = G.isfree()
        assert int((free == 0).sum()) == k["occupied"] and sha(free) == k["sha256_isfree"]
        xs, ys, zs = G.coords()
        assert (sha(xs), sha(ys), sha(zs)) == (k["sha256_xs"], k["sha256_ys"], k["sha256_zs"])


def test_brute_and_restricted_voxelisers_agree(oracle, meshes):
    a = oracle.Grid.from_triangles(meshes["simplified_piece"], 0.02, 3, oracle.VOX_BRUTE)
    b = oracle.Grid.from_triangles(meshes["simplified_piece"], 0.02, 3, oracle.VOX_AABB)
    assert np.array_equal(a.isfree(), b.isfree()) and a.tests() > b.tests()


@pytest.fixture(scope="module")
def cubic_grid(oracle, meshes):
    return oracle.Grid.from_triangles(meshes["cubic"], 0.005, 10, oracle.VOX_AABB)


def test_select_step_kats(oracle, cubic_grid, kat):
    """Single selectNext calls of the unmodified reference, incl. NaN-plane and boundary cases."""
    A = oracle.Acs(cubic_grid)
    dead = 0
    for c in kat["steps"]:
        more, infos, d, nxt, L = A.select_step(c["cur"], c["goal"], c["tabu"], c["r31"])
        assert (more, d, nxt) == (c["more"], c["dir"], c["next"]), c
        assert int(np.float32(L).view(np.int32)) == c["L_bits"]
        assert [int(v) for v in infos.view(np.int32)] == c["infos_bits"], c   # bit-exact, NaN payloads included
        dead += d < 0
    assert dead > 0


def test_c1_searches_match_reference(oracle, cubic_grid, ref_acs, kat):
    """The 15 pair searches (150 iterations, adaptive colony) + the NaN pair, under the sequential
    Philox stream and std::sort — exactly what the unmodified reference computed."""
    A = oracle.Acs(cubic_grid, rng_mode=oracle.RNG_SEQUENTIAL, sort_mode=oracle.SORT_STD, seed=kat["seed"])
    for i in range(6):
        for j in range(i + 1, 6):
            if (i + j) % 3 == 2:
                continue   # 10 of the 15 pairs keep the CPU suite short; the GPU box runs all of them
            meta = ref_acs["pair_%d_%d_meta" % (i, j)]
            ok, s, e = A.set_points(C1_POINTS[i], C1_POINTS[j])
            assert ok and (s, e) == (int(meta[0]), int(meta[1]))
            A.begin(0.5); A.iterate(150)
            ids, dirs, L = A.best()
            A.reset()
            assert int(np.float32(L).view(np.int32)) == int(meta[2])
            assert np.array_equal(ids, ref_acs["pair_%d_%d_ids" % (i, j)]) and np.array_equal(dirs, ref_acs["pair_%d_%d_dirs" % (i, j)])
    meta = ref_acs["nan_pair_meta"]
    ok, s, e = A.set_points(*C1_NAN_PAIR)
    assert (int(ok), s, e) == (int(meta[0]), int(meta[1]), int(meta[2]))
    A.begin(0.5); A.iterate(20)
    assert np.isinf(A.best()[2]) and np.isinf(np.int32(meta[3]).view(np.float32))
    c = A.counters()
    assert c["finite_fallthrough"] == 0   # the J.back() UB of ACSRank_3D.hpp:174 never decided anything


def test_c1_pheromone_snapshots(oracle, cubic_grid, kat):
    A = oracle.Acs(cubic_grid, rng_mode=oracle.RNG_SEQUENTIAL, sort_mode=oracle.SORT_STD, seed=kat["seed"])
    A.reset()   # the fixture was taken after earlier pairs: reset() (:307-315) also lifts out-of-bounds slots from 0 to tau0
    for iters, want in kat["c1_tau_snapshots_pair_0_5"].items():
        A.set_points(C1_POINTS[0], C1_POINTS[5])
        A.begin(0.5); A.iterate(int(iters))
        ids, dirs, L = A.best()
        assert sha(A.pheromone()) == want["sha256_tau"]
        assert int(np.float32(L).view(np.int32)) == want["L_bits"] and len(ids) == want["nodes"] and sha(ids.astype(np.int32)) == want["sha256_ids"]
        A.reset()


def test_snap_rule_scan_equals_separable(oracle, cubic_grid):
    A = oracle.Acs(cubic_grid)
    rng = np.random.default_rng(3)
    xs, ys, zs = cubic_grid.coords()
    for _ in range(25):
        p = [rng.uniform(v.min() - 0.01, v.max() + 0.01) for v in (xs, ys, zs)]
        q = [rng.uniform(v.min(), v.max()) for v in (xs, ys, zs)]
        assert A.set_points(p, q) == A.set_points(p, q, scan=True)


def test_gtsp_matches_reference(oracle, kat):
    for g in kat["gtsp"]:
        n = g["n"]
        P = np.random.default_rng(g["points_seed"]).random((n, 3))
        D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        D = np.array([[float("%.6f" % v) for v in row] for row in D])
        T = oracle.Gtsp(D, seed=g["seed"], rng_mode=oracle.RNG_SEQUENTIAL)
        ran = T.iterate(g["iters"], early_stop=True)
        tour, L = T.best()
        assert ran == g["ran"] and L == g["L"] and tour.ravel().tolist() == g["tour"]
        assert T.tau0() == g["tau0"] and sha(T.pheromone()) == g["sha256_pheromone"]


def test_keyed_mode_is_order_independent_and_deterministic(oracle, meshes):
    """The product mode (Philox keyed by (iteration, ant, step), total order): two runs agree, and
    a step cap only turns long walks into dead ants."""
    G = oracle.Grid.from_triangles(meshes["simplified_piece"], 0.02, 3, oracle.VOX_AABB)
    ids = np.flatnonzero(G.isfree())
    runs = []
    for cap in (0, 0, 60):
        A = oracle.Acs(G, seed=5, fixed_colony=64, step_cap=cap)
        A.set_endpoints(int(ids[3]), int(ids[-3]))
        A.begin(1.0); A.iterate(3)
        runs.append((A.pheromone(), A.best(), A.counters()))
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1][2] == runs[1][1][2]
    assert runs[2][2]["dead_step_cap"] > 0 and runs[0][2]["dead_step_cap"] == 0


def test_bspline_matches_reference_fixture(oracle):
    """BS_Basic<float, 3, D, CI, CF> (core/BSplineBasic.h): knots, control points, curve points and return values of the
    restatement against what the unmodified header produced (tests/golden/ref_bspline.npz), bit for bit incl. the NaN time,
    the out-of-range times (clamped) and fin_time itself (the SP_IS_EQUAL branch of _findSpan)."""
    import os
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    fx = np.load(os.path.join(GOLDEN, "ref_bspline.npz"))
    for i, case in enumerate(MG.BSPLINE_CASES):
        init, fin, mid, tf, u, pre = MG.bspline_inputs(case)
        pts, ok, knots, cps = oracle.bspline(case[0], case[1], case[2], init, fin, mid, tf, u, out=pre)
        assert np.array_equal(knots.view(np.uint32), fx["knots%d" % i].view(np.uint32)), case
        assert np.array_equal(cps.view(np.uint32), fx["cps%d" % i].view(np.uint32)), case
        assert np.array_equal(ok, fx["ok%d" % i]), case
        assert np.array_equal(pts.view(np.uint32), fx["pts%d" % i].view(np.uint32)), case


def test_bspline_oracle_properties(oracle):
    """Properties of the restated BS_Basic that need no fixture: a degree-0 curve returns its control points (start point twice, then
    the path, BSplineBasic.h:381-384 + :441-447); a clamped curve starts at the first and ends at the last point; times outside
    [0, fin_time] are clamped (:88-92); equal control points give that point back (partition of unity up to rounding)."""
    rng = np.random.default_rng(12)
    path = np.cumsum(rng.random((40, 3)), axis=0).astype(np.float32)
    pts, ok, knots, cps = oracle.bspline(0, 0, 0, path[0], path[-1], path, 150.0, np.zeros(0, np.float32))
    mids = ((knots[:-1] + knots[1:]) * np.float32(0.5)).astype(np.float32)
    pts, ok, _, _ = oracle.bspline(0, 0, 0, path[0], path[-1], path, 150.0, mids)
    assert ok.all() and np.array_equal(pts, cps) and np.array_equal(cps[0], path[0]) and np.array_equal(cps[1:-1], path)
    for d in (1, 2, 3, 5):
        pts, ok, knots, cps = oracle.bspline(d, 0, 0, path[0], path[-1], path, 10.0, np.array([-3.0, 0.0, 10.0, 99.0], np.float32))
        assert ok.all()
        # (to rounding: the reference's basis functions give N_0(0) = K * (1 / K), one ulp short of 1)
        assert np.array_equal(pts[0], pts[1]) and np.allclose(pts[1], path[0], rtol=1e-6, atol=0)      # before 0 -> clamped to the start
        assert np.array_equal(pts[2], pts[3]) and np.allclose(pts[3], path[-1], rtol=1e-6, atol=0)     # after fin_time -> clamped to the end
        same = np.tile(np.float32([0.25, -1.5, 3.0]), (20, 1))
        q, ok2, _, _ = oracle.bspline(d, 0, 0, same[0], same[0], same, 7.0, np.linspace(0, 7, 57).astype(np.float32))
        assert ok2.all() and np.allclose(q, same[0], rtol=0, atol=2e-6)
