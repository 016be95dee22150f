"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, and against the committed reference fixtures.  Bit-exact for grids, visited-node
sequences and — in the rank-ordered update modes — the pheromone field."""
import hashlib

import numpy as np
import pytest

from conftest import C1_NAN_PAIR, C1_POINTS

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def wr():
    import welding_robot_b200 as wr
    return wr


def gpu_grid(wr, tris, precision, wall, cls=None, **kw):
    g = (cls or wr.GridMap)(**kw)
    g.creatGridMap(tris, precision, wall)
    return g


# ---------------------------------------------------------------------------------------------
# K1 voxeliser
# ---------------------------------------------------------------------------------------------
def test_voxel_grid_matches_reference_fixtures(wr, meshes, kat):
    for k in kat["grids"]:
        g = gpu_grid(wr, meshes[k["mesh"]], k["precision"], k["wall"])
        assert [g.rangeX, g.rangeY, g.rangeZ] == k["dims"]
        free = g.isfree()
        assert int((free == 0).sum()) == k["occupied"] == g.stats()["occupied"]
        assert sha(free) == k["sha256_isfree"], k
        xs, ys, zs = g.coords()
        assert (sha(xs), sha(ys), sha(zs)) == (k["sha256_xs"], k["sha256_ys"], k["sha256_zs"])


@pytest.mark.parametrize("mesh,precision,wall", [("simplified_piece", 0.0061, 4), ("origin_piece", 0.0047, 6), ("cubic", 0.0123, 1),
                                                 ("simplified_piece", 0.0034981, 10)])
def test_voxel_grid_matches_oracle(wr, oracle, meshes, mesh, precision, wall):
    G = oracle.Grid.from_triangles(meshes[mesh], precision, wall, oracle.VOX_AABB)
    g = gpu_grid(wr, meshes[mesh], precision, wall)
    assert (g.rangeX, g.rangeY, g.rangeZ) == G.dims
    assert np.array_equal(g.isfree(), G.isfree())
    assert g.stats()["tests"] == G.tests()   # same set of (triangle, node) pairs passes the box test
    bits = g.bits()
    assert np.array_equal(np.unpackbits(bits.view(np.uint8), bitorder="little")[:g.size_of_map()], 1 - G.isfree())


def test_occupancy_entry_point_round_trip(wr):
    rng = np.random.default_rng(5)
    rx, ry, rz = 37, 21, 13
    free = (rng.random(rx * ry * rz) > 0.3).astype(np.uint8)
    g = wr.GridMap().creatFromOccupancy(free, np.arange(rx), np.arange(ry), np.arange(rz), 1.0)
    assert np.array_equal(g.isfree(), free)


# ---------------------------------------------------------------------------------------------
# K2 walk + ranking + K3 update against the oracle (keyed Philox, total order)
# ---------------------------------------------------------------------------------------------
def make_pair(wr, oracle, tris, precision, wall, **params):
    G = oracle.Grid.from_triangles(tris, precision, wall, oracle.VOX_AABB)
    op = {k: v for k, v in params.items() if k in ("alpha", "beta", "rho", "tau0", "fixed_colony", "step_cap", "seed", "K")}
    A = oracle.Acs(G, **op)
    g = wr.ACS_Rank(**params)
    g.creatGridMap(tris, precision, wall)
    g.initFromGridMap()
    return A, g


def compare_iteration(A, g, check_tau=True, exact_tau=True):
    oc, olam, oq = A.last_colony()
    gc, glam, gq = g.lastColony()
    assert (oc, np.float32(olam).tobytes(), np.float32(oq).tobytes()) == (gc, np.float32(glam).tobytes(), np.float32(gq).tobytes())
    for k in range(oc):
        oid, odir, oL, oorder = A.last_ant(k)
        gid, gdir, gL, gorder = g.lastAnt(k)
        assert gorder == oorder, (k, gorder, oorder)
        if np.isinf(oL):
            assert np.isinf(gL), k
        else:
            assert np.float32(oL).tobytes() == np.float32(gL).tobytes(), (k, oL, gL)
            assert np.array_equal(oid, gid), k          # visited-node sequence, bit-exact
            assert np.array_equal(odir, gdir), k
    ob, gb = A.best(), g.bestPath()
    assert np.float32(ob[2]).tobytes() == np.float32(gb[2]).tobytes()
    if not np.isinf(ob[2]):
        assert np.array_equal(ob[0], gb[0]) and np.array_equal(ob[1], gb[1])
    if check_tau:
        to, tg = A.pheromone(), g.pheromone()
        if exact_tau:
            assert np.array_equal(to.view(np.uint32), tg.view(np.uint32)), "pheromone field differs: max |d| = %g" % np.abs(to - tg).max()
        else:
            assert np.allclose(to, tg, rtol=1e-5, atol=0)


@pytest.mark.parametrize("update_mode", [0, 1, 4])
def test_c1_adaptive_colony_bit_exact(wr, oracle, meshes, update_mode):
    """cubic.stl @ (0.005, 10), reference defaults (adaptive colony), pair (0,5): every ant of
    every iteration, the best path and the whole pheromone field, bit for bit."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED, update_mode=update_mode)
    ok, s, e = A.set_points(C1_POINTS[0], C1_POINTS[5])
    assert g.setPoints(C1_POINTS[0], C1_POINTS[5]) and ok
    assert (g._start_id, g._goal_id) == (s, e)
    A.begin(0.5); g.begin(0.5)
    for it in range(12):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g, check_tau=it in (0, 1, 9, 11))
    A.iterate(40); g.iterate(40)
    compare_iteration(A, g)
    oc, gc = A.counters(), g.counters()
    for key in ("ant_steps", "ants", "arrived", "dead_no_candidate", "dead_fallthrough", "dead_step_cap", "iterations"):
        assert oc[key] == gc[key], key
    assert gc["table_overflows"] == 0


def test_c1_all_pairs_lengths_and_nan_plane(wr, oracle, meshes):
    """searchBestPathOfPoints' pair loop (ACSRank_3D.hpp:472-499) incl. reset() between pairs, and
    the pair beyond the duplicate-coordinate plane: NaN must propagate and every ant must die."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED)
    for (i, j) in [(0, 1), (2, 5), (1, 4)]:
        A.set_points(C1_POINTS[i], C1_POINTS[j]); assert g.setPoints(C1_POINTS[i], C1_POINTS[j])
        A.begin(0.5); A.iterate(150); g.computeSolution(0.5)
        compare_iteration(A, g)
        A.reset(); g.reset()
        assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))
    ok, s, e = A.set_points(*C1_NAN_PAIR)
    assert g.setPoints(*C1_NAN_PAIR) and (g._start_id, g._goal_id) == (s, e)
    A.begin(0.5); A.iterate(5); g.begin(0.5); g.iterate(5)
    compare_iteration(A, g)
    assert np.isinf(g.bestPath()[2]) and g.counters()["dead_fallthrough"] > 0


@pytest.mark.parametrize("table_log2", [4, 9])
def test_fixed_colony_step_cap_and_table_overflow(wr, oracle, meshes, table_log2):
    """Fixed colony of 512 ants, step cap 300 (some ants hit it), and — with a 16-slot visited table —
    the exact re-run path (pass 2 with tables in HBM)."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=7, fixed_colony=512, step_cap=300, walk_table_log2=table_log2)
    G = A.grid
    free = G.isfree()
    ids = np.flatnonzero(free)
    s, e = int(ids[10]), int(ids[-10])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    for it in range(6):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    gc = g.counters()
    assert (gc["table_overflows"] > 0) == (table_log2 == 4)
    assert gc["dead_step_cap"] == A.counters()["dead_step_cap"]


# ---------------------------------------------------------------------------------------------
# K = 26: the neighbourhood the reference scaffolds and disables (ACSRank_3D.hpp:367-385); the oracle's K = 26 mode is
# the specification (same enumeration, lengths precision*1.414f / precision*1.732f, every slot evaporated)
# ---------------------------------------------------------------------------------------------
def slot26_offsets():
    out = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if (dx, dy, dz) != (0, 0, 0):
                    out.append((dx, dy, dz))
    return out


@pytest.mark.parametrize("update_mode", [0, 1, 4])
def test_k26_adaptive_colony_bit_exact(wr, oracle, meshes, update_mode):
    """cubic.stl @ (0.005, 10), adaptive colony, K = 26: every ant's visited-node sequence, length (float sums of three
    step lengths in path order), rank, the best path and the whole 26-slot pheromone field, bit for bit."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED, update_mode=update_mode, K=26)
    ok, s, e = A.set_points(C1_POINTS[0], C1_POINTS[5])
    assert g.setPoints(C1_POINTS[0], C1_POINTS[5]) and ok
    A.begin(0.5); g.begin(0.5)
    assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))   # initFromGridMap: tau0 / 0 on out-of-bounds slots
    for it in range(10):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g, check_tau=it in (0, 1, 9))
    A.iterate(30); g.iterate(30)
    compare_iteration(A, g)
    oc, gc = A.counters(), g.counters()
    for key in ("ant_steps", "ants", "arrived", "dead_no_candidate", "dead_fallthrough", "dead_step_cap", "iterations"):
        assert oc[key] == gc[key], key
    assert gc["arrived"] > 0
    # the moves really are 26-connected: some diagonal slot was taken, and ids follow the slot enumeration
    ids, dirs, L = g.bestPath()
    rx, ry = g.rangeX, g.rangeY
    off = slot26_offsets()
    assert np.array_equal(np.diff(ids), np.array([off[d][0] + off[d][1] * rx + off[d][2] * rx * ry for d in dirs]))
    assert any(sum(1 for c in off[d] if c) > 1 for d in dirs)


@pytest.mark.parametrize("table_log2", [4, 9])
def test_k26_fixed_colony_step_cap_table_overflow_and_nan(wr, oracle, meshes, table_log2):
    """K = 26 with a fixed colony, a step cap that some ants hit and — with a 16-slot visited table — the park/resume path
    (pass 2 with tables in HBM, L carried in the parked state); then the NaN-plane pair (every ant dies)."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=11, fixed_colony=384, step_cap=200, walk_table_log2=table_log2, K=26)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[10]), int(ids[-10])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    for it in range(5):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    gc = g.counters()
    assert (gc["table_overflows"] > 0) == (table_log2 == 4)
    assert gc["dead_step_cap"] == A.counters()["dead_step_cap"]
    A.reset(); g.reset()
    assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))


def test_k26_nan_plane(wr, oracle, meshes):
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED, K=26)
    ok, s, e = A.set_points(*C1_NAN_PAIR)
    assert g.setPoints(*C1_NAN_PAIR) and (g._start_id, g._goal_id) == (s, e)
    A.begin(0.5); A.iterate(4); g.begin(0.5); g.iterate(4)
    compare_iteration(A, g)
    assert A.counters()["dead_fallthrough"] == g.counters()["dead_fallthrough"]


def test_k26_rejected_for_sharded_handles(wr, meshes):
    g = wr.ACS_Rank(K=26)
    g.creatGridMap(meshes["cubic"], 0.0123, 1)
    g.initFromGridMap()
    from welding_robot_b200 import _lib
    assert _lib.lib().wr_acs_set_shard(g._a, 0, 2) == -1   # WR_ERR_INVALID


@pytest.mark.parametrize("K,policy", [(6, "1"), (26, "1"), (6, "0")])
def test_rankset_update_mode_bit_exact(wr, oracle, meshes, K, policy, monkeypatch):
    """WR_UPDATE_RANKSET (per-slot sets of depositing ranks instead of sorted records): same bits as the oracle with a
    fixed colony (hot slots crossed by most eligible ranks take the warp-cooperative chain), across reset(), a second
    search on the same handle, and a second handle that inherits the first one's parked (all-zero) table.  Policy 1 forces
    the rank-set path on every iteration; policy 0 is the shipped device-side switch (thresholds lowered so that both
    paths and both transitions occur within the run)."""
    monkeypatch.setenv("WR_RANKSET_POLICY", policy)
    if policy == "0":
        monkeypatch.setenv("WR_RANKSET_ON", "100000"); monkeypatch.setenv("WR_RANKSET_OFF", "9000")
    for rep in range(2):
        A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=21 + rep, fixed_colony=640, step_cap=300, update_mode=4, K=K)
        ids = np.flatnonzero(A.grid.isfree())
        for (s, e) in [(int(ids[10]), int(ids[-10])), (int(ids[200]), int(ids[-300]))]:
            A.set_endpoints(s, e); g.setEndpoints(s, e)
            A.begin(1.0); g.begin(1.0)
            for it in range(4):
                A.iterate(1); g.iterate(1)
                compare_iteration(A, g)
            A.iterate(12); g.iterate(12)
            compare_iteration(A, g)
            A.reset(); g.reset()
            assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))
        assert g.counters()["deposit_records"] > 0
        st = g.updateStats()
        assert st["rankset_iterations"] > 0, st
        del g


@pytest.mark.parametrize("update_mode", [0, 4])
def test_uploaded_pheromone_field_and_clean_tiles(wr, oracle, meshes, update_mode):
    """wr_acs_upload_pheromone on a clean-tile handle (every tile becomes explicit) and the clean-tile bookkeeping itself:
    a fresh field has no dirty tile, deposits dirty only the tiles they touch, reset() cleans them again — with the
    downloaded field equal to the oracle's at every point."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=5, fixed_colony=300, step_cap=250, update_mode=update_mode)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[30]), int(ids[len(ids) // 3])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    dirty, tiles = g.fieldStats()
    assert dirty == 0 and tiles > 0
    assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))
    A.iterate(3); g.iterate(3)
    compare_iteration(A, g)
    dirty, _ = g.fieldStats()
    assert 0 < dirty <= tiles
    A.reset(); g.reset()
    assert g.fieldStats()[0] == 0
    assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))
    A.begin(1.0); g.begin(1.0)
    A.iterate(2); g.iterate(2)
    compare_iteration(A, g)
    # an explicit field: random values everywhere (out-of-bounds slots included)
    rng = np.random.default_rng(3)
    tau = (0.25 + rng.random(A.pheromone().size)).astype(np.float32)
    A.set_pheromone(tau); g.setPheromone(tau)
    assert g.fieldStats()[0] == tiles
    A.begin(1.0); g.begin(1.0)
    for _ in range(3):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)


def test_rankset_widest_rows(wr, oracle, meshes, monkeypatch):
    """4900 ants: 981 eligible ranks = 31 row words, the widest rank-set row (summary bits 0..30 + the on-best bit 31);
    the rank-set path forced on every iteration, a converging colony so that slots collect ranks from every word."""
    monkeypatch.setenv("WR_RANKSET_POLICY", "1")
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=41, fixed_colony=4900, step_cap=120, update_mode=4)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[60]), int(ids[len(ids) // 4])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    for _ in range(4):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    st = g.updateStats()
    assert st["rankset_iterations"] == 4 and g.counters()["deposit_records"] > 100000


@pytest.mark.parametrize("K", [6, 26])
def test_large_colony_ranking(wr, oracle, meshes, K):
    """20 000 ants: beyond the single-kernel ranking, the colony goes through the multi-kernel radix sort; massive ties in
    the keys (a few hundred distinct step counts), dead ants, every ant's rank compared with the oracle's total order."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=31, fixed_colony=20000, step_cap=160, K=K)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[40]), int(ids[len(ids) // 2])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    for _ in range(2):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    c = g.counters()
    assert c["arrived"] > 0 and c["dead_step_cap"] > 0


def test_atomic_update_within_tolerance(wr, oracle, meshes):
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=9, fixed_colony=256, step_cap=400, update_mode=2)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[5]), int(ids[-5])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    A.iterate(1); g.iterate(1)
    compare_iteration(A, g, exact_tau=False)   # tolerance 1e-5 relative (north_star) for the atomic mode


def test_path_properties_full_size(wr, meshes):
    """BASELINE configs[1] scale (256-long grid, 4096 ants): properties that need no oracle —
    6-connected, self-avoiding, free cells only, starts/ends at the endpoints, L = steps * precision."""
    precision = 0.823812 / 235.5
    g = wr.ACS_Rank(seed=1, fixed_colony=4096, step_cap=4096)
    g.creatGridMap(meshes["simplified_piece"], precision, 10)
    assert max(g.rangeX, g.rangeY, g.rangeZ) == 256
    free = g.isfree()
    rx, ry, rz = g.rangeX, g.rangeY, g.rangeZ
    nid = lambda x, y, z: (z * ry + y) * rx + x  # noqa: E731
    s, e = nid(5, 5, 5), nid(rx - 6, ry - 6, rz - 6)
    assert free[s] and free[e]
    g.setEndpoints(s, e)
    g.begin(1.0); g.iterate(3)
    c = g.counters()
    assert c["ants"] == 3 * 4096 and c["arrived"] > 0
    strides = {0: -rx * ry, 1: -rx, 2: -1, 3: 1, 4: rx, 5: rx * ry}
    checked = 0
    for k in list(range(0, 4096, 97)):
        ids, dirs, L, order = g.lastAnt(k)
        if np.isinf(L):
            continue
        assert ids[0] == s and ids[-1] == e
        assert len(set(ids.tolist())) == len(ids)
        assert free[ids].all()
        assert np.array_equal(np.diff(ids), np.array([strides[d] for d in dirs]))
        x = ids % rx
        assert (np.abs(np.diff(x)) <= 1).all()     # no wrap across a row
        Lacc = np.float32(0)
        for _ in range(len(dirs)):
            Lacc = np.float32(Lacc + np.float32(precision))
        assert np.float32(L) == Lacc
        checked += 1
    assert checked > 5
    ids, dirs, L = g.bestPath()
    assert ids[0] == s and ids[-1] == e and len(ids) == len(dirs) + 1


# ---------------------------------------------------------------------------------------------
# larger grids (BASELINE configs[2] / configs[4] shapes)
# ---------------------------------------------------------------------------------------------
def test_voxel_grid_512_long_matches_oracle(wr, oracle, meshes):
    """C3: origin_piece.stl (29 888 triangles) at a 512-long grid (~20 M nodes).  The reference's
    O(T*N) loop cannot run at this size (4e15 tests); the oracle's box-restricted form of the same
    predicate can."""
    precision = 0.823812 / 491.5
    G = oracle.Grid.from_triangles(meshes["origin_piece"], precision, 10, oracle.VOX_AABB)
    g = gpu_grid(wr, meshes["origin_piece"], precision, 10)
    assert (g.rangeX, g.rangeY, g.rangeZ) == G.dims and max(G.dims) == 512
    assert np.array_equal(g.isfree(), G.isfree())
    st = g.stats()
    assert st["tests"] == G.tests() and st["occupied"] == int((G.isfree() == 0).sum())


def synthetic_boxes(n, nboxes, seed):
    """C5-style obstacle grid: union of axis-aligned boxes, 10-cell free wall, coordinates = indices."""
    rng = np.random.default_rng(seed)
    occ = np.zeros((n, n, n), bool)
    for _ in range(nboxes):
        e = rng.integers(4, 49, 3)
        c = [int(rng.integers(10, n - 10 - int(e[k]))) for k in range(3)]
        occ[c[2]:c[2] + e[2], c[1]:c[1] + e[1], c[0]:c[0] + e[0]] = True
    return (~occ).astype(np.uint8).ravel()


def test_synthetic_512_cubed_query(wr, oracle):
    """C5 shape: one start/goal query on a synthetic 512^3 obstacle grid (3.2 GB pheromone field),
    256 ants, against the oracle for two iterations — every ant, best path and the whole field."""
    n = 512
    free = synthetic_boxes(n, 1500, 4)
    axis = np.arange(n, dtype=np.float32)
    ids = np.flatnonzero(free)
    rng = np.random.default_rng(5)
    s = int(ids[rng.integers(0, len(ids))])
    sz, sy, sx = s // (n * n), (s // n) % n, s % n
    cand = ids[(np.abs(ids // (n * n) - sz) + np.abs((ids // n) % n - sy) + np.abs(ids % n - sx)) == 96]
    e = int(cand[rng.integers(0, len(cand))])
    G = oracle.Grid.from_occupancy(free, axis, axis, axis, 1.0)
    A = oracle.Acs(G, seed=5, fixed_colony=256, step_cap=2048)
    g = wr.ACS_Rank(seed=5, fixed_colony=256, step_cap=2048)
    g.creatFromOccupancy(free, axis, axis, axis, 1.0)
    g.initFromGridMap()
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(100.0); g.begin(100.0)
    for _ in range(2):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    assert g.counters()["arrived"] > 0


# ---------------------------------------------------------------------------------------------
# all-pairs driver on the device
# ---------------------------------------------------------------------------------------------
def test_snap_points_kernel_matches_host_scan_and_oracle(wr, oracle, meshes):
    """setPoints' full-grid scan (ACSRank_3D.hpp:545-562) as a kernel == the host scan == the oracle's scan, for points
    on nodes, between nodes, on the duplicate-coordinate plane, inside the solid and outside the grid."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=1)
    rng = np.random.default_rng(11)
    xs, ys, zs = g.coords()
    pts = [tuple(p) for p in C1_POINTS] + list(C1_NAN_PAIR)
    for _ in range(200):
        pts.append((np.float32(rng.uniform(xs[0] - 0.01, xs[-1] + 0.01)), np.float32(rng.uniform(ys[0] - 0.01, ys[-1] + 0.01)),
                    np.float32(rng.uniform(zs[0] - 0.01, zs[-1] + 0.01))))
    ids = g.snapPoints(pts)
    hits = 0
    for p, i in zip(pts, ids):
        ok = g.setPoints(p, p)
        assert (g._start_id if ok else -1) == int(i)
        ok_o, s, _ = A.set_points(p, p)
        assert (s if ok_o else -1) == int(i)
        hits += int(i) >= 0
    assert 20 < hits < len(pts)


def test_search_pairs_equals_pair_by_pair_loop(wr, oracle, meshes):
    """wr_acs_search_pairs (no host synchronisation between pairs) == begin/iterate/best/reset pair by pair, bit for bit,
    including the pair whose ants all die (NaN plane) and the pheromone field left behind."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED)
    pts = [C1_POINTS[0], C1_POINTS[3], C1_POINTS[5], C1_NAN_PAIR[0], C1_NAN_PAIR[1]]
    node = g.snapPoints(pts)
    assert (node >= 0).all()
    pairs = [(0, 1), (1, 2), (3, 4), (0, 2)]
    loop = []
    for i, j in pairs:
        g.setEndpoints(int(node[i]), int(node[j]))
        g.begin(0.5); g.iterate(30)
        loop.append(g.bestPath())
        g.reset()
    tau_loop = g.pheromone()
    g.setNextSearch(0)   # the pairs of searchPairs take consecutive Philox search indices, like the begin() calls of the loop above
    res = g.searchPairs([node[i] for i, _ in pairs], [node[j] for _, j in pairs], 0.5, iterations=30)
    for (ids, dirs, L), (ids2, dirs2, L2) in zip(loop, res):
        assert np.array_equal(ids, ids2) and np.array_equal(dirs, dirs2)
        assert np.float32(L) == np.float32(L2) or (np.isinf(L) and np.isinf(L2))
    assert np.isinf(res[2][2]) and len(res[2][0]) == 0 and np.isfinite(res[0][2])
    assert np.array_equal(tau_loop.view(np.uint32), g.pheromone().view(np.uint32))
    # and the oracle agrees on the first pair
    A.set_endpoints(int(node[0]), int(node[1])); A.begin(0.5); A.iterate(30)
    ids, dirs, L = A.best()
    assert np.array_equal(ids, res[0][0]) and np.float32(L) == np.float32(res[0][2])
