"""GPU: concurrent searches (wr_acs_search_batch) against the same searches run one after the other (wr_acs_search_pairs)
and against the oracle — best path, its slots and its length of every query, bit for bit.  Covers the reference's all-pairs
workload (C1: adaptive colonies, the pair whose ants all die on the NaN plane), independent queries on an obstacle grid
(C5 shape), non-default parameters, the park/resume path of ants whose shared-memory visited table fills up, and the
fall-back of a chunk whose pheromone table overflows."""
import numpy as np
import pytest

from conftest import C1_NAN_PAIR, C1_POINTS
from test_gpu_parity import make_pair, synthetic_boxes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wr():
    import welding_robot_b200 as wr
    return wr


def same(res_a, res_b):
    assert len(res_a) == len(res_b)
    for q, ((i1, d1, L1), (i2, d2, L2)) in enumerate(zip(res_a, res_b)):
        assert (np.isinf(L1) and np.isinf(L2)) or np.float32(L1) == np.float32(L2), (q, L1, L2)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2), q


def test_c1_all_pairs_batch_equals_sequential_and_oracle(wr, oracle, meshes):
    """cubic.stl @ (0.005, 10), the reference's defaults: 15 pairs + the NaN-plane pair, 150 iterations, adaptive colonies."""
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED)
    _, g2 = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0x5EED)
    pts = list(C1_POINTS) + list(C1_NAN_PAIR)
    node = g.snapPoints(pts)
    assert (node >= 0).all()
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)] + [(6, 7)]
    s = [int(node[i]) for i, _ in pairs]; e = [int(node[j]) for _, j in pairs]
    seq = g.searchPairs(s, e, 0.5, iterations=150)
    bat = g2.searchBatch(s, e, 0.5, iterations=150)
    same(seq, bat)
    st = g2.batchStats()
    assert st["fallbacks"] == 0 and st["entries_used"] > 0, st
    assert np.isinf(bat[15][2]) and len(bat[15][0]) == 0 and all(np.isfinite(r[2]) for r in bat[:15])
    assert np.array_equal(g.pheromone().view(np.uint32), g2.pheromone().view(np.uint32))   # both end in the reset() state
    for key in ("ant_steps", "ants", "arrived", "dead_no_candidate", "dead_fallthrough", "dead_step_cap"):
        assert g.counters()[key] == g2.counters()[key], key
    # the oracle, query by query (search index = position in the batch)
    for q in (0, 7, 14, 15):
        A.set_endpoints(s[q], e[q]); A.set_next_search(q)
        A.begin(0.5); A.iterate(150)
        ids, dirs, L = A.best()
        assert (np.isinf(L) and np.isinf(bat[q][2])) or (np.float32(L) == np.float32(bat[q][2]) and np.array_equal(ids, bat[q][0]) and np.array_equal(dirs, bat[q][1])), q
        A.reset()
    # a second batch on the same handle continues the search indices, like a second sequential call
    seq2 = g.searchPairs(s[:4], e[:4], 0.5, iterations=40)
    bat2 = g2.searchBatch(s[:4], e[:4], 0.5, iterations=40)
    same(seq2, bat2)


@pytest.mark.parametrize("params,env", [
    (dict(fixed_colony=200, step_cap=300), {}),
    (dict(fixed_colony=96, step_cap=300, alpha=2, beta=0.9, rho=0.7, tau0=0.5), {}),
    (dict(fixed_colony=150, step_cap=400), {"WR_BATCH_TABLE": "16"}),       # 16-entry visited tables: most ants park and resume from HBM tables
    (dict(fixed_colony=150, step_cap=300), {"WR_BATCH_LOG2": "10"}),        # 1024-entry pheromone table: the chunk falls back to the sequential loop
    (dict(fixed_colony=64, step_cap=300), {"WR_BATCH_MEM_MB": "4", "WR_BATCH_LOG2": "18"}),   # memory for a few queries per chunk only
])
def test_fixed_colony_queries(wr, oracle, meshes, params, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=17, **params)
    _, g2 = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=17, **params)
    free = np.flatnonzero(A.grid.isfree())
    rng = np.random.default_rng(2)
    s = [int(v) for v in free[rng.integers(0, len(free), 24)]]
    e = [int(v) for v in free[rng.integers(0, len(free), 24)]]
    s[5] = e[5] = int(free[100])            # a query whose start is its goal: no ant takes a step
    seq = g.searchPairs(s, e, 1.0, iterations=12)
    bat = g2.searchBatch(s, e, 1.0, iterations=12)
    same(seq, bat)
    st = g2.batchStats()
    assert (st["fallbacks"] > 0) == (env.get("WR_BATCH_LOG2") == "10"), st
    if "WR_BATCH_TABLE" in env:
        assert g2.counters()["table_overflows"] > 0
    if "WR_BATCH_MEM_MB" in env:
        assert st["queries_per_chunk"] < 24, st
    assert g.counters()["ant_steps"] == g2.counters()["ant_steps"]
    for q in (0, 11, 23):
        A.set_endpoints(s[q], e[q]); A.set_next_search(q)
        A.begin(1.0); A.iterate(12)
        ids, dirs, L = A.best()
        assert (np.isinf(L) and np.isinf(bat[q][2])) or (np.float32(L) == np.float32(bat[q][2]) and np.array_equal(ids, bat[q][0])), q
        A.reset()


def test_queries_on_an_obstacle_grid(wr):
    """C5 shape at 1/64 of the volume: 128^3 boxes grid, coordinates = indices, 96 queries x 256 ants."""
    n = 128
    free = synthetic_boxes(n, 60, 4)
    axis = np.arange(n, dtype=np.float32)
    ids = np.flatnonzero(free)
    rng = np.random.default_rng(5)
    s = [int(v) for v in ids[rng.integers(0, len(ids), 96)]]
    e = [int(v) for v in ids[rng.integers(0, len(ids), 96)]]
    res = []
    for batch in (False, True):
        g = wr.ACS_Rank(seed=5, fixed_colony=256, step_cap=1024)
        g.creatFromOccupancy(free, axis, axis, axis, 1.0)
        g.initFromGridMap()
        res.append(g.searchPairs(s, e, 100.0, iterations=10, batch=batch))
        if batch:
            assert g.batchStats()["fallbacks"] == 0
        del g
    same(res[0], res[1])
    assert sum(np.isfinite(r[2]) for r in res[1]) > 48


def test_batch_on_a_used_handle_runs_sequentially(wr, oracle, meshes):
    """A handle whose pheromone field is not in its initial state: the first search of the sequential loop would start from
    that field, so the batch call takes the sequential path — same answers."""
    _, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=4, fixed_colony=128, step_cap=300)
    _, g2 = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=4, fixed_colony=128, step_cap=300)
    free = np.flatnonzero(g.isfree())
    s = [int(free[10]), int(free[500])]; e = [int(free[-10]), int(free[-700])]
    for h in (g, g2):
        h.setEndpoints(s[0], e[1]); h.begin(1.0); h.iterate(3)     # no reset(): the field carries deposits
    same(g.searchPairs(s, e, 1.0, iterations=8), g2.searchBatch(s, e, 1.0, iterations=8))
    assert g2.batchStats()["queries_per_chunk"] == 0               # the batch buffers were never needed
