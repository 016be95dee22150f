"""GPU: batched seam-ordering colonies (K4) against the CPU oracle — tours bit-exact, lengths and
pheromone matrices bit-exact (the spec asks 1e-5; the FP64 addition order is reproduced)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def matrix(n, seed):
    P = np.random.default_rng(seed).random((n, 3))
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
    return np.array([[float("%.6f" % v) for v in row] for row in D])   # as a graph file would carry it


@pytest.mark.parametrize("n,iters,batch", [(8, 12, 3), (33, 6, 4), (64, 5, 3), (200, 4, 2), (256, 10, 3)])
def test_gtsp_matches_oracle(oracle, n, iters, batch):
    import welding_robot_b200 as wr
    D = matrix(n, n)
    g = wr.ACS_GTSP(seed=77)
    g.dis, g.city_num, g.cnt = D, n, n * (n - 1) // 2
    res = g.computeBatch(batch, iters, colony_first=5)
    for b in range(batch):
        T = oracle.Gtsp(D, seed=77, colony_id=5 + b)
        assert T.iterate(iters) == iters
        tour, L = T.best()
        assert g.tau0() == T.tau0()
        assert np.array_equal(res[b][0], tour), (n, b)
        assert res[b][1] == L
        assert np.array_equal(g.pheromone(b).view(np.uint64), T.pheromone().view(np.uint64))
    assert len({tuple(r[0].ravel()) for r in res}) > 1 or n <= 8   # colonies really are independent streams


def test_gtsp_facade_early_stop(oracle, tmp_path, capsys):
    """computeSolution's stagnation rule (ACS_GTSP.hpp:261-276) through the graph-file interface."""
    import welding_robot_b200 as wr
    n = 12
    D = matrix(n, 3)
    f = tmp_path / "graph.in"
    with open(f, "w") as fp:
        fp.write("%d %d\n" % (n, n * (n - 1) // 2))
        for i in range(n):
            for j in range(i + 1, n):
                fp.write("%.6f\n" % D[i, j])
    g = wr.ACS_GTSP(seed=4)
    assert g.readFromGraphFile(str(f)) and g.computeSolution()
    T = oracle.Gtsp(D, seed=4, colony_id=0)
    ran = T.iterate(n * n, early_stop=True)
    tour, L = T.best()
    assert np.array_equal(g.best_path, tour) and g.best_L == L
    assert ("iteration %d:Best so far" % (ran - 1)) in capsys.readouterr().out
    assert g.path_segment_nums() == n - 1
