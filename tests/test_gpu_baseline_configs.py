"""GPU parity on BASELINE.json's own configurations and on the parameter space the defaults never reach.

  * C2 exactly as bench.py builds it (256^3, 4096 ants, seed 1, step cap 8192): iterations 1 and 2 against the oracle —
    every ant's visited-node sequence and rank, lambda, Q, the best path and the whole 100 M-slot pheromone field;
  * a 65 536-ant colony (C3's colony size) through the ranking path for colonies beyond one CTA;
  * alpha != 1 (power<T>, ACSRank_3D.hpp:48-60) and non-default beta / rho / tau0;
  * grids the packed-coordinate walk cannot address are rejected, not mis-walked;
  * the per-search Philox index: successive searches on one handle are independent and reproducible.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_parity import compare_iteration, make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wr():
    import welding_robot_b200 as wr
    return wr


@pytest.mark.timeout(900)
def test_c2_bench_workload_two_iterations_bit_exact(wr, oracle):
    """The headline configuration itself, not a scaled-down stand-in."""
    sys.path.insert(0, ROOT)
    import bench
    bench.select_workload("C2")
    wl = bench.build_workload_gpu()                  # K1 voxeliser -> natural grid -> 256^3 lattice
    wl_cpu = bench.build_workload_cpu()              # the oracle's voxeliser, same embedding
    assert np.array_equal(wl["isfree"], wl_cpu["isfree"]) and wl["start"] == wl_cpu["start"] and wl["goal"] == wl_cpu["goal"]
    for ax in ("xs", "ys", "zs"):
        assert np.array_equal(wl[ax], wl_cpu[ax])
    G = oracle.Grid.from_occupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
    A = oracle.Acs(G, seed=bench.SEED, fixed_colony=bench.ANTS_PER_GPU, step_cap=bench.STEP_CAP)
    g = wr.ACS_Rank(seed=bench.SEED, fixed_colony=bench.ANTS_PER_GPU, step_cap=bench.STEP_CAP, update_mode=4)
    g.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], bench.PRECISION)
    g.initFromGridMap()
    A.set_endpoints(wl["start"], wl["goal"]); g.setEndpoints(wl["start"], wl["goal"])
    A.begin(bench.PREDICT); g.begin(bench.PREDICT)
    for _ in range(2):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
    oc, gc = A.counters(), g.counters()
    for key in ("ant_steps", "ants", "arrived", "dead_no_candidate", "dead_fallthrough", "dead_step_cap", "iterations"):
        assert oc[key] == gc[key], key
    assert gc["ants"] == 2 * 4096 and gc["arrived"] > 0


@pytest.mark.timeout(900)
def test_colony_of_65536_ants(wr, oracle, meshes):
    """C3's colony size on a small grid: ranking beyond the single-CTA kernel, 13 108 eligible ranks, the record path and
    (forced on the second handle) rank sets with rows of 410 words."""
    for policy in ("2", "1"):
        os.environ["WR_RANKSET_POLICY"] = policy
        try:
            A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=77, fixed_colony=65536, step_cap=128, update_mode=4)
        finally:
            del os.environ["WR_RANKSET_POLICY"]
        ids = np.flatnonzero(A.grid.isfree())
        s, e = int(ids[40]), int(ids[len(ids) // 3])
        A.set_endpoints(s, e); g.setEndpoints(s, e)
        A.begin(1.0); g.begin(1.0)
        for it in range(2):
            A.iterate(1); g.iterate(1)
            oc, olam, oq = A.last_colony(); gc, glam, gq = g.lastColony()
            assert (oc, np.float32(olam).tobytes(), np.float32(oq).tobytes()) == (gc, np.float32(glam).tobytes(), np.float32(gq).tobytes())
            for k in list(range(0, 65536, 97)) + [65535]:
                oid, odir, oL, oorder = A.last_ant(k)
                gid, gdir, gL, gorder = g.lastAnt(k)
                assert gorder == oorder, (k, gorder, oorder)
                assert (np.isinf(oL) and np.isinf(gL)) or (np.float32(oL) == np.float32(gL) and np.array_equal(oid, gid) and np.array_equal(odir, gdir)), k
            ob, gb = A.best(), g.bestPath()
            assert np.float32(ob[2]) == np.float32(gb[2]) and np.array_equal(ob[0], gb[0])
            assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32)), "pheromone field differs (policy %s, iteration %d)" % (policy, it)
        c = g.counters()
        assert c["ants"] == 2 * 65536 and c["arrived"] > 1000 and c["ant_steps"] == A.counters()["ant_steps"]
        assert (g.updateStats()["rankset_iterations"] > 0) == (policy == "1")
        del g


@pytest.mark.parametrize("alpha,beta,rho,tau0", [(2, 0.9, 0.7, 0.5), (3, 0.3, 0.93, 2.5), (0, 0.6, 0.8, 1.0)])
@pytest.mark.parametrize("update_mode", [0, 4])
def test_non_default_parameters_bit_exact(wr, oracle, meshes, alpha, beta, rho, tau0, update_mode):
    """power<T>(tau, alpha) with alpha != 1 (square-and-multiply in the reference's order; alpha = 0 gives the constant 1),
    and evaporation / initial pheromone / heuristic weight away from the literals of :319-324."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=13, fixed_colony=700, step_cap=300,
                     alpha=alpha, beta=beta, rho=rho, tau0=tau0, update_mode=update_mode)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[25]), int(ids[-40])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    A.begin(1.0); g.begin(1.0)
    for it in range(5):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g, check_tau=it in (0, 4))
    A.iterate(15); g.iterate(15)
    compare_iteration(A, g)
    A.reset(); g.reset()
    assert np.array_equal(A.pheromone().view(np.uint32), g.pheromone().view(np.uint32))
    assert g.counters()["arrived"] > 0


def test_adaptive_colony_with_non_default_parameters(wr, oracle, meshes):
    """reference colony rule (:247) together with alpha = 2 on the C1 grid"""
    from conftest import C1_POINTS
    A, g = make_pair(wr, oracle, meshes["cubic"], 0.005, 10, seed=0xBEEF, alpha=2, beta=1.1, rho=0.85, tau0=0.7)
    ok, s, e = A.set_points(C1_POINTS[1], C1_POINTS[4])
    assert g.setPoints(C1_POINTS[1], C1_POINTS[4]) and ok
    A.begin(0.5); g.begin(0.5)
    for it in range(8):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g, check_tau=it in (0, 7))
    A.iterate(60); g.iterate(60)
    compare_iteration(A, g)


def test_axis_longer_than_1024_is_rejected(wr):
    from welding_robot_b200 import _lib
    g = wr.ACS_Rank()
    g.creatFromOccupancy(np.ones(1025 * 2 * 2, np.uint8), np.arange(1025), np.arange(2), np.arange(2), 1.0)
    with pytest.raises(_lib.WrError) as ei:
        g.initFromGridMap()
    assert ei.value.status == -1 and "1024" in str(ei.value)
    g = wr.ACS_Rank()
    g.creatFromOccupancy(np.ones(1024 * 2 * 2, np.uint8), np.arange(1024), np.arange(2), np.arange(2), 1.0)
    g.initFromGridMap()
    g.setEndpoints(0, 1024 * 4 - 1)
    g.begin(10.0); g.iterate(1)
    assert g.counters()["ants"] > 0


def test_successive_searches_draw_independent_streams(wr, oracle, meshes):
    """The reference draws every search from one continuous rand() stream (ACSRank_3D.hpp:169, :327).  Here the search index
    is part of the Philox counter: the same endpoints searched twice on one handle give different colonies, the oracle
    follows search by search, and setNextSearch(0) reproduces the first search exactly."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=3, fixed_colony=256, step_cap=300)
    ids = np.flatnonzero(A.grid.isfree())
    s, e = int(ids[25]), int(ids[-40])
    A.set_endpoints(s, e); g.setEndpoints(s, e)
    firsts = []
    for search in range(3):
        A.begin(1.0); g.begin(1.0)
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g)
        firsts.append([g.lastAnt(k)[0].tobytes() for k in range(0, 256, 16)])
        A.reset(); g.reset()
    assert firsts[0] != firsts[1] and firsts[1] != firsts[2]
    A.set_next_search(0); g.setNextSearch(0)
    A.begin(1.0); g.begin(1.0)
    A.iterate(1); g.iterate(1)
    compare_iteration(A, g)
    assert [g.lastAnt(k)[0].tobytes() for k in range(0, 256, 16)] == firsts[0]
    # search indices beyond 16 bits use the upper half of the block counter
    A.reset(); g.reset()
    A.set_next_search(0x12345); g.setNextSearch(0x12345)
    A.begin(1.0); g.begin(1.0)
    A.iterate(2); g.iterate(2)
    compare_iteration(A, g)


def test_upload_of_negative_zero(wr, oracle, meshes):
    """-0.0f is the clean-tile sentinel's bit pattern; an uploaded -0.0f must behave like the value 0."""
    A, g = make_pair(wr, oracle, meshes["simplified_piece"], 0.02, 3, seed=5, fixed_colony=200, step_cap=250)
    ids = np.flatnonzero(A.grid.isfree())
    A.set_endpoints(int(ids[30]), int(ids[len(ids) // 3])); g.setEndpoints(int(ids[30]), int(ids[len(ids) // 3]))
    rng = np.random.default_rng(9)
    tau = (0.25 + rng.random(A.pheromone().size)).astype(np.float32)
    tau[rng.integers(0, tau.size, 5000)] = -0.0
    A.set_pheromone(np.where(tau == 0, np.float32(0.0), tau)); g.setPheromone(tau)
    A.begin(1.0); g.begin(1.0)
    for _ in range(3):
        A.iterate(1); g.iterate(1)
        compare_iteration(A, g, check_tau=False)
        assert np.array_equal(A.pheromone(), g.pheromone())     # values (0 == -0)
