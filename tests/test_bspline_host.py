"""CPU: the host half of the trajectory smoothing — BS_Basic::SetParam (core/BSplineBasic.h:70-76: knots, constrained control points,
middle points) as wr_bspline_eval computes it before any kernel runs — against the fixtures of the unmodified reference header and
the oracle.  (With m = 0 sample times the entry point makes no CUDA call, so this runs without a GPU; the curve points themselves
are tests/test_gpu_bspline.py.)"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN


def test_setparam_host_side_matches_reference_fixture_and_oracle(oracle):
    import welding_robot_b200 as wr
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    fx = np.load(os.path.join(GOLDEN, "ref_bspline.npz"))
    for i, case in enumerate(MG.BSPLINE_CASES):
        init, fin, mid, tf, u, pre = MG.bspline_inputs(case)
        c = wr.BS_Basic(mid.shape[0], case[0], case[1], case[2])
        assert c.SetParam(init, fin, mid, tf)
        assert np.array_equal(c.Knots_.view(np.uint32), fx["knots%d" % i].view(np.uint32)), case
        assert np.array_equal(c.CPoints_.view(np.uint32), fx["cps%d" % i].view(np.uint32)), case
        _, _, knots, cps = oracle.bspline(case[0], case[1], case[2], init, fin, mid, tf, np.zeros(0, np.float32))
        assert np.array_equal(c.Knots_.view(np.uint32), knots.view(np.uint32)) and np.array_equal(c.CPoints_.view(np.uint32), cps.view(np.uint32))


def test_setparam_rejects_invalid_setups():
    import welding_robot_b200 as wr
    with pytest.raises(wr.WrError):     # NumKnots < 2 * (DEGREE + 1), BSplineBasic.h:54-56
        wr.BS_Basic(1, 3, 0, 0).SetParam(np.zeros(3, np.float32), np.ones(3, np.float32), np.zeros((1, 3), np.float32), 1.0)
    with pytest.raises(wr.WrError):     # a constraint level above the degree
        wr.BS_Basic(8, 1, 2, 2).SetParam(np.zeros(9, np.float32), np.ones(9, np.float32), np.zeros((8, 3), np.float32), 1.0)
    with pytest.raises(wr.WrError):     # degree out of range
        wr.BS_Basic(8, 6, 0, 0).SetParam(np.zeros(3, np.float32), np.ones(3, np.float32), np.zeros((8, 3), np.float32), 1.0)
