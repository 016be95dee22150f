"""GPU: trajectory smoothing (wr_bspline_eval / BS_Basic, core/BSplineBasic.h:33-120 as main.cpp:287-352 uses it) against the CPU
oracle and the fixtures of the unmodified reference header — bit-exact: knots, control points, curve points, return values."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def bits(a):
    """Bit pattern of a float array with every NaN mapped to one value: a NaN time yields NaN points on both sides, but x86 propagates
    the input's payload (0x7FC00000) where the GPU returns its canonical NaN (0x7FFFFFFF)."""
    if a.dtype != np.float32:
        return a
    b = np.ascontiguousarray(a).view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def run_gpu(case, init, fin, mid, tf, u, pre):
    import welding_robot_b200 as wr
    c = wr.BS_Basic(mid.shape[0], case[0], case[1], case[2])
    c.SetParam(init, fin, mid, tf)
    pts, ok = c.getCurvePoints(u, out=pre.copy())
    return pts, ok, c.Knots_, c.CPoints_


def test_bspline_equals_reference_fixture_and_oracle(oracle):
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    fx = np.load(os.path.join(GOLDEN, "ref_bspline.npz"))
    for i, case in enumerate(MG.BSPLINE_CASES):
        init, fin, mid, tf, u, pre = MG.bspline_inputs(case)
        pts, ok, knots, cps = run_gpu(case, init, fin, mid, tf, u, pre)
        assert np.array_equal(bits(knots), bits(fx["knots%d" % i])), case
        assert np.array_equal(bits(cps), bits(fx["cps%d" % i])), case
        assert np.array_equal(ok, fx["ok%d" % i]), case
        assert np.array_equal(bits(pts), bits(fx["pts%d" % i])), case


@pytest.mark.parametrize("degree,ci,cf", [(0, 0, 0), (2, 2, 2), (3, 2, 2), (5, 2, 2), (4, 1, 0), (1, 1, 1)])
def test_bspline_random_setups_equal_oracle(oracle, degree, ci, cf):
    rng = np.random.default_rng(100 * degree + 10 * ci + cf)
    for n, tf, m in [(max(1, degree + 1 - ci - cf), 1.0, 100), (218, 150.0, 4096), (5000, 6000.0, 200000)]:
        mid = rng.random((n, 3), dtype=np.float32) * 4 - 2
        init = np.concatenate([mid[0], rng.random(3 * ci) - 0.5]).astype(np.float32)
        fin = np.concatenate([mid[-1], rng.random(3 * cf) - 0.5]).astype(np.float32)
        u = (rng.random(m, dtype=np.float32) * 1.2 - 0.1) * np.float32(tf)
        u[:3] = (0.0, tf, np.nan)
        pre = rng.random((m, 3), dtype=np.float32)
        want = oracle.bspline(degree, ci, cf, init, fin, mid, tf, u, out=pre)
        got = run_gpu((degree, ci, cf), init, fin, mid, tf, u, pre)
        for g, w in zip(got, want):
            assert np.array_equal(bits(np.ascontiguousarray(g)), bits(np.ascontiguousarray(w))), (degree, ci, cf, n)


def test_bspline_demo_pipeline_properties():
    """The demo's two curves (main.cpp:299-352) at the size of a stitched path, through properties that need no oracle: a degree-0
    curve returns its control points (the start point twice, then the path); both curves start at the first and end at the last
    path point; a million sample times in one launch."""
    import welding_robot_b200 as wr
    rng = np.random.default_rng(9)
    path = np.cumsum(rng.random((218, 3), dtype=np.float32) * 0.01, axis=0).astype(np.float32)
    c0 = wr.BS_Basic(len(path), 0, 0, 0)
    c0.SetParam(path[0], path[-1], path, 150.0)
    k = c0.Knots_
    mids = ((k[:-1] + k[1:]) * np.float32(0.5)).astype(np.float32)            # one time inside every knot interval
    pts, ok = c0.getCurvePoints(mids)
    assert ok.all() and np.array_equal(bits(pts), bits(c0.CPoints_[:len(mids)]))
    first = c0.getCurvePoints(np.arange(10, 161, 10, dtype=np.float32))[0]
    c2 = wr.BS_Basic(len(first), 2, 2, 2)
    s2 = np.concatenate([path[0], np.zeros(6, np.float32)]); e2 = np.concatenate([path[-1], np.zeros(6, np.float32)])
    c2.SetParam(s2, e2, first, 6000.0)
    t = np.linspace(0, 6000, 1 << 20).astype(np.float32)
    sm, ok2 = c2.getCurvePoints(t)
    assert ok2.all() and np.isfinite(sm).all()
    assert np.array_equal(bits(sm[0]), bits(path[0])) and np.array_equal(bits(sm[-1]), bits(path[-1]))
    lo, hi = np.minimum(path.min(0), 0) - 1e-3, path.max(0) + 1e-3
    assert (sm >= lo).all() and (sm <= hi).all()                               # convex hull of the control points (zero end derivatives)


def test_bspline_rejects_invalid_setups():
    import welding_robot_b200 as wr
    c = wr.BS_Basic(1, 3, 0, 0)                                               # NumKnots < 2 * (DEGREE + 1), BSplineBasic.h:54-56
    with pytest.raises(wr.WrError):
        c.SetParam(np.zeros(3, np.float32), np.ones(3, np.float32), np.zeros((1, 3), np.float32), 1.0)
    c = wr.BS_Basic(8, 1, 2, 2)                                               # a constraint level above the degree
    with pytest.raises(wr.WrError):
        c.SetParam(np.zeros(9, np.float32), np.ones(9, np.float32), np.zeros((8, 3), np.float32), 1.0)
