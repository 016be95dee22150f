"""TEST INFRASTRUCTURE (oracle/) — ctypes bindings for the CPU oracle.

Two libraries live here:
  * ``liboracle.so``     – the CPU restatement of the reference hot path (wr_oracle.cpp);
  * ``_ref/libwrref.so`` – the UNMODIFIED reference headers behind ref_harness.cpp
                           (only buildable where /root/reference exists).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libwrref.so")


def build(quiet=True):
    """Compile liboracle.so and, when /root/reference is present, _ref/libwrref.so."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


class AcsParams(C.Structure):
    _fields_ = [("alpha", C.c_int), ("beta", C.c_float), ("rho", C.c_float), ("tau0", C.c_float),
                ("fixed_colony", C.c_int), ("step_cap", C.c_int), ("K", C.c_int), ("seed", C.c_uint64),
                ("rng_mode", C.c_int), ("sort_mode", C.c_int)]


RNG_KEYED, RNG_SEQUENTIAL = 0, 1
SORT_TOTAL, SORT_STD = 0, 1
VOX_BRUTE, VOX_AABB = 0, 1

_lib = None
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE):
            build()
        L = C.CDLL(_ORACLE)
        vp = C.c_void_p
        L.wro_grid_from_triangles.restype = vp
        L.wro_grid_from_triangles.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_int]
        L.wro_grid_from_occupancy.restype = vp
        L.wro_grid_from_occupancy.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_float]
        L.wro_grid_destroy.argtypes = [vp]
        L.wro_grid_dims.argtypes = [vp, vp]
        L.wro_grid_precision.restype = C.c_float
        L.wro_grid_precision.argtypes = [vp]
        L.wro_grid_isfree.argtypes = [vp, vp]
        L.wro_grid_coords.argtypes = [vp, vp, vp, vp]
        L.wro_grid_tests.restype = C.c_uint64
        L.wro_grid_tests.argtypes = [vp]
        L.wro_grid_write_file.argtypes = [vp, C.c_char_p, C.c_int]
        L.wro_grid_read_file.restype = vp
        L.wro_grid_read_file.argtypes = [C.c_char_p]
        L.wro_stl_parse.argtypes = [vp, C.c_size_t, vp, C.c_int]
        L.wro_acs_default_params.argtypes = [C.POINTER(AcsParams)]
        L.wro_acs_create.restype = vp
        L.wro_acs_create.argtypes = [vp, C.POINTER(AcsParams)]
        L.wro_acs_destroy.argtypes = [vp]
        L.wro_acs_set_points.argtypes = [vp, vp, vp, vp]
        L.wro_acs_set_points_scan.argtypes = [vp, vp, vp, vp]
        L.wro_acs_set_endpoints.argtypes = [vp, C.c_int64, C.c_int64]
        L.wro_acs_begin.argtypes = [vp, C.c_float]
        L.wro_acs_iterate.argtypes = [vp, C.c_int]
        L.wro_acs_set_next_search.argtypes = [vp, C.c_uint32]
        L.wro_acs_seq_seek.argtypes = [vp, C.c_uint64]
        L.wro_acs_seq_tell.restype = C.c_uint64
        L.wro_acs_seq_tell.argtypes = [vp]
        L.wro_acs_reset.argtypes = [vp]
        L.wro_acs_best.argtypes = [vp, vp, vp, C.c_int, vp]
        L.wro_acs_pheromone.argtypes = [vp, vp]
        L.wro_acs_set_pheromone.argtypes = [vp, vp]
        L.wro_acs_last_colony.argtypes = [vp, vp, vp, vp]
        L.wro_acs_last_ant.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
        L.wro_acs_counters.argtypes = [vp, vp]
        L.wro_acs_phase_seconds.argtypes = [vp, vp]
        L.wro_acs_select_step.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_int, C.c_uint32, vp, vp, vp, vp]
        L.wro_gtsp_create.restype = vp
        L.wro_gtsp_create.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]
        L.wro_gtsp_destroy.argtypes = [vp]
        L.wro_gtsp_iterate.argtypes = [vp, C.c_int, C.c_int]
        L.wro_gtsp_best.argtypes = [vp, vp, vp]
        L.wro_gtsp_pheromone.argtypes = [vp, vp]
        L.wro_gtsp_tau0.restype = C.c_double
        L.wro_gtsp_tau0.argtypes = [vp]
        L.wro_gtsp_steps.restype = C.c_uint64
        L.wro_gtsp_steps.argtypes = [vp]
        L.wro_philox.argtypes = [vp, vp, vp]
        L.wro_bspline.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_float, vp, C.c_int, vp, vp, vp, vp]
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(_REF)
        vp = C.c_void_p
        L.wrref_stl_read.argtypes = [C.c_char_p, vp, C.c_int]
        L.wrref_create.restype = vp
        L.wrref_destroy.argtypes = [vp]
        L.wrref_voxelize.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, vp]
        L.wrref_voxelize_to_file.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, C.c_char_p, vp]
        L.wrref_read_grid_file.argtypes = [vp, C.c_char_p, vp]
        L.wrref_grid_read.argtypes = [vp, vp, vp, vp, vp]
        L.wrref_grid_set_free.argtypes = [vp, vp]
        L.wrref_acs_init.argtypes = [vp]
        L.wrref_acs_set_points.argtypes = [vp, vp, vp, vp]
        L.wrref_acs_set_endpoints.argtypes = [vp, C.c_int64, C.c_int64]
        L.wrref_acs_compute.argtypes = [vp, C.c_float, C.c_int, C.c_uint64, vp]
        L.wrref_acs_best.argtypes = [vp, vp, vp, C.c_int, vp]
        L.wrref_acs_pheromone.argtypes = [vp, vp]
        L.wrref_acs_reset.argtypes = [vp]
        L.wrref_acs_select_step.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_int, C.c_int, vp, vp, vp, vp]
        L.wrref_acs_search_all.argtypes = [vp, vp, C.c_int, C.c_float, C.c_uint64, C.c_char_p, C.c_char_p, vp]
        L.wrref_acs_pair_best.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
        L.wrref_gtsp_create.restype = vp
        L.wrref_gtsp_create.argtypes = [vp, C.c_int, C.c_char_p]
        L.wrref_gtsp_run.argtypes = [vp, C.c_int, C.c_uint64, vp]
        L.wrref_gtsp_best.argtypes = [vp, vp, vp]
        L.wrref_gtsp_pheromone.argtypes = [vp, vp]
        L.wrref_gtsp_tau0.restype = C.c_double
        L.wrref_gtsp_tau0.argtypes = [vp]
        L.wrref_bspline.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_float, vp, C.c_int, vp, vp, vp, vp]
        _ref = L
    return _ref


# --------------------------------------------------------------------------------------
# thin object wrappers
# --------------------------------------------------------------------------------------
def stl_parse(data: bytes):
    """read_STL.hpp:131-156 — returns (T,12) float32: normal, v0, v1, v2."""
    buf = np.frombuffer(data, np.uint8)
    n = lib().wro_stl_parse(_p(buf), len(data), None, 0)
    if n < 0:
        raise ValueError("STL parse error %d" % n)
    t = np.zeros((n, 12), np.float32)
    lib().wro_stl_parse(_p(buf), len(data), _p(t), n)
    return t


class Grid:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle grid creation failed")
        self.h = C.c_void_p(handle)
        d = (C.c_int * 3)()
        lib().wro_grid_dims(self.h, d)
        self.dims = (d[0], d[1], d[2])  # rx, ry, rz
        self.precision = lib().wro_grid_precision(self.h)

    @classmethod
    def from_triangles(cls, tris, precision, wall, mode=VOX_AABB):
        tris = np.ascontiguousarray(tris, np.float32)
        return cls(lib().wro_grid_from_triangles(_p(tris), len(tris), precision, wall, mode))

    @classmethod
    def from_occupancy(cls, isfree, xs, ys, zs, precision):
        isfree = np.ascontiguousarray(isfree, np.uint8)
        xs, ys, zs = (np.ascontiguousarray(v, np.float32) for v in (xs, ys, zs))
        assert isfree.size == len(xs) * len(ys) * len(zs)
        return cls(lib().wro_grid_from_occupancy(_p(isfree), len(xs), len(ys), len(zs), _p(xs), _p(ys), _p(zs), precision))

    @classmethod
    def read_file(cls, path):
        return cls(lib().wro_grid_read_file(path.encode()))

    def write_file(self, path, compat=False):
        return lib().wro_grid_write_file(self.h, path.encode(), int(compat))

    @property
    def n(self):
        return self.dims[0] * self.dims[1] * self.dims[2]

    def isfree(self):
        out = np.zeros(self.n, np.uint8)
        lib().wro_grid_isfree(self.h, _p(out))
        return out

    def coords(self):
        xs, ys, zs = (np.zeros(d, np.float32) for d in self.dims)
        lib().wro_grid_coords(self.h, _p(xs), _p(ys), _p(zs))
        return xs, ys, zs

    def tests(self):
        return lib().wro_grid_tests(self.h)

    def __del__(self):
        try:
            lib().wro_grid_destroy(self.h)
        except Exception:
            pass


def default_params(**kw):
    p = AcsParams()
    lib().wro_acs_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Acs:
    def __init__(self, grid: Grid, params: AcsParams = None, **kw):
        self.grid = grid
        self.params = params or default_params(**kw)
        self.K = self.params.K
        self.h = C.c_void_p(lib().wro_acs_create(grid.h, C.byref(self.params)))
        if not self.h:
            raise RuntimeError("oracle acs creation failed")

    def set_points(self, s, e, scan=False):
        s = np.asarray(s, np.float32); e = np.asarray(e, np.float32)
        ids = np.zeros(2, np.int64)
        f = lib().wro_acs_set_points_scan if scan else lib().wro_acs_set_points
        ok = f(self.h, _p(s), _p(e), _p(ids))
        return bool(ok), int(ids[0]), int(ids[1])

    def set_endpoints(self, s, e):
        return bool(lib().wro_acs_set_endpoints(self.h, s, e))

    def begin(self, predict, seq_pos=0):
        """seq_pos: position of the sequential stream (rng_mode SEQUENTIAL only); None = continue."""
        keep = lib().wro_acs_seq_tell(self.h)
        lib().wro_acs_begin(self.h, predict)
        lib().wro_acs_seq_seek(self.h, keep if seq_pos is None else seq_pos)

    def set_next_search(self, idx):
        """Keyed stream: the search index the next begin() takes (default: the number of begin() calls so far)."""
        lib().wro_acs_set_next_search(self.h, idx)

    def iterate(self, n):
        r = lib().wro_acs_iterate(self.h, n)
        if r != 0:
            raise RuntimeError("oracle iterate failed")

    def reset(self):
        lib().wro_acs_reset(self.h)

    def best(self, cap=1 << 20):
        ids = np.zeros(cap, np.int64); dirs = np.zeros(cap, np.int32); L = C.c_float()
        n = lib().wro_acs_best(self.h, _p(ids), _p(dirs), cap, C.byref(L))
        return ids[:n].copy(), dirs[:max(n - 1, 0)].copy(), L.value

    def pheromone(self):
        out = np.zeros(self.grid.n * self.K, np.float32)
        lib().wro_acs_pheromone(self.h, _p(out))
        return out

    def set_pheromone(self, tau):
        tau = np.ascontiguousarray(tau, np.float32)
        assert tau.size == self.grid.n * self.K
        lib().wro_acs_set_pheromone(self.h, _p(tau))

    def last_colony(self):
        c = C.c_int(); lam = C.c_float(); q = C.c_float()
        lib().wro_acs_last_colony(self.h, C.byref(c), C.byref(lam), C.byref(q))
        return c.value, lam.value, q.value

    def last_ant(self, k, cap=1 << 16):
        ids = np.zeros(cap, np.int64); dirs = np.zeros(cap, np.int32); L = C.c_float(); order = C.c_int()
        n = lib().wro_acs_last_ant(self.h, k, _p(ids), _p(dirs), cap, C.byref(L), C.byref(order))
        if n > cap:
            return self.last_ant(k, n)
        return ids[:n].copy(), dirs[:max(n - 1, 0)].copy(), L.value, order.value

    def counters(self):
        out = np.zeros(9, np.uint64)
        lib().wro_acs_counters(self.h, _p(out))
        keys = ["ant_steps", "ants", "arrived", "dead_no_candidate", "dead_fallthrough", "dead_step_cap",
                "finite_fallthrough", "iterations", "rng_draws"]
        return dict(zip(keys, (int(v) for v in out)))

    def phase_seconds(self):
        out = np.zeros(3, np.float64)
        lib().wro_acs_phase_seconds(self.h, _p(out))
        return dict(walk=out[0], evaporate=out[1], sort_deposit=out[2])

    def select_step(self, cur, goal, tabu, r31):
        tabu = np.asarray(tabu, np.int64)
        infos = np.zeros(self.K, np.float32); d = C.c_int(); nxt = C.c_int64(); L = C.c_float()
        more = lib().wro_acs_select_step(self.h, cur, goal, _p(tabu), len(tabu), r31, _p(infos), C.byref(d), C.byref(nxt), C.byref(L))
        return more, infos, d.value, nxt.value, L.value

    def __del__(self):
        try:
            lib().wro_acs_destroy(self.h)
        except Exception:
            pass


class Gtsp:
    def __init__(self, dis, seed=0, colony_id=0, rng_mode=RNG_KEYED, cnt=None):
        dis = np.ascontiguousarray(dis, np.float64)
        self.n = dis.shape[0]
        cnt = self.n * (self.n - 1) // 2 if cnt is None else cnt
        self.h = C.c_void_p(lib().wro_gtsp_create(_p(dis), self.n, cnt, seed, colony_id, rng_mode))

    def iterate(self, iters, early_stop=False):
        return lib().wro_gtsp_iterate(self.h, iters, int(early_stop))

    def best(self):
        tour = np.zeros(2 * (self.n + 1), np.int32); L = C.c_double()
        m = lib().wro_gtsp_best(self.h, _p(tour), C.byref(L))
        return tour[:2 * m].reshape(m, 2).copy(), L.value

    def pheromone(self):
        out = np.zeros((self.n, self.n), np.float64)
        lib().wro_gtsp_pheromone(self.h, _p(out))
        return out

    def tau0(self):
        return lib().wro_gtsp_tau0(self.h)

    def steps(self):
        return lib().wro_gtsp_steps(self.h)

    def __del__(self):
        try:
            lib().wro_gtsp_destroy(self.h)
        except Exception:
            pass


# --------------------------------------------------------------------------------------
# the unmodified reference (only where oracle/_ref/libwrref.so exists)
# --------------------------------------------------------------------------------------
class Ref:
    """Drives the UNMODIFIED reference headers (ref_harness.cpp)."""

    def __init__(self):
        self.L = ref()
        self.h = C.c_void_p(self.L.wrref_create())
        self.dims = None

    @staticmethod
    def stl_read(path, cap=40000):
        t = np.zeros((cap, 12), np.float32)
        n = ref().wrref_stl_read(path.encode(), _p(t), cap)
        return t[:n].copy()

    def voxelize(self, tris, precision, wall, file=None):
        tris = np.ascontiguousarray(tris, np.float32)
        d = (C.c_int * 3)()
        if file:
            n = self.L.wrref_voxelize_to_file(self.h, _p(tris), len(tris), precision, wall, file.encode(), d)
        else:
            n = self.L.wrref_voxelize(self.h, _p(tris), len(tris), precision, wall, d)
        self.dims = (d[0], d[1], d[2])
        return n

    def read_grid_file(self, path):
        d = (C.c_int * 3)()
        n = self.L.wrref_read_grid_file(self.h, path.encode(), d)
        self.dims = (d[0], d[1], d[2])
        return n

    def grid(self):
        n = self.dims[0] * self.dims[1] * self.dims[2]
        free = np.zeros(n, np.uint8)
        xs, ys, zs = (np.zeros(d, np.float32) for d in self.dims)
        self.L.wrref_grid_read(self.h, _p(free), _p(xs), _p(ys), _p(zs))
        return free, xs, ys, zs

    def set_free(self, isfree):
        isfree = np.ascontiguousarray(isfree, np.uint8)
        self.L.wrref_grid_set_free(self.h, _p(isfree))

    def acs_init(self):
        self.L.wrref_acs_init(self.h)

    def set_points(self, s, e):
        s = np.asarray(s, np.float32); e = np.asarray(e, np.float32)
        ids = np.zeros(2, np.int64)
        ok = self.L.wrref_acs_set_points(self.h, _p(s), _p(e), _p(ids))
        return bool(ok), int(ids[0]), int(ids[1])

    def set_endpoints(self, start_id, goal_id):
        return bool(self.L.wrref_acs_set_endpoints(self.h, start_id, goal_id))

    def compute(self, predict, max_iter, seed):
        calls = C.c_uint64()
        self.L.wrref_acs_compute(self.h, predict, max_iter, seed, C.byref(calls))
        return calls.value

    def best(self, cap=1 << 20):
        ids = np.zeros(cap, np.int64); dirs = np.zeros(cap, np.int32); L = C.c_float()
        n = self.L.wrref_acs_best(self.h, _p(ids), _p(dirs), cap, C.byref(L))
        return ids[:n].copy(), dirs[:max(n - 1, 0)].copy(), L.value

    def pheromone(self):
        n = self.dims[0] * self.dims[1] * self.dims[2]
        out = np.zeros(n * 6, np.float32)
        self.L.wrref_acs_pheromone(self.h, _p(out))
        return out

    def reset(self):
        self.L.wrref_acs_reset(self.h)

    def select_step(self, cur, goal, tabu, r31):
        tabu = np.asarray(tabu, np.int64)
        infos = np.zeros(6, np.float32); d = C.c_int(); nxt = C.c_int64(); L = C.c_float()
        more = self.L.wrref_acs_select_step(self.h, cur, goal, _p(tabu), len(tabu), r31, _p(infos), C.byref(d), C.byref(nxt), C.byref(L))
        return more, infos, d.value, nxt.value, L.value

    def search_all(self, pts, predict, seed, tmpdir):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        lens = np.zeros((n, n), np.float32)
        cnt = self.L.wrref_acs_search_all(self.h, _p(pts), n, predict, seed, os.path.join(tmpdir, "points.in").encode(),
                                          os.path.join(tmpdir, "graph.in").encode(), _p(lens))
        return cnt, lens

    def pair_best(self, i, j, cap=1 << 20):
        ids = np.zeros(cap, np.int64); L = C.c_float()
        n = self.L.wrref_acs_pair_best(self.h, i, j, _p(ids), cap, C.byref(L))
        return ids[:n].copy(), L.value


class RefGtsp:
    def __init__(self, dis, tmp_graph):
        dis = np.ascontiguousarray(dis, np.float64)
        self.n = dis.shape[0]
        self.L = ref()
        self.h = C.c_void_p(self.L.wrref_gtsp_create(_p(dis), self.n, tmp_graph.encode()))

    def run(self, iters, seed):
        calls = C.c_uint64()
        ran = self.L.wrref_gtsp_run(self.h, iters, seed, C.byref(calls))
        return ran, calls.value

    def best(self):
        tour = np.zeros(2 * (self.n + 1), np.int32); L = C.c_double()
        m = self.L.wrref_gtsp_best(self.h, _p(tour), C.byref(L))
        return tour[:2 * m].reshape(m, 2).copy(), L.value

    def pheromone(self):
        out = np.zeros((self.n, self.n), np.float64)
        self.L.wrref_gtsp_pheromone(self.h, _p(out))
        return out

    def tau0(self):
        return self.L.wrref_gtsp_tau0(self.h)


def bspline(degree, ci, cf, init, fin, middle, fin_time, times, out=None, use_ref=False):
    """BS_Basic<float, 3, degree, ci, cf>(len(middle)).SetParam(init, fin, middle, fin_time) + getCurvePoint at `times`
    (BSplineBasic.h:70-111): the restatement, or — use_ref — the unmodified header behind oracle/_ref.
    Returns (points [m][3], ok [m], knots, control points [ncp][3])."""
    init = np.ascontiguousarray(init, np.float32).ravel(); fin = np.ascontiguousarray(fin, np.float32).ravel()
    middle = np.ascontiguousarray(middle, np.float32)
    n = middle.shape[0] if middle.size else 0
    stride = middle.shape[1] if n else 3
    u = np.ascontiguousarray(times, np.float32).ravel()
    out = np.zeros((u.size, 3), np.float32) if out is None else np.ascontiguousarray(out, np.float32).copy()
    ok = np.zeros(u.size, np.uint8)
    knots = np.zeros(degree + n + 2 + ci + cf + 1, np.float32)
    cps = np.zeros((n + 2 + ci + cf, 3), np.float32)
    fn = ref().wrref_bspline if use_ref else lib().wro_bspline
    rc = fn(degree, ci, cf, _p(init), _p(fin), _p(middle), n, stride, fin_time, _p(u), u.size, _p(out), _p(ok), _p(knots), _p(cps))
    if rc != 0:
        raise ValueError("bspline: unsupported (degree, ci, cf) = (%d, %d, %d)" % (degree, ci, cf))
    return out, ok.astype(bool), knots, cps
