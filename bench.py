#!/usr/bin/env python
"""bench.py — ant-steps/s of the rank-based 3-D ACS search on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

Workload "C2" (BASELINE.json configs[1]): the simplified_piece mesh voxelised at a
256-long grid (precision 0.823812/235.5, wall 10) and embedded in a 256^3 lattice whose node
coordinates follow the reference's formula (model_grid_map.hpp:204-211) for a cubic bounding box;
4096 ants per GPU, K = 6, Philox seed 1, start/goal at opposite corners inset by 5 cells, step cap
8192.  One "step" = `--iters` ACS iterations (ant construction + ranking + pheromone update) of one
search.  With N GPUs the colony is 4096*N ants sharded by ant index (weak scaling) with one exchange
per iteration.  The pheromone field (403 MB) is larger than L2 (126 MB): every iteration streams it
from HBM, so no L2 flush is needed between timed iterations.

Prints ONE JSON line.  `value` = whole-job ant-steps/s with the grid already in HBM (CUDA events on
the launching stream, max over ranks); `e2e` = the same metric through the public API from HOST
buffers, every step: occupancy upload, handle creation, search, best-path download.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ANTS_PER_GPU = 4096
STEP_CAP = 8192
SEED = 1
PRECISION = 0.823812 / 235.5
WALL = 10
CUBE = 256
MESH = "simplified_piece"
LABEL = "C2: ACSRank_3D, simplified_piece @256-long grid in 256^3"


def select_workload(name):
    """C2 (BASELINE.json configs[1]) is the bench workload; C3 (configs[2]: origin_piece at a 512-long grid in 512^3, 8192 ants per
    GPU, the pheromone field 3.2 GB per GPU) is the multi-GPU scaling case north_star names — an exploration run, never the default."""
    global CUBE, PRECISION, MESH, LABEL, ANTS_PER_GPU, STEP_CAP
    if name == "C3":
        CUBE, PRECISION, MESH = 512, 0.823812 / 491.5, "origin_piece"
        ANTS_PER_GPU, STEP_CAP = 8192, 16384
        LABEL = "C3: ACSRank_3D, origin_piece @512-long grid in 512^3"
PREDICT = 1.0                # with a fixed colony only Q of the pre-arrival iterations depends on it (ACSRank_3D.hpp:249)
WALK_BYTES_PER_STEP = 30     # 4*K tau + K/8 occupancy + 4 id + 1 dir, K = 6
UPDATE_BYTES_PER_SLOT = 8    # 4 read + 4 write per directed slot per iteration


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and clock-event reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in
    process every few ms (a 4 ms step leaves no room for `nvidia-smi -lms`' start-up); where pynvml cannot initialise the
    recipe's `nvidia-smi --query-gpu` loop is the fallback.  Only samples taken between mark_begin() and mark_end() count."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None     # rows: (t, sm_mhz, sm_max_mhz, [reasons])
        self.t0 = self.t1 = None
        self.quit = threading.Event()
        self.ready = threading.Event()
        self.source = None

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        self.source = "nvml"
        while not self.quit.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            self.rows.append((time.perf_counter(), sm, mx, [n for n, b in names if bits & b]))
            self.ready.set()
            time.sleep(0.003)

    def _smi(self):
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            c = [v.strip() for v in line.split(",")]
            try:
                sm, mx = float(c[0]), float(c[1])
            except (ValueError, IndexError):
                continue
            rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]) if v.lower().startswith("active")]
            self.rows.append((time.perf_counter(), sm, mx, rs))
            self.ready.set()
            if self.quit.is_set():
                break

    def run(self):
        try:
            self._nvml()
        except Exception:
            try:
                self._smi()
            except OSError:
                pass
        self.ready.set()

    def mark_begin(self):
        self.ready.wait(10.0)       # first sample is in before the timed region starts
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.quit.set()
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or float("inf"))]
        window = "timed region"
        if not rows:                # never silently empty: say that the window is wider
            rows, window = list(self.rows), "whole run (no sample fell inside the timed region)"
        sm = [r[1] for r in rows]
        mx = [r[2] for r in rows]
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": self.source, "window": window}


# ---------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------
def axis_coords(rng, wall, mn, mx, precision):
    """model_grid_map.hpp:204-205 in float32 (bit-identical to the C expression, see tests)."""
    f = np.float32
    out = np.zeros(rng, np.float32)
    for i in range(rng):
        if i < wall:
            out[i] = f(mn) - f(f(wall - i) * f(precision))
        elif i >= rng - wall:
            out[i] = f(mx) + f(f(i - rng + wall) * f(precision))
        else:
            out[i] = f(mn) + f(f(i - wall) * f(precision))
    return out


def embed(free_zyx, gmin, gmax):
    """Natural voxelisation -> CUBE^3 workload: cubic bounding box (the long axis' extent on all axes),
    the reference's coordinate formula, the piece's occupancy centred, everything else free."""
    rz, ry, rx = free_zyx.shape
    assert max(rx, ry, rz) == CUBE, (rx, ry, rz)
    f = np.float32
    ext = max(f(gmax[k]) - f(gmin[k]) for k in range(3))
    cmin = [f(gmin[k]) for k in range(3)]
    cmax = [f(f(gmin[k]) + f(ext)) for k in range(3)]
    axes = [axis_coords(CUBE, WALL, cmin[k], cmax[k], f(PRECISION)) for k in range(3)]
    cube = np.ones((CUBE, CUBE, CUBE), np.uint8)
    oz, oy, ox = (CUBE - rz) // 2, (CUBE - ry) // 2, (CUBE - rx) // 2
    cube[oz:oz + rz, oy:oy + ry, ox:ox + rx] = free_zyx
    nid = lambda x, y, z: (z * CUBE + y) * CUBE + x  # noqa: E731
    start, goal = nid(5, 5, 5), nid(CUBE - 6, CUBE - 6, CUBE - 6)
    flat = np.ascontiguousarray(cube.ravel())
    assert flat[start] and flat[goal]
    return dict(isfree=flat, xs=axes[0], ys=axes[1], zs=axes[2], start=start, goal=goal, natural=(rx, ry, rz),
                cmin=[float(v) for v in cmin], cmax=[float(v) for v in cmax])


def build_workload_gpu():
    """Natural grid from the product's GPU voxeliser (K1)."""
    import welding_robot_b200 as wr
    tris = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))[MESH]
    g = wr.GridMap()
    with contextlib.redirect_stdout(io.StringIO()):
        g.creatGridMap(tris, PRECISION, WALL)
    free = g.isfree().reshape(g.rangeZ, g.rangeY, g.rangeX)
    mn, mx = g.bbox()
    wl = embed(free, mn, mx)
    wl["vox"] = g.stats(); wl["ntri"] = len(tris)
    return wl


def build_workload_cpu():
    """The same workload built on the CPU (reference arm): the oracle's box-restricted voxeliser gives
    the natural grid (the reference's own O(T*N) loop would need ~140 s for it)."""
    from oracle import oracle as O
    tris = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))[MESH]
    G = O.Grid.from_triangles(tris, PRECISION, WALL, O.VOX_AABB)
    rx, ry, rz = G.dims
    v = tris[:, 3:].reshape(-1, 3)
    wl = embed(G.isfree().reshape(rz, ry, rx), v.min(0), v.max(0))
    wl["ntri"] = len(tris)
    return wl


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import welding_robot_b200 as wr
    from welding_robot_b200 import _lib
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with `python -m torch.distributed.run --nproc-per-node %d bench.py --gpus %d ...`" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().wr_set_device(local))
    torch.cuda.set_stream(torch.cuda.Stream())   # a stream of our own (the default stream cannot be captured into CUDA graphs); events, NCCL and the handles all use it
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = build_workload_gpu()
    n_nodes = CUBE ** 3
    colony = args.ants * world

    def make_search(host_free):
        acs = wr.ACS_Rank(seed=SEED, fixed_colony=colony, step_cap=STEP_CAP, update_mode=args.update_mode)
        acs.creatFromOccupancy(host_free, wl["xs"], wl["ys"], wl["zs"], PRECISION)     # H2D: occupancy + coordinates
        with contextlib.redirect_stdout(io.StringIO()):
            acs.initFromGridMap()
        return acs

    acs = make_search(wl["isfree"])
    stream = torch.cuda.current_stream()
    _lib.check(_lib.lib().wr_acs_set_stream(acs._a, stream.cuda_stream))
    acs.setEndpoints(wl["start"], wl["goal"])
    from welding_robot_b200.dist import ShardedSearch
    if world > 1:
        driver = ShardedSearch(acs, rank, world)
        driver.begin(PREDICT)
        step = lambda: driver.iterate(args.iters)  # noqa: E731
    else:
        acs.begin(PREDICT)
        step = lambda: acs.iterate(args.iters)  # noqa: E731

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- timed region, device resident -------------------------------------------------------------
    # Single GPU: the timed region runs as a user would run it (no per-phase events, so the steady-state iterations go out as
    # one CUDA graph each); the per-phase / per-kernel times are taken afterwards on a REPLAY of the same iterations (a fresh
    # handle, same seed: the search is deterministic) with the event pairs enabled.  Sharded: events inside the timed region.
    phase_in_timed = world > 1
    if phase_in_timed:
        acs.setTiming(True)
    c0 = acs.counters()
    rs0 = acs.updateStats()["rankset_iterations"]
    dirty0, tiles_total = acs.fieldStats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.mark_end()
    sampler.stop()
    c1 = acs.counters()
    if phase_in_timed:
        kms = acs.kernelMs()
        sk_ms, sk_n = acs.streamKernelMs()                  # the streaming kernel alone, by its own events inside the loop
    else:
        rep = make_search(wl["isfree"])
        _lib.check(_lib.lib().wr_acs_set_stream(rep._a, stream.cuda_stream))
        rep.setEndpoints(wl["start"], wl["goal"])
        rep.begin(PREDICT)
        for _ in range(args.warmup):
            rep.iterate(args.iters)
        rep.sync()
        rep.setTiming(True)
        for _ in range(args.steps):
            rep.iterate(args.iters)
        kms = rep.kernelMs()
        sk_ms, sk_n = rep.streamKernelMs()
        assert rep.counters()["ant_steps"] == c1["ant_steps"], "replay diverged from the timed search"
        del rep
    upd_stats = acs.updateStats()
    rs_iters = upd_stats["rankset_iterations"] - rs0        # timed iterations whose deposits went through rank sets (mode 4)
    dirty1, _ = acs.fieldStats()
    dirty_tiles = 0.5 * (dirty0 + dirty1)                    # tiles the evaporation pass streams (clean-tile field), mean over the timed region
    acs.setTiming(False)
    local_steps = c1["ant_steps"] - c0["ant_steps"]
    steps_done = local_steps
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        s = torch.tensor([local_steps], device="cuda", dtype=torch.int64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        steps_done = int(s.item())
    iters_done = args.steps * args.iters
    value = steps_done / (ms * 1e-3)

    # ---- kernels timed alone (burst roofline of K3): the single-pass fused kernel over the WHOLE field, i.e. on a handle whose
    #      field is materialised (every tile dirty, WR_LAZY_TAU=0) and that carries the deposit records of a real iteration -------
    alone = {}
    if world == 1 and args.workload == "C2":
        os.environ["WR_LAZY_TAU"] = "0"
        dense = wr.ACS_Rank(seed=SEED, fixed_colony=colony, step_cap=STEP_CAP, update_mode=0)
        dense.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], PRECISION)
        with contextlib.redirect_stdout(io.StringIO()):
            dense.initFromGridMap()
        del os.environ["WR_LAZY_TAU"]
        _lib.check(_lib.lib().wr_acs_set_stream(dense._a, stream.cuda_stream))
        dense.setEndpoints(wl["start"], wl["goal"])
        dense.begin(PREDICT)
        dense.iterate(3 * args.iters)
        for name, which in (("update_fused", 0), ("evaporate_float4", 1), ("d2d_copy", 2)):
            t_ms = dense.benchKernel(which, 20)
            alone[name] = {"ms": t_ms, "GBps": UPDATE_BYTES_PER_SLOT * n_nodes * 6 / (t_ms * 1e-3) / 1e9}
        alone["what"] = "dense field (every tile streamed), records of iteration %d of the same search; 805 MB per launch" % (3 * args.iters)
        del dense
        # the same kernel on the bench handle itself: its clean-tile field after the timed region, no records (rank-set phase)
        t_ms = acs.benchKernel(0, 20)
        dt, _ = acs.fieldStats()
        alone["update_fused_dirty_tiles_only"] = {"ms": t_ms, "GBps": dt * 4096 * UPDATE_BYTES_PER_SLOT / (t_ms * 1e-3) / 1e9, "tiles_streamed": dt,
                                                  "bytes": dt * 4096 * UPDATE_BYTES_PER_SLOT}

    # ---- the K = 26 neighbourhood (north_star's "26-neighbour" scores; the reference disables it, so it is reported next to
    #      the K = 6 parity workload, never instead of it): same grid, colony and endpoints, 112 B per ant-step --------------
    k26 = None
    if world == 1 and not args.no_k26:
        a26 = wr.ACS_Rank(seed=SEED, fixed_colony=args.ants, step_cap=STEP_CAP, update_mode=args.update_mode, K=26)
        a26.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], PRECISION)
        with contextlib.redirect_stdout(io.StringIO()):
            a26.initFromGridMap()
        _lib.check(_lib.lib().wr_acs_set_stream(a26._a, stream.cuda_stream))
        a26.setEndpoints(wl["start"], wl["goal"])
        a26.begin(PREDICT)
        a26.iterate(3 * args.iters)
        a26.sync()
        a26.setTiming(True)
        d0 = a26.counters()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n26 = 6 * args.iters
        f0.record(stream)
        a26.iterate(n26)
        f1.record(stream)
        torch.cuda.synchronize()
        ms26 = f0.elapsed_time(f1)
        d1 = a26.counters()
        k26ms = a26.kernelMs()
        st26 = d1["ant_steps"] - d0["ant_steps"]
        hbm26, _ = peaks()
        w26 = 112 * st26 / (k26ms["walk"] * 1e-3) / 1e9
        dt26, tt26 = a26.fieldStats()
        u26 = UPDATE_BYTES_PER_SLOT * dt26 * 4096 * n26 / (k26ms["update"] * 1e-3) / 1e9
        k26 = {"ant_steps_per_s": st26 / (ms26 * 1e-3), "acs_iterations_per_s": n26 / (ms26 * 1e-3), "iterations": n26,
               "mean_steps_per_ant": st26 / max(1, d1["ants"] - d0["ants"]), "arrived": d1["arrived"] - d0["arrived"],
               "kernel_ms_per_iteration": {k: v / n26 for k, v in k26ms.items()},
               "walk_GBps_at_112B_per_step": w26, "walk_frac_of_hbm": w26 / hbm26,
               "update_GBps": u26, "update_frac_of_hbm": u26 / hbm26, "pheromone_field_bytes": n_nodes * 26 * 4, "dirty_tiles": dt26, "tiles": tt26}
        del a26

    # ---- end to end through the public API from HOST buffers -------------------------------------------
    e2e = None
    if True:
        pinned = torch.from_numpy(wl["isfree"]).pin_memory()
        host_free = pinned.numpy()
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            a2 = make_search(host_free)
            a2.setEndpoints(wl["start"], wl["goal"])
            if world > 1:
                _lib.check(_lib.lib().wr_acs_set_stream(a2._a, stream.cuda_stream))
                d2 = ShardedSearch(a2, rank, world)
                d2.begin(PREDICT)                                                       # + exchange of the peer-slab IPC handles
                d2.iterate(args.iters)
            else:
                a2.begin(PREDICT)
                a2.iterate(args.iters)
            ids, dirs, L = a2.bestPath()                                                # D2H: the result
            n = a2.counters()["ant_steps"]
            del a2
            return n, len(ids) * 4 + len(dirs) + 4 + 9 * 8

        for _ in range(2):          # untimed: the first searches grow the stream-ordered memory pool by a second handle's worth
            e2e_step()
        tot_steps, d2h = 0, 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n, d2h = e2e_step()
            tot_steps += n
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            sN = torch.tensor([tot_steps], device="cuda", dtype=torch.int64)
            dist.all_reduce(sN, op=dist.ReduceOp.SUM)
            tot_steps = int(sN.item())
        e2e = {"value": tot_steps / dt, "unit": "ant-steps/s", "h2d_bytes_per_step": int(host_free.nbytes + 3 * CUBE * 4 + 16),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "what": "per step and per rank, from host buffers: wr_grid_create_from_occupancy (16.8 MB H2D) -> wr_acs_create -> wr_acs_begin -> "
                       "%d x iterate%s -> wr_acs_best (D2H); wall clock between barriers, max over ranks"
                       % (args.iters, " (sharded: + peer-slab handle exchange)" if world > 1 else "")}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    hbm, hbm_src = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")   # dram__bytes_read+write per launch from the committed ncu --set full captures
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}
    walk_ms = kms["walk"] / iters_done
    upd_ms = kms["update"] / iters_done
    walk_gbs = WALK_BYTES_PER_STEP * (local_steps / iters_done) / (walk_ms * 1e-3) / 1e9
    upd_bytes = dirty_tiles * 4096 * UPDATE_BYTES_PER_SLOT     # what the pass reads + writes: 4096-float tiles, 8 B per slot
    upd_phase_gbs = upd_bytes / (upd_ms * 1e-3) / 1e9
    upd_kernel_ms = sk_ms / max(1, sk_n) if world == 1 else upd_ms
    upd_gbs = upd_bytes / (upd_kernel_ms * 1e-3) / 1e9
    upd_dense_gbs = UPDATE_BYTES_PER_SLOT * n_nodes * 6 / (upd_ms * 1e-3) / 1e9
    ants_done = max(1, c1["ants"] - c0["ants"])
    out = {
        "metric": "ant-steps/s", "value": value, "unit": "ant-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic: reference mesh fixture voxelised on the GPU, embedded in 256^3 free space; synthetic start/goal",
        "config": {"workload": "%s, %d ants/GPU, K=6" % (LABEL, args.ants),
                   "grid": [CUBE, CUBE, CUBE], "natural_grid": list(wl["natural"]), "ants": colony, "iters_per_step": args.iters,
                   "step_cap": STEP_CAP, "seed": SEED, "update_mode": ["fused", "split", "atomic", "fused_tma", "rankset"][args.update_mode],
                   "parallelism": "ants sharded x%d" % world,
                   "exchange": ("none (1 GPU)" if world == 1 else
                                "NVLink peer memory inside the library's kernels (barrier flags, step counts, trails, rank-set blocks | final slot values); "
                                "no collective in the iteration loop"),
                   "l2_rule": "inputs larger than L2: the %d MB pheromone field is streamed from HBM every iteration" % (n_nodes * 6 * 4 // 1000000)},
        "acs_iterations_per_s": iters_done / (ms * 1e-3),
        "ant_steps": steps_done, "arrived_local": c1["arrived"] - c0["arrived"], "ants_local": c1["ants"] - c0["ants"],
        "mean_steps_per_ant": local_steps / ants_done,
        "gpu_launches": int(round(launches_per_iteration(0 if args.update_mode == 4 else args.update_mode, colony, args.iters, world > 1) * (iters_done - rs_iters)
                                  + launches_per_iteration(4, colony, args.iters, world > 1) * rs_iters)),
        "deposit_path": {"rank_set_iterations": rs_iters, "record_iterations": iters_done - rs_iters, "last": upd_stats} if args.update_mode == 4 else None,
        "kernel_ms_per_iteration": {k: v / iters_done for k, v in kms.items()},
        "kernel_ms_source": "events inside the timed region" if world > 1 else
                            "replay of the timed iterations on a fresh handle with per-phase events (the timed region itself runs without them)",
        "roofline": {"kernel": "k_walk2 (K2 ant construction; dominant by time)", "bound": "hbm", "achieved": walk_gbs, "peak": hbm, "unit": "GB/s",
                     "frac": walk_gbs / hbm, "traffic": traffic.get("k_walk2", traffic.get("k_walk")), "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": WALK_BYTES_PER_STEP * (local_steps / iters_done),
                     "note": "30 B algorithmic per ant-step; a walk is a chain of dependent gathers, so with 4096 ants the kernel is "
                             "latency-bound, not bandwidth-bound (see DESIGN.md)"},
        "roofline_update": {"kernel": ["k_update_fused (K3: evaporation + rank-ordered deposits, one HBM pass)", "k_evaporate + k_deposit_apply (K3 split)",
                                       "k_evaporate + atomic deposits (K3 atomic)", "(removed)",
                                       "k_update_fused on record-path iterations | k_evaporate_tiles (+ k_rankset_apply) on rank-set iterations (K3 adaptive)"][args.update_mode],
                            "bound": "hbm", "achieved": upd_gbs, "peak": hbm, "unit": "GB/s", "frac": upd_gbs / hbm,
                            "traffic": traffic.get("k_update_fused") if args.update_mode == 0 else (traffic.get("k_evaporate_tiles") if args.update_mode == 4 else None),
                            "algorithmic_bytes_per_launch": upd_bytes,
                            "reference_sweep_bytes": UPDATE_BYTES_PER_SLOT * n_nodes * 6, "reference_sweep_equivalent_GBps": upd_dense_gbs,
                            "dirty_tiles": dirty_tiles, "tiles": tiles_total, "kernel_ms": upd_kernel_ms, "launches_timed": sk_n,
                            "phase_ms": upd_ms, "phase_GBps": upd_phase_gbs,
                            "note": "clean-tile field: the pass streams only the 16 KB tiles that ever received a deposit (8 B per slot of those); "
                                    "`achieved` counts those bytes, `reference_sweep_equivalent_GBps` the reference's sweep of every slot (SURVEY 8d) over the same time",
                            "peak_source": hbm_src,
                            "timed": "the streaming kernel by its own CUDA-event pair inside the iteration loop (`kernel_ms`); `phase_ms` is the whole update "
                                     "phase (k_tile_offsets / k_rankset_apply and launch gaps included)"},
        "kernels_alone": alone,
        "voxelise": {"triangles": wl["ntri"], "grid": list(wl["natural"]), "kernel_ms": wl["vox"]["kernel_ms"], "tests": wl["vox"]["tests"]},
        "clocks": sampler.summary(),
    }
    if e2e:
        out["e2e"] = e2e
    if k26:
        out["k26_extension"] = k26
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(out))


def launches_per_iteration(update_mode, colony=ANTS_PER_GPU, iters=5, sharded=False):
    """Kernels of OURS launched per ACS iteration (welding_robot_b200/csrc/acs.cu: wr_acs_iterate / the sharded sequence)."""
    cap_bits = int(np.ceil(np.log2(STEP_CAP + 2)))
    slot_bits = int(np.ceil(np.log2(CUBE ** 3 * 6)))
    sort = lambda bits: 3 * ((bits + 9) // 10)  # noqa: E731  hist + scan + scatter per pass of <= 10-bit digits (radix_sort.cu)
    rank = 1 if colony <= 16384 else 1 + sort(cap_bits) + 2      # k_rank_small | keys + sort + finish + best clear
    # L2 warm-up, iter_begin, walk pass 1 + 2, ranking, best copy, iter_end (single GPU: once per wr_acs_iterate call)
    n = 1 + 1 + 2 + rank + 1 + (1 if sharded else 1.0 / iters)
    if update_mode == 2 or (update_mode == 4 and not sharded):
        return n + (2 if update_mode == 2 else 3)                 # rank-set: gen, evaporate, apply
    n += 1 + sort(slot_bits) + 2                                  # deposit gen, slot sort, (tile offsets + fused) | (evaporate + apply)
    if sharded:
        n += 6                                                    # partition pass (4), pull of the peers' final values, memset of the queue words
    return n


# ---------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref = UNMODIFIED reference headers), or the oracle port
# ---------------------------------------------------------------------------------------------------
def reference_worker(args):
    """One process: `warmup + steps` first-iterations of computeSolution (ACSRank_3D.hpp:220-305) with a
    colony of `cpu_ants` on the C2 grid.  Prints {"kind", "ant_steps": [...], "seconds": [...], "init_s"}."""
    from oracle import oracle as O
    wl = build_workload_cpu()
    seed = SEED + 1000 * args.worker
    # the reference sizes its colony as int(0.35*predict/precision) while no path is known (:247)
    predict = (args.cpu_ants + 0.5) * PRECISION / 0.35
    t0 = time.perf_counter()
    use_ref = O.have_ref() and not args.port
    steps, secs = [], []
    if use_ref:
        R = O.Ref()
        c0, c1 = wl["cmin"], wl["cmax"]
        dummy = np.array([[0, 0, 1, c0[0], c0[1], c0[2], c1[0], c1[1], c1[2], c0[0], c1[1], c0[2]]], np.float32)
        R.voxelize(dummy, PRECISION, WALL)      # same lattice: cubic box through creatGridMap's own formula
        assert R.dims == (CUBE, CUBE, CUBE), R.dims
        _, xs, ys, zs = R.grid()
        assert np.array_equal(xs, wl["xs"]) and np.array_equal(ys, wl["ys"]) and np.array_equal(zs, wl["zs"])
        R.set_free(wl["isfree"])
        R.acs_init()
        assert R.set_endpoints(wl["start"], wl["goal"])
        init_s = time.perf_counter() - t0
        for _ in range(args.warmup + args.steps):
            t1 = time.perf_counter()
            calls = R.compute(predict, 1, seed)    # every successful step draws once (ACSRank_3D.hpp:169)
            secs.append(time.perf_counter() - t1); steps.append(int(calls))
    else:
        G = O.Grid.from_occupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], PRECISION)
        A = O.Acs(G, seed=seed, rng_mode=O.RNG_SEQUENTIAL, sort_mode=O.SORT_STD)
        A.set_endpoints(wl["start"], wl["goal"])
        init_s = time.perf_counter() - t0
        for _ in range(args.warmup + args.steps):
            before = A.counters()["ant_steps"]
            t1 = time.perf_counter()
            A.begin(predict); A.iterate(1)
            secs.append(time.perf_counter() - t1); steps.append(A.counters()["ant_steps"] - before)
    print(json.dumps({"kind": "reference" if use_ref else "port", "ant_steps": steps[args.warmup:], "seconds": secs[args.warmup:], "init_s": init_s}))


def run_reference_pool(args, steps, warmup, workers):
    """`workers` concurrent single-thread processes (the reference has no threads: independent searches
    are the only parallelism that keeps its arithmetic)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--worker-mode", "--steps", str(steps), "--warmup", str(warmup),
           "--cpu-ants", str(args.cpu_ants)] + (["--port"] if args.port else [])
    procs = [subprocess.Popen(cmd + ["--worker", str(i)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(workers)]
    res = []
    for p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference worker failed:\n" + err[-2000:])
        res.append(json.loads(out.strip().splitlines()[-1]))
    tot = sum(sum(r["ant_steps"]) for r in res)
    wall = max(sum(r["seconds"]) for r in res)
    return res, tot, wall


def host_workers(args):
    cores = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
    except Exception:
        avail_kb = 16 << 20
    by_mem = max(1, int(avail_kb * 0.6 / (7 << 20)))    # ~6 GB per unmodified-reference process at 256^3
    return max(1, min(cores, by_mem, args.max_workers))


def cpu_baseline(args):
    """Bounded sample for the default run: ONE process, 1 warm-up + 2 timed first-iterations."""
    res, tot, wall = run_reference_pool(args, steps=2, warmup=1, workers=1)
    return {"value": tot / wall, "unit": "ant-steps/s", "cores": 1, "kind": res[0]["kind"],
            "sample": "2 iterations of computeSolution with %d of the %d ants on the same 256^3 grid (evaporation sweep of the whole field "
                      "included), 1 process" % (args.cpu_ants, ANTS_PER_GPU),
            "seconds": wall, "init_seconds": res[0]["init_s"], "ant_steps": tot}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 4)), min(args.warmup, 1)
    workers = host_workers(args)
    res, tot, wall = run_reference_pool(args, steps, warmup, workers)
    value = tot / wall
    print(json.dumps({
        "impl": "reference", "metric": "ant-steps/s", "value": value, "unit": "ant-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic: same C2 grid as the GPU arm, built on the CPU",
        "config": {"workload": "%s, %d ants/GPU, K=6" % (LABEL, ANTS_PER_GPU),
                   "grid": [CUBE, CUBE, CUBE], "sample_ants": args.cpu_ants, "processes": workers},
        "cpu_baseline": {"value": value, "unit": "ant-steps/s", "cores": workers, "kind": res[0]["kind"],
                         "sample": "%d timed first-iterations of computeSolution (ACSRank_3D.hpp:220-305) with %d-ant colonies on the 256^3 grid, "
                                   "%d concurrent single-thread processes (independent searches), evaporation sweep included"
                                   % (steps, args.cpu_ants, workers)},
        "e2e": {"value": value, "unit": "ant-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5, help="ACS iterations per step")
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"], help="C2 = the bench workload; C3 = the 512^3 scaling case (exploration run)")
    ap.add_argument("--ants", type=int, default=0, help="ants per GPU (default: the workload's colony; other values are exploration runs)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--update-mode", type=int, default=4, help="WR_UPDATE_*: 4 = adaptive rank sets | sorted records + fused pass (default), 0 = fused")
    ap.add_argument("--cpu-ants", type=int, default=1024, help="colony size of the CPU sample")
    ap.add_argument("--max-workers", type=int, default=16)
    ap.add_argument("--port", action="store_true", help="time the oracle port instead of oracle/_ref")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-k26", action="store_true", help="skip the K = 26 extension leg")
    ap.add_argument("--worker-mode", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--worker", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    select_workload(args.workload)
    if args.ants <= 0:
        args.ants = ANTS_PER_GPU
    if args.workload != "C2":
        args.no_k26 = True
    if args.impl == "reference":
        if args.worker_mode:
            reference_worker(args)
        else:
            run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
