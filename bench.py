#!/usr/bin/env python
"""bench.py — ant-steps/s of the rank-based 3-D ACS search on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU code (oracle/_ref)

Workload (config.workload "C2"): BASELINE.json configs[1] — the simplified_piece mesh voxelised on
the GPU at a 256-long grid (precision 0.823812/235.5, wall 10) and embedded in 256^3 free space
(SURVEY.md §8d), 4096 ants per GPU, K = 6, Philox seed 1, start/goal at opposite corners inset by 5
cells.  One "step" = `--iters` ACS iterations (ant construction + ranking + fused pheromone update)
of one search.  With N GPUs the colony is 4096*N ants sharded by ant index (weak scaling), one
exchange per iteration.  The pheromone field (403 MB) is larger than L2 (126 MB), so every
iteration streams it from HBM.

Prints ONE JSON line (see the contract in the task description): `value` is whole-job
ant-steps/s with the grid already in HBM; `e2e` goes through the public API from HOST buffers
(occupancy upload, handle creation, search, best-path download) every step.
"""
import argparse
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
PREDICT = 1.0   # unused with a fixed colony except for Q of the first iterations (ACSRank_3D.hpp:249)
WALK_BYTES_PER_STEP = 30     # 4*K tau + K/8 occupancy + 4 id + 1 dir, K = 6
UPDATE_BYTES_PER_SLOT = 8    # 4 read + 4 write per directed slot per iteration


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag:
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload():
    """Host copy of the C2 occupancy: GPU-voxelised natural grid embedded in CUBE^3 free space."""
    import welding_robot_b200 as wr
    tris = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))["simplified_piece"]
    g = wr.GridMap()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        g.creatGridMap(tris, PRECISION, WALL)
    rx, ry, rz = g.rangeX, g.rangeY, g.rangeZ
    assert max(rx, ry, rz) == CUBE, (rx, ry, rz)
    free = g.isfree().reshape(rz, ry, rx)
    xs, ys, zs = g.coords()
    vox = g.stats()
    p = np.float32(PRECISION)

    def extend(c, n):
        lo = (n - len(c)) // 2
        hi = n - len(c) - lo
        return np.concatenate([c[0] - p * np.arange(lo, 0, -1, dtype=np.float32), c, c[-1] + p * np.arange(1, hi + 1, dtype=np.float32)]).astype(np.float32), lo

    X, ox = extend(xs, CUBE); Y, oy = extend(ys, CUBE); Z, oz = extend(zs, CUBE)
    cube = np.ones((CUBE, CUBE, CUBE), np.uint8)
    cube[oz:oz + rz, oy:oy + ry, ox:ox + rx] = free
    nid = lambda x, y, z: (z * CUBE + y) * CUBE + x  # noqa: E731
    start, goal = nid(5, 5, 5), nid(CUBE - 6, CUBE - 6, CUBE - 6)
    assert cube.ravel()[start] and cube.ravel()[goal]
    return dict(isfree=np.ascontiguousarray(cube.ravel()), xs=X, ys=Y, zs=Z, start=start, goal=goal, natural=(rx, ry, rz), vox=vox, ntri=len(tris))


def run_ours(args):
    import torch
    import welding_robot_b200 as wr
    from welding_robot_b200 import _lib
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().wr_set_device(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = build_workload()
    n_nodes = CUBE ** 3
    colony = ANTS_PER_GPU * world

    def make_search():
        acs = wr.ACS_Rank(seed=SEED, fixed_colony=colony, step_cap=STEP_CAP, update_mode=args.update_mode)
        acs.creatFromOccupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], PRECISION)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            acs.initFromGridMap()
        return acs

    acs = make_search()
    stream = torch.cuda.current_stream()
    _lib.check(_lib.lib().wr_acs_set_stream(acs._a, stream.cuda_stream))
    if world > 1:
        from welding_robot_b200.dist import ShardedSearch
        driver = ShardedSearch(acs, rank, world)
    else:
        driver = None
    acs.setEndpoints(wl["start"], wl["goal"])
    acs.begin(PREDICT) if driver is None else driver.begin(PREDICT)

    def step():
        if driver is None:
            acs.iterate(args.iters)
        else:
            driver.iterate(args.iters)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- timed region: device-resident (value) ---------------------------------------------------
    acs.setTiming(True)
    c0 = acs.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    c1 = acs.counters()
    kms = acs.kernelMs()
    acs.setTiming(False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        s = torch.tensor([c1["ant_steps"] - c0["ant_steps"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        steps_done = int(s.item())
    else:
        steps_done = c1["ant_steps"] - c0["ant_steps"]
    iters_done = args.steps * args.iters
    value = steps_done / (ms * 1e-3)

    # ---- end to end through the public API from HOST buffers ---------------------------------------
    e2e = None
    if world == 1:
        pinned = torch.from_numpy(wl["isfree"]).pin_memory()
        host_free = pinned.numpy()
        e2e_steps = max(1, min(args.steps, 5))
        tot_steps = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            a2 = wr.ACS_Rank(seed=SEED, fixed_colony=colony, step_cap=STEP_CAP, update_mode=args.update_mode)
            a2.creatFromOccupancy(host_free, wl["xs"], wl["ys"], wl["zs"], PRECISION)     # H2D: occupancy + coordinates
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                a2.initFromGridMap()
            a2.setEndpoints(wl["start"], wl["goal"])
            a2.begin(PREDICT)
            a2.iterate(args.iters)
            ids, dirs, L = a2.bestPath()                                                # D2H: the result
            tot_steps += a2.counters()["ant_steps"]
            d2h = ids.nbytes // 2 + dirs.nbytes // 4 + 4 + 9 * 8
            del a2
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": tot_steps / dt, "unit": "ant-steps/s", "h2d_bytes_per_step": int(host_free.nbytes + 3 * CUBE * 4 + 16),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "what": "per step: occupancy upload -> wr_grid_create_from_occupancy -> wr_acs_create -> begin -> %d iterations -> wr_acs_best" % args.iters}
    sampler.stop_flag = True

    if rank != 0:
        return
    hbm, hbm_src = peaks()
    walk_ms = kms["walk"] / iters_done
    upd_ms = kms["update"] / iters_done
    local_steps = (c1["ant_steps"] - c0["ant_steps"])
    walk_gbs = WALK_BYTES_PER_STEP * (local_steps / iters_done) / (walk_ms * 1e-3) / 1e9
    upd_gbs = UPDATE_BYTES_PER_SLOT * n_nodes * 6 / (upd_ms * 1e-3) / 1e9
    out = {
        "metric": "ant-steps/s", "value": value, "unit": "ant-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic: reference mesh fixture voxelised on the GPU, embedded in 256^3 free space; synthetic start/goal",
        "config": {"workload": "C2: ACSRank_3D, simplified_piece @256-long grid in 256^3, %d ants/GPU, K=6" % ANTS_PER_GPU,
                   "grid": [CUBE, CUBE, CUBE], "natural_grid": list(wl["natural"]), "ants": colony, "iters_per_step": args.iters,
                   "step_cap": STEP_CAP, "seed": SEED, "update_mode": ["fused", "split", "atomic"][args.update_mode],
                   "parallelism": "ants sharded x%d" % world,
                   "l2_rule": "pheromone field 403 MB > 126 MB L2: every iteration streams it from HBM (no flush needed)"},
        "acs_iterations_per_s": iters_done / (ms * 1e-3),
        "ant_steps": steps_done, "arrived": c1["arrived"] - c0["arrived"], "ants": c1["ants"] - c0["ants"],
        "gpu_launches": None,
        "kernel_ms_per_iteration": {k: v / iters_done for k, v in kms.items()},
        "roofline": {"kernel": "k_walk (K2 ant construction, dominant by time)", "bound": "hbm", "achieved": walk_gbs, "peak": hbm, "unit": "GB/s",
                     "frac": walk_gbs / hbm, "traffic": None, "peak_source": hbm_src,
                     "note": "30 B algorithmic per ant-step; latency-bound: %d dependent steps per ant" % (local_steps // max(1, c1["ants"] - c0["ants"]))},
        "roofline_update": {"kernel": "k_update_fused (K3 evaporate+deposit, TMA)" if args.update_mode == 0 else "k_evaporate+deposit", "bound": "hbm",
                            "achieved": upd_gbs, "peak": hbm, "unit": "GB/s", "frac": upd_gbs / hbm, "traffic": None, "peak_source": hbm_src},
        "voxelise": {"triangles": wl["ntri"], "grid": list(wl["natural"]), "kernel_ms": wl["vox"]["kernel_ms"], "tests": wl["vox"]["tests"]},
        "clocks": sampler.summary(),
    }
    if e2e:
        out["e2e"] = e2e
    launches_per_iter = launches_per_iteration(acs, args.update_mode)
    out["gpu_launches"] = launches_per_iter * iters_done
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(wl, sample_ants=args.cpu_ants)
    print(json.dumps(out))


def launches_per_iteration(acs, update_mode):
    """Kernels of OURS launched per ACS iteration (see welding_robot_b200/csrc/acs.cu)."""
    cap_bits = int(np.ceil(np.log2(STEP_CAP + 2)))
    slot_bits = int(np.ceil(np.log2(CUBE ** 3 * 6)))
    sort = lambda bits: 3 * ((bits + 7) // 8)  # noqa: E731  hist+scan+scatter per 8-bit pass
    n = 1 + 3 + 1 + sort(cap_bits) + 1 + 2 + 1   # iter_begin, walk x2 + queue reset, rank keys, sort, rank finish, best x2, iter_end
    if update_mode == 2:
        return n + 2
    return n + 1 + sort(slot_bits) + 2           # deposit gen, sort, (tile offsets + fused) or (evaporate + apply)


def reference_sample(wl, sample_ants, seed):
    """One iteration of the UNMODIFIED reference (oracle/_ref) on the C2 grid with `sample_ants` ants.
    The reference sizes its colony as int(0.35*predict/precision) (ACSRank_3D.hpp:247), so predict
    is chosen to give exactly the sample size."""
    from oracle import oracle as O
    R = O.Ref()
    # a 1-triangle mesh would not reproduce the grid: drive the reference on the SAME occupancy by
    # voxelising a tiny mesh of the right extent and overwriting isFree (coordinates are its own).
    raise NotImplementedError


def cpu_baseline(wl, sample_ants):
    """The CPU oracle (a port of the reference's arithmetic, single thread like the reference) on a
    bounded sample of the same workload: ONE iteration of `sample_ants` of the 4096 ants on the
    same 256^3 grid, same seed."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    G = O.Grid.from_occupancy(wl["isfree"], wl["xs"], wl["ys"], wl["zs"], PRECISION)
    A = O.Acs(G, seed=SEED, fixed_colony=sample_ants, step_cap=STEP_CAP)
    A.set_endpoints(wl["start"], wl["goal"])
    A.begin(PREDICT)
    t1 = time.perf_counter()
    A.iterate(1)
    t2 = time.perf_counter()
    c = A.counters(); ph = A.phase_seconds()
    return {"value": c["ant_steps"] / (t2 - t1), "unit": "ant-steps/s", "cores": 1, "kind": "port",
            "sample": "1 iteration, %d of %d ants, full 256^3 grid (evaporation sweep included)" % (sample_ants, ANTS_PER_GPU),
            "seconds": t2 - t1, "init_seconds": t1 - t0, "ant_steps": c["ant_steps"],
            "phase_seconds": {k: float(v) for k, v in ph.items()},
            "walk_only_ant_steps_per_s": c["ant_steps"] / max(ph["walk"], 1e-9),
            "evaporate_GBps": UPDATE_BYTES_PER_SLOT * CUBE ** 3 * 6 / max(ph["evaporate"], 1e-9) / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_reference import run as run_ref
    print(json.dumps(run_ref(args, build_workload_cpu())))


def build_workload_cpu():
    raise NotImplementedError


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5, help="ACS iterations per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--update-mode", type=int, default=0)
    ap.add_argument("--cpu-ants", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
