#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native scan-matching backend.

Workload (BASELINE.json configs[1], SURVEY.md section 8(d) "C2"): real-time correlative sweep,
1081-beam 270-degree scans, +-0.5 m / +-30 deg window at 0.05 m / 0.5 deg, lowRes 5,
ScanRangeMax 5.7296 m, against one ~800x800-cell map.  One STEP = the coarse win-max
precompute of the map + one batch of `--matches` independent matches (different scans and
perturbed initial poses) through the C ABI.  metric = pose hypotheses scored per second
(fine + coarse hypotheses, each a full sum over the kept beams).

  value      kernels only, inputs resident in HBM, CUDA events on the context stream,
             max over ranks (N > 1: every rank runs its own batch = weak scaling, no collective;
             the single-scan front-end match does not shard).
  e2e        same metric through the public API with HOST buffers: grid + scans H2D, kernels,
             result records D2H, wall clock with a device sync on both sides.
  roofline   the sweep kernel: algorithmic gathered bytes / its CUDA-event duration vs the
             measured HBM copy peak (MEASURED_PEAKS.json).  The map is L1/L2 resident by
             design, so frac > 1 is expected; DESIGN.md explains the L1-wavefront bound.
  cpu_baseline / --impl reference
             the UNMODIFIED reference matcher (oracle/_ref/liblgs_ref.so) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from my_lidar_graph_slam_b200 import synth  # noqa: E402

C2 = dict(low_res=5, range_x=1.0, range_y=1.0, range_theta=1.0471975512, scan_range_max=5.7296)
METRIC = "pose hypotheses scored/sec (real-time correlative sweep, C2)"
UNIT = "hypotheses/s"


def c2_workload(n_matches: int, seed: int = 1):
    """Map from 10 scans + n_matches (scan, perturbed initial pose) pairs, all seeded."""
    world = synth.RoomsWorld(40.0, 5.0, seed=seed + 1)   # office-like: most ranges < 5.7 m
    angles = synth.beam_angles(1081, 270.0)
    traj = synth.trajectory(world, 10, step=0.4, seed=seed + 1)
    noise = np.random.default_rng(seed + 1)           # range noise
    map_scans = [synth.make_scan(world, p, angles, noise) for p in traj]
    dense, min_x, min_y = synth.rasterize_map(traj, angles, map_scans)
    pert = np.random.default_rng(seed + 2)            # pose perturbations
    ranges, inits = [], []
    while len(ranges) < n_matches:
        base = traj[pert.integers(0, len(traj))]
        true = base + np.array([pert.uniform(-0.6, 0.6), pert.uniform(-0.6, 0.6),
                                pert.uniform(-0.5, 0.5)])
        if not world.is_free(true[0], true[1], 0.3):
            continue
        ranges.append(synth.make_scan(world, true, angles, noise))
        inits.append(true + np.array([pert.uniform(-0.3, 0.3), pert.uniform(-0.3, 0.3),
                                      pert.uniform(-0.2, 0.2)]))
    return dense, min_x, min_y, angles, ranges, np.asarray(inits)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ref_time_matches(dense, min_x, min_y, angles, ranges, inits, n_threads):
    """Time the reference matcher (incl. its own coarse-map precompute, as OptimizePose(query)
    does once per call) over the given matches on n_threads host threads."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import refapi as R
    refmap = R.RefMap.from_dense(dense, min_x, min_y)

    def one(k):
        return R.rtcsm_match(refmap, angles, ranges[k], inits[k], **C2)   # pre=None: precompute inside

    one(0)   # warm-up (page in, allocate)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=n_threads) as ex:   # ctypes releases the GIL
        res = list(ex.map(one, range(len(ranges))))
    return time.perf_counter() - t0, res


def hyps_per_match(res):
    nbx = (2 * res.winX) // 5 + 1
    nby = (2 * res.winY) // 5 + 1
    return (2 * res.winT + 1) * (nbx * 5 * nby * 5 + nbx * nby)


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    from oracle import refapi as R
    cores = os.cpu_count() or 1
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/liblgs_ref.so not built"}))
        return
    per_step = max(cores, 8)
    total = per_step * (args.steps + args.warmup)
    dense, min_x, min_y, angles, ranges, inits = c2_workload(total, seed=1)
    times, hyps = [], 0
    for s in range(args.steps + args.warmup):
        sl = slice(s * per_step, (s + 1) * per_step)
        dt, res = ref_time_matches(dense, min_x, min_y, angles, ranges[sl], inits[sl], cores)
        if s >= args.warmup:
            times.append(dt)
            hyps += sum(hyps_per_match(r) for r in res)
    total_t = sum(times)
    value = hyps / total_t
    sample = (f"{per_step} C2 matches per step x {args.steps} steps on {cores} threads; "
              "hypotheses = full window size per match (the CPU prunes inside it)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2 real-time correlative sweep, reference CPU matcher",
                   "matches_per_step": per_step, **C2},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args, rank, world_size, local_rank):
    from my_lidar_graph_slam_b200 import capi
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    M = args.matches
    dense, min_x, min_y, angles, ranges, inits = c2_workload(M, seed=1 + rank)
    ctx = capi.Context(local_rank)
    grid = capi.Grid.from_dense(ctx, dense, min_x, min_y, 0.05, apron=32)
    coarse = grid.like()
    scans = capi.Scans([angles] * M, ranges, inits)
    batch = capi.RtcsmBatch(ctx, **C2)
    lib = capi.lib()

    def step_device():
        ctx.check(lib.lgs_precompute(ctx.h, grid.h, 5, coarse.h))
        batch.run(grid, coarse)

    def step_e2e():
        grid.upload(dense)
        ctx.check(lib.lgs_precompute(ctx.h, grid.h, 5, coarse.h))
        batch.upload(grid, scans)
        batch.run(grid, coarse)
        return batch.results(grid, coarse)

    # ---- device-resident timing ------------------------------------------------------------
    batch.upload(grid, scans)
    hyp, gathers = batch.work()
    for _ in range(max(args.warmup, 3)):
        step_device()
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    total_hyp = sum_over_ranks(float(hyp)) * args.steps
    value = total_hyp / (ms * 1e-3)

    # ---- per-kernel timing for the roofline (sweep kernel) ---------------------------------------
    sweep_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        ctx.check(lib.lgs_precompute(ctx.h, grid.h, 5, coarse.h))
        sweep_ms.append(batch.run_timed(grid, coarse))
    k_proj, k_sweep, k_sel = (statistics.mean(x[i] for x in sweep_ms) for i in range(3))
    peak, peak_src = measured_peaks()
    achieved = gathers * 8 / (k_sweep * 1e-3) / 1e9
    results_dev = batch.results(grid, coarse)

    # ---- end to end through the public API with host buffers ---------------------------------
    for _ in range(2):
        step_e2e()
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step_e2e()
    ctx.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total_hyp / e2e_s
    h2d = dense.nbytes + scans.nbytes + M * 96            # grid + scans + match descriptors
    d2h = M * 32 + 4                                      # result records + fix-up counter
    assert all((a.found, a.ix, a.iy, a.it, a.score) == (b.found, b.ix, b.iy, b.it, b.score)
               for a, b in zip(results_dev, res_e2e))

    if rank != 0:
        return
    # ---- CPU baseline: the unmodified reference on a bounded sample, parity-checked ------------
    cpu = None
    if world_size == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refapi as R
            if R.available():
                cores = os.cpu_count() or 1
                ns = min(M, max(2 * cores, 16))
                dt, ref = ref_time_matches(dense, min_x, min_y, angles, ranges[:ns], inits[:ns], cores)
                bad = sum((a.found, a.ix, a.iy, a.it, a.score) != (b.found, b.ix, b.iy, b.it, b.score)
                          for a, b in zip(ref, results_dev[:ns]))
                cpu = {"value": sum(hyps_per_match(r) for r in ref) / dt, "unit": UNIT,
                       "cores": cores, "kind": "reference",
                       "sample": f"first {ns} matches of the step on {cores} threads "
                                 f"({dt:.1f} s); hypotheses = full window per match; "
                                 f"GPU winners/scores identical on {ns - bad}/{ns}"}
        except Exception as e:   # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 real-time correlative sweep: 1081 beams 270 deg, +-0.5 m/+-30 deg "
                               "at 0.05 m/0.5 deg, lowRes 5, one map of %dx%d cells" % dense.shape[::-1],
                   "matches_per_step": M, "hypotheses_per_step": hyp, **C2,
                   "l2": "no flush: per-step working set (projected offsets + score tables) %.0f MB > 126 MB "
                         "L2; the map itself is cache-resident by design"
                         % ((hyp * 8 + hyp / 650.0 * (gathers / max(hyp, 1)) * 4) / 1e6),
                   "parallelism": "replicas" if world_size > 1 else "1 GPU"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "kernel": "csm_sweep_kernel",
                     "peak_source": peak_src, "kernel_ms": k_sweep,
                     "algorithmic_bytes_per_launch": gathers * 8,
                     "note": "gathers are served from L1/L2 (map is cache resident), so the HBM "
                             "roofline is not the binding limit; see DESIGN.md"},
        "kernel_ms": {"csm_project": k_proj, "csm_sweep": k_sweep, "csm_select": k_sel},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--matches", type=int, default=1000, help="matches per step (C2: 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_b200(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
