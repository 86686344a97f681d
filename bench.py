#!/usr/bin/env python
"""bench.py — WoSt walk-steps/s on B200(s) next to the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — the mixed Dirichlet/Neumann polyline domain with
reflecting walks (square +-2 Dirichlet, 32-gon r=0.5 Neumann, Laplace, g = x; SURVEY §8(d) cfg 2) — with a
throughput-sized evaluation set: POINTS uniform points per GPU, WALKS walks each.  One "step" = one pass of the hot
path (wost_solve) over that batch.  Weak scaling: every rank owns POINTS points of a global N*POINTS set; there is
no data-path collective (only the timing reduction), see DESIGN.md.

Prints ONE JSON line (rank 0).  `value` = whole-job walk-steps/s with inputs resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (H2D of the points and D2H of the statistics inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

POINTS = int(os.environ.get("WOST_BENCH_POINTS", 65536))
WALKS = int(os.environ.get("WOST_BENCH_WALKS", 256))
F_STEP_CFG2 = 1625.0          # algorithmic fp32 flops per walk step for S_D=4, V_N=33
# dram__bytes_read.sum + dram__bytes_write.sum of one walk_kernel launch of this workload, from the committed
# `ncu --set full` capture profiles/r1_v7_walk_kernel_ncu_full.csv (0.60 MB read + 14.77 MB written)
DRAM_TRAFFIC_PER_LAUNCH = 15_365_888
METRIC, UNIT = "wost_walk_steps_per_sec", "walk-steps/s"


def scenario(n_points):
    from dcrmontecarlo_b200 import scenarios as sc

    return sc.cfg2_throughput(n_points=n_points, n_walks=WALKS)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons |= {n for bit, n in names.items() if mask & bit}
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=10)                                            # an NVML query can take a while under load
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(n_threads=0, target_seconds=12.0):
    """The CPU oracle (C port of the reference walk loop, Philox mode, OpenMP over evaluation points) on a bounded
    sample of the same workload."""
    from oracle import wost_oracle as orc

    cores = n_threads or len(os.sched_getaffinity(0))
    s = scenario(4096)
    prob = orc.Problem.from_scenario(s)
    t0 = time.perf_counter()
    r = prob.solve(s.points[: 8 * cores], 16, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=1, n_threads=cores)
    rate = r["steps"] / (time.perf_counter() - t0)
    n_pts = len(s.points)
    walks = int(max(16, rate * target_seconds / (16.6 * n_pts)))
    t0 = time.perf_counter()
    r = prob.solve(s.points[:n_pts], walks, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=2, n_threads=cores)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    r1 = prob.solve(s.points[:256], max(16, walks // 64), s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=3, n_threads=1)
    single = r1["steps"] / (time.perf_counter() - t1)
    return {"value": r["steps"] / dt, "unit": UNIT, "cores": cores, "kind": "port", "single_core_value": single,
            "sample": f"{n_pts} points x {walks} walks of the same scene ({r['steps']} steps in {dt:.2f} s), oracle/wost_oracle.c, OpenMP over points",
            "python_reference_probe": "the unmodified Python reference measured ~2.7e3 walk-steps/s on 1 core for this scene (BASELINE.md §2); it cannot travel to the GPU box"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the reference itself is
    Python and cannot travel) with all host threads on the same config."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import wost_oracle as orc

    cores = len(os.sched_getaffinity(0))
    s = scenario(4096)
    prob = orc.Problem.from_scenario(s)
    # calibrate so that one step is ~3 s of all-core CPU work
    t0 = time.perf_counter()
    r = prob.solve(s.points[: 8 * cores], 16, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=1, n_threads=cores)
    rate = r["steps"] / (time.perf_counter() - t0)
    n_pts = len(s.points)
    walks = int(max(8, rate * 3.0 / (16.6 * n_pts)))
    t_steps, total = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = prob.solve(s.points[:n_pts], walks, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=100 + it, n_threads=cores)
        if it >= args.warmup:
            t_steps.append(time.perf_counter() - t0); total += r["steps"]
    T = sum(t_steps)
    val = total / T
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2 mixed Dirichlet(square +-2)/Neumann(32-gon r=0.5) Laplace g=x, eps=1e-4, maxSteps=500",
                       "points_per_step": n_pts, "walks_per_point": walks},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = {n_pts} points x {walks} walks, oracle/wost_oracle.c (C port of the reference loop), OpenMP"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def hbm_view(steps_per_launch, ms_per_launch):
    """Algorithmic bytes per launch (points in, per-walk totals out) over the launch time against the measured copy
    bandwidth (MEASURED_PEAKS.json, else the profiling recipe's 6.4 TB/s): three orders of magnitude below the roof."""
    peak, src = 6400.0, "fallback (B200_PROFILING.md)"
    try:
        peak, src = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    alg = POINTS * 8 + POINTS * WALKS * 4
    gbs = alg / (ms_per_launch * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_launch": alg,
            "peak_source": src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from dcrmontecarlo_b200 import _native as nat

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    nat.require_cuda()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    s = scenario(POINTS * world)
    solver = s.make_solver()
    my = s.points[rank * POINTS:(rank + 1) * POINTS].contiguous()
    pts_dev = my.cuda()
    pts_pinned = my.pin_memory()
    base = rank * POINTS

    def step_resident(it):
        return solver.solve_raw(pts_dev, WALKS, s.max_steps, s.eps, seed=1000 + it, point_index_base=base, device_outputs=True)

    def step_e2e(it):
        return solver.solve_raw(pts_pinned, WALKS, s.max_steps, s.eps, seed=1000 + it, point_index_base=base)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident arm ------------------------------------------------------------------------
    for it in range(args.warmup):
        step_resident(it)
    sync_all()
    sampler = ClockSampler(local); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    outs = []
    ev0.record()
    for it in range(args.steps):
        outs.append(step_resident(args.warmup + it)["steps"])
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    my_steps = int(sum(int(o[0]) for o in outs))
    # the clocks belong to the device-timed region; NVML polling stays out of the host-timed arm below
    clocks = sampler.stop()

    # ---- end-to-end arm: host buffers through the C ABI -----------------------------------------------
    for it in range(args.warmup):
        step_e2e(it)
    # Three passes of K steps, the fastest one counts and all are reported: this arm is timed on the host clock, so it sees
    # what the device-timed arm above does not (an unlucky scheduling of this process, a collector pause).
    import gc
    gc.collect(); gc.disable()                                           # no collector pauses inside the host-timed region
    e2e_passes = []
    for rep in range(3):
        sync_all()
        t0 = time.perf_counter()
        n_steps = 0
        for it in range(args.steps):
            n_steps += int(step_e2e(args.warmup + it)["steps"][0])    # returns after the D2H of the statistics
        torch.cuda.synchronize()
        e2e_passes.append((1e3 * (time.perf_counter() - t0), n_steps))
    gc.enable()
    e2e_ms, e2e_steps = min(e2e_passes)

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([my_steps, e2e_steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(t[0]), float(t[1])
    tot_steps, tot_e2e = float(c[0]), float(c[1])

    if rank == 0:
        value = tot_steps / (ms * 1e-3)
        peak_tf, eff_mhz = nat.fp32_peak(local)
        # dominant kernel = the walk kernel (the two statistics kernels take < 1 % of a step, see profiles/)
        achieved_tf = (my_steps / (ms * 1e-3)) * F_STEP_CFG2 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg2 mixed Dirichlet(square +-2)/Neumann(32-gon r=0.5) Laplace g=x, eps=1e-4, maxSteps=500 (BASELINE.json configs[1])",
                       "points_per_gpu": POINTS, "walks_per_point": WALKS, "walk_steps_per_step_per_gpu": my_steps / args.steps,
                       "l2_note": "each step rewrites a %d MiB per-walk buffer and uses a fresh Philox key; the working set is registers/shared memory, not L2" % (POINTS * WALKS * 4 >> 20),
                       "parallelism": f"points sharded over {world} GPU(s), no data-path collective", "compat": "reference"},
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": DRAM_TRAFFIC_PER_LAUNCH if (POINTS, WALKS) == (65536, 256) else None,
                         "traffic_note": "bytes per launch from profiles/r1_v7_walk_kernel_ncu_full.csv; algorithmic bytes per launch = "
                                         f"{POINTS * 8 + POINTS * WALKS * 4} (points in, per-walk totals out, mostly L2-resident)",
                         "flops_per_walk_step": F_STEP_CFG2,
                         "peak_source": "FMA-chain microbenchmark in this run (wost_fp32_peak); MEASURED_PEAKS.json has no fp32 entry",
                         "kernel": "walk_kernel<NEU=1,SRC=0,DELTA=0>", "fp32_peak_effective_sm_mhz": eff_mhz,
                         # the same launch seen as HBM traffic, to show which roof applies: algorithmic bytes / launch time
                         "hbm_view": hbm_view(my_steps / args.steps, ms / args.steps)},
            "e2e": {"value": tot_e2e / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": POINTS * 8, "d2h_bytes_per_step": POINTS * 16 + 8,
                    "ms_per_step": e2e_ms / args.steps, "passes_ms_per_step": [p[0] / args.steps for p in e2e_passes]},
            "gpu_launches": 3 * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
