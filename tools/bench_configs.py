#!/usr/bin/env python
"""Per-scenario throughput and RMSE-vs-time on one B200 next to the CPU oracle (BASELINE.md §3 plan).

    python tools/bench_configs.py [--out gpurun_out/configs.json] [--quick]

Not the driver's bench (that is bench.py); this fills the per-config table in DESIGN.md / profiles/.
For each SURVEY §8(d) scenario: walk-steps/s on the GPU (CUDA events, device-resident inputs, 3 warm-up + 5 timed
passes), steps per walk, the CPU oracle (C port, all host cores) on a bounded sample, and for the scenarios with an
analytic solution RMSE over the evaluation points against wall time for growing walk counts.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from dcrmontecarlo_b200 import _native as nat  # noqa: E402
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402
from oracle import wost_oracle as orc  # noqa: E402


def f_step(s):
    """Algorithmic fp32 flops per walk step; fields counted at 14 flops per evaluation."""
    SD = len(s.dirichlet) - 1
    f = 27 * SD + 12
    if s.neumann is not None:
        VN = len(s.neumann)
        f += 23 * (VN - 2) + 24 * (VN - 1) + 24
    if s.f is not None:
        f += 22 + 14
    if s.delta:
        f += 25 + 3 * 14 + 0.5 * 14
    return float(f)


def gpu_rate(s, points, walks, reps=5, warm=3):
    solver = s.make_solver()
    pts = points.cuda()
    for i in range(warm):
        solver.solve_raw(pts, walks, s.max_steps, s.eps, seed=10 + i, device_outputs=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    outs = []
    e0.record()
    for i in range(reps):
        outs.append(solver.solve_raw(pts, walks, s.max_steps, s.eps, seed=100 + i, device_outputs=True)["steps"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = sum(int(o[0]) for o in outs)
    return dict(steps_per_s=steps / (ms * 1e-3), ms_per_pass=ms / reps, steps_per_walk=steps / (reps * len(points) * walks),
                points=len(points), walks=walks)


def cpu_rate(s, points, target_s=4.0):
    cores = len(os.sched_getaffinity(0))
    prob = orc.Problem.from_scenario(s, sigma_bar=s.make_solver().sigma_bar if s.delta else 0.0)
    pts = points[: max(cores * 4, 16)]
    t0 = time.perf_counter()
    r = prob.solve(pts, 8, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=1, n_threads=cores)
    rate = r["steps"] / (time.perf_counter() - t0)
    walks = int(max(8, min(4096, rate * target_s / max(r["steps"] / 8, 1))))
    t0 = time.perf_counter()
    r = prob.solve(pts, walks, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=2, n_threads=cores)
    dt = time.perf_counter() - t0
    return dict(steps_per_s=r["steps"] / dt, cores=cores, points=len(pts), walks=walks, seconds=dt)


def tile_points(points, n):
    reps = (n + len(points) - 1) // len(points)
    return points.repeat(reps, 1)[:n].contiguous()


def rmse_vs_time(s, walks_list, cpu_walks_list):
    solver = s.make_solver()
    exact = s.analytic(s.points).double()
    prob = orc.Problem.from_scenario(s, sigma_bar=solver.sigma_bar if s.delta else 0.0)
    cores = len(os.sched_getaffinity(0))
    solver.solve(s.points, nWalks=64, maxSteps=s.max_steps, eps=s.eps, seed=0)          # warm-up
    out = {"gpu": [], "cpu": []}
    for W in walks_list:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        est = solver.solve(s.points, nWalks=W, maxSteps=s.max_steps, eps=s.eps, seed=1000 + W)   # host in, host out
        dt = time.perf_counter() - t0
        out["gpu"].append(dict(walks=W, wall_s=dt, rmse=float(torch.sqrt(((est[:, 0].double() - exact) ** 2).mean()))))
    for W in cpu_walks_list:
        t0 = time.perf_counter()
        r = prob.solve(s.points, W, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=1000 + W, n_threads=cores)
        dt = time.perf_counter() - t0
        out["cpu"].append(dict(walks=W, wall_s=dt, rmse=float(np.sqrt(((r["mean"] - exact.numpy()) ** 2).mean())), cores=cores))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "configs.json"))
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    nat.require_cuda()
    peak_tf, mhz = nat.fp32_peak(0)
    N = 16384 if a.quick else 65536
    plan = [
        ("cfg1a", sc.cfg1a(), N, 256), ("cfg1b", sc.cfg1b(), N, 64), ("cfg2", sc.cfg2(), N, 256), ("cfg3", sc.cfg3(), N, 256),
        ("cfg4", sc.cfg4(), N, 64), ("cfg5_9e", sc.cfg5(9), 9, 32768 if not a.quick else 8192),
        ("cfg5_175e", sc.cfg5(175), 175, 4096 if not a.quick else 1024),
        ("scale_32", sc.scale_scene(32, N, 64), None, 64), ("scale_1024", sc.scale_scene(1024, N, 16), None, 16),
        ("scale_16384", sc.scale_scene(16384, N, 16), None, 16),
    ]
    res = {"fp32_peak_tflops": peak_tf, "fp32_peak_effective_sm_mhz": mhz, "configs": {}}
    for name, s, n_pts, walks in plan:
        pts = s.points if n_pts is None else tile_points(s.points, n_pts)
        g = gpu_rate(s, pts, walks)
        c = cpu_rate(s, pts) if not name.startswith("scale_16384") else cpu_rate(s, pts, target_s=2.0)
        F = f_step(s)
        g.update(f_step=F, fp32_frac=g["steps_per_s"] * F / 1e12 / peak_tf, speedup_vs_cpu_port=g["steps_per_s"] / c["steps_per_s"])
        res["configs"][name] = dict(gpu=g, cpu_port=c)
        print(name, json.dumps(res["configs"][name]), flush=True)
    if not a.quick:
        res["rmse_vs_time"] = {}
        for key in ("cfg1a", "cfg1b", "cfg3"):
            res["rmse_vs_time"][key] = rmse_vs_time(sc.ALL[key](), [10, 25, 50, 150, 600, 2400, 9600, 38400, 153600, 614400], [10, 25, 50, 150, 600, 2400, 9600])
            print(key, json.dumps(res["rmse_vs_time"][key]), flush=True)
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
