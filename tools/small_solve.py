#!/usr/bin/env python
"""Time-to-solution of the reference's SHIPPED problem sizes: cfg 5 at 9 x 100, cfg 1b at 16 x 150, ...

    python tools/small_solve.py [reps]

Prints microseconds per solve, device-resident (inputs / outputs on the device, CUDA events over `reps` back-to-back
solves) and with host buffers through the C ABI (wall clock, H2D + D2H inside)."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import _native as nat  # noqa: E402
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402


def small_solve_us(s, reps=200, jit="auto"):
    solver = s.make_solver(jit=jit)
    pts_d, pts_h = s.points.cuda(), s.points.clone()
    W = s.n_walks
    for i in range(5):
        solver.solve_raw(pts_d, W, s.max_steps, s.eps, seed=i, device_outputs=True)
        solver.solve_raw(pts_h, W, s.max_steps, s.eps, seed=i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        r = solver.solve_raw(pts_d, W, s.max_steps, s.eps, seed=100 + i, device_outputs=True)
    e1.record(); torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / reps
    t0 = time.perf_counter()
    for i in range(reps):
        solver.solve_raw(pts_h, W, s.max_steps, s.eps, seed=100 + i)
    host_us = (time.perf_counter() - t0) * 1e6 / reps
    return dict(points=len(s.points), walks=W, device_resident_us=dev_us, host_buffers_us=host_us, steps_per_solve=int(r["steps"][0]),
                jit=nat.jit_last_note() == "")


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for key in ("cfg5", "cfg1b", "cfg1a", "cfg2", "cfg3", "cfg4"):
        for jit in ("off", "on"):
            print(key, jit, json.dumps(small_solve_us(sc.ALL[key](), reps, jit)), flush=True)
