#!/usr/bin/env python
"""Solve time against job size for one scenario (device-resident, CUDA events): python tools/midsize.py cfg3 [WOST_TIMING]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import distributed as dm  # noqa: E402
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
s = sc.ALL[key]()
solver = s.make_solver()
pts = s.points.cuda()
for W in (150, 600, 2400, 9600, 38400):
    for i in range(3):
        solver.solve_raw(pts, W, s.max_steps, s.eps, seed=i, device_outputs=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        r = solver.solve_raw(pts, W, s.max_steps, s.eps, seed=10 + i, device_outputs=True)
    e1.record(); torch.cuda.synchronize()
    raw_us = e0.elapsed_time(e1) * 100.0
    steps = int(r["steps"][0])
    for i in range(3):
        dm.solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=i)["mean"].cpu()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10):
        dm.solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=10 + i)["mean"].cpu()
    sh_us = (time.perf_counter() - t0) * 1e5
    print(f"{key} {len(pts)} x {W}: raw solve {raw_us:8.1f} us  ({steps / raw_us * 1e6:.3e} steps/s)   solve_sharded + D2H {sh_us:8.1f} us", flush=True)
