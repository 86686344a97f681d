#!/usr/bin/env python
"""Run a few passes of one scenario on the GPU (profiling target): python tools/run_one.py cfg5_175e [passes]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402

name = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
walks_override = int(sys.argv[3]) if len(sys.argv) > 3 else 0
table = {"cfg1a": (sc.cfg1a, 65536, 256), "cfg1b": (sc.cfg1b, 65536, 64), "cfg2": (sc.cfg2, 65536, 256), "cfg3": (sc.cfg3, 65536, 256),
         "cfg4": (sc.cfg4, 65536, 64), "cfg5_9e": (lambda: sc.cfg5(9), 9, 32768), "cfg5_175e": (lambda: sc.cfg5(175), 175, 4096),
         "cfg5": (lambda: sc.cfg5(175), 175, 32768)}          # cfg5 = bench.py's headline job
mk, n, w = table[name]
w = walks_override or w
s = mk()
pts = s.points.repeat((n + len(s.points) - 1) // len(s.points), 1)[:n].contiguous().cuda()
solver = s.make_solver()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(passes):
    e0.record()
    r = solver.solve_raw(pts, w, s.max_steps, s.eps, seed=i, device_outputs=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "pass", i, "ms", round(e0.elapsed_time(e1), 3), "steps/s %.3e" % (int(r["steps"][0]) / e0.elapsed_time(e1) * 1e3), "steps", int(r["steps"][0]))
