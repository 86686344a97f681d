#!/usr/bin/env python
"""Run a few passes of a scale scene (profiling target): python tools/run_scale.py N_SEG [points] [walks] [passes]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402

n = int(sys.argv[1]); P = int(sys.argv[2]) if len(sys.argv) > 2 else 65536; W = int(sys.argv[3]) if len(sys.argv) > 3 else 16
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = sc.scale_scene(n, P, W); solver = s.make_solver(); pts = s.points.cuda()
for i in range(passes):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = solver.solve_raw(pts, W, s.max_steps, s.eps, seed=i, device_outputs=True); e1.record(); torch.cuda.synchronize()
    print(n, "pass", i, "ms %.3f" % e0.elapsed_time(e1), "steps/s %.3e" % (int(r["steps"][0]) / e0.elapsed_time(e1) * 1e3), flush=True)
