#!/usr/bin/env python
"""DCR survey throughput: per-source launches (1 stream / 8 streams) vs shared walks.  python tools/survey_bench.py"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402
from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple  # noqa: E402
from dcrmontecarlo_b200.survey import DCRSurvey, DipoleSource  # noqa: E402

for n_el, n_src, W in ((21, 32, 4096), (175, 64, 1024)):
    s = sc.cfg5(n_el)
    srcs = [DipoleSource((float(x), 0.0), (float(x) + 20.0, 0.0)) for x in np.linspace(-40, 20, n_src)]
    sv = DCRSurvey(PolyLinesSimple(s.dirichlet), PolyLinesSimple(s.neumann), s.alpha, s.points, srcs)
    sv.run(nWalks=128, seed=1); sv.run(nWalks=128, seed=1, shared_walks=True)
    for label, kw in (("1 stream", dict(streams=1)), ("8 streams", dict(streams=8)), ("shared walks", dict(shared_walks=True))):
        dt = 1e9
        for rep in range(3):                                               # best of 3: the first run at a size grows the memory pool
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = sv.run(nWalks=W, maxSteps=500, eps=0.9, seed=2, **kw)
            dt = min(dt, time.perf_counter() - t0)
        print(f"{n_el} electrodes x {n_src} sources x {W} walks, {label:12s}: {dt * 1e3:8.2f} ms   walk-steps taken {out['steps']:.3e}"
              f"   source-electrode estimates/s {n_el * n_src / dt:.3e}", flush=True)
