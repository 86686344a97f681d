#!/usr/bin/env python
"""The per-rank share of bench.py's survey strong-scaling job on ONE GPU: every `world`-th of 175 electrodes x 64 source
dipoles x W shared walks (what one of `world` ranks computes), to see how the single-rank time falls with the share.

    python tools/survey_rank_job.py [W]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402
from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple  # noqa: E402
from dcrmontecarlo_b200.survey import DCRSurvey, DipoleSource  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
c5 = sc.cfg5(175)
srcs = [DipoleSource((-38.0 + 1.1 * k, 0.0), (38.0 - 1.1 * k, 0.0)) for k in range(64)]
t1 = None
for world in (1, 2, 4, 8):
    pts = c5.points[0::world].contiguous()
    sv = DCRSurvey(PolyLinesSimple(c5.dirichlet), PolyLinesSimple(c5.neumann), c5.alpha, pts, srcs, sink_sign=+1.0)
    for i in range(2):
        sv.run(nWalks=W, maxSteps=c5.max_steps, eps=c5.eps, seed=99, shared_walks=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        res = sv.run(nWalks=W, maxSteps=c5.max_steps, eps=c5.eps, seed=99, shared_walks=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    t1 = t1 or ms
    print(f"share 1/{world}: {len(pts)} electrodes x 64 sources x {W} walks: {ms:.3f} ms  ({res['steps']:.3e} steps)  speed-up over the whole job {t1 / ms:.2f}", flush=True)
