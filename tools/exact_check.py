#!/usr/bin/env python
"""Per-walk bit comparison of the CUDA kernel with the CPU oracle on the same Philox stream (all six scenarios)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import _native as nat  # noqa: E402
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402
from oracle import wost_oracle as orc  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for key in ("cfg1a", "cfg1b", "cfg2", "cfg3", "cfg4", "cfg5"):
    s = sc.ALL[key]()
    solver = s.make_solver()
    pts = s.points[:: max(1, len(s.points) // 16)][:16].contiguous()
    r = solver.solve_raw(pts, W, s.max_steps, s.eps, seed=77, want_walk_vals=True, n_trace=len(pts) * W, trace_cap=4)
    icdf = solver._cache[("icdf", float(solver.sigma_bar), nat.current_device())].cpu().numpy() if s.delta else None
    o = orc.Problem.from_scenario(s, sigma_bar=solver.sigma_bar if s.delta else 0.0).solve(
        pts, W, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=77, icdf=icdf, walk_vals=True, walk_steps=True)
    gv, ov = np.asarray(r["walk_vals"], np.float32).ravel(), np.asarray(o["walk_vals"], np.float32).ravel()
    same = (gv.view(np.uint32) == ov.view(np.uint32)) | (np.isnan(gv) & np.isnan(ov))
    print(f"{key}: {same.mean() * 100:.4f}% of {same.size} walks bit-equal; steps gpu {int(r['steps'][0])} oracle {o['steps']}; "
          f"max |dv| {np.nanmax(np.abs(gv - ov)):.3e}", flush=True)
    bad = np.flatnonzero(~same)[:5]
    for b in bad:
        print("   walk", b, "gpu", gv[b], "oracle", ov[b], "oracle steps", o["walk_steps"].ravel()[b])
