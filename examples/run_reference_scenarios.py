#!/usr/bin/env python
"""The reference's driver scripts (tests/testWoStCorrectness.py, testWostWithSource.py, testWostVariableCoefficients.py,
testGeophysicalScenario.py) as one command on the B200 solver.

    python examples/run_reference_scenarios.py cfg1b --walks 10 25 50 150
    python examples/run_reference_scenarios.py cfg1b --walks 1000 100000 --compat physical
    python examples/run_reference_scenarios.py cfg5 --walks 100 10000

For each walk count it prints what the reference scripts print -- mean / max / RMSE against the analytic solution where
one exists (cfg 1a, 1b, 3), the estimates otherwise -- plus wall time and walk-steps/s.  `--compat physical` runs the
textbook estimator instead of the reference's (for cfg 1b this removes the reference's bias floor of 0.028; the scenarios
with Neumann boundaries differ by design).  Needs a GPU; there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("scenario", choices=sorted(sc.ALL))
    ap.add_argument("--walks", type=int, nargs="+", default=[10, 25, 50, 150])
    ap.add_argument("--compat", choices=["reference", "physical"], default="reference")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()

    s = sc.ALL[args.scenario]()
    s.compat = args.compat
    if args.compat == "physical":
        s.sigma_bar = None
    torch.manual_seed(args.seed)                                          # the Philox key is drawn from torch's generator
    solver = s.make_solver()
    exact = s.analytic(s.points).double() if s.analytic is not None else None
    print(f"{s.name}: {len(s.points)} points, maxSteps {s.max_steps}, eps {s.eps}, compat {args.compat}"
          + (f", delta tracking with sigma_bar = {solver.sigma_bar:.5g}" if solver.use_delta_tracking else ""))
    for W in args.walks:
        t0 = time.perf_counter()
        est, stats = solver.solve(s.points, nWalks=W, maxSteps=s.max_steps, eps=s.eps, return_stats=True)
        dt = time.perf_counter() - t0
        u = est[:, 0].double()
        line = f"  nWalks {W:>8}: {dt * 1e3:9.2f} ms  {stats['total_steps'] / dt:10.3e} walk-steps/s  mean stderr {stats['stderr'].mean():.4g}"
        if exact is not None:
            err = (u - exact).abs()
            line += f"  mean|err| {err.mean():.5f}  max|err| {err.max():.5f}  RMSE {torch.sqrt((err ** 2).mean()):.5f}"
        else:
            line += "  u = " + " ".join(f"{v:.4g}" for v in u[:9].tolist())
        print(line)


if __name__ == "__main__":
    main()
