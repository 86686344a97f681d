"""Extendable / resumable estimates: add walks to an existing solve, checkpoint in between.

Not in the reference (no checkpoint / resume).  It falls out of the design: Philox counters are global
(point, walk, step) indices and the statistics are reduced per block of ``WALK_BLOCK`` walks in a fixed order, so walks
``[0, n)`` followed later by walks ``[n, m)`` give bit for bit the estimate of one solve with ``m`` walks, as long as
``n`` is a multiple of the block size.  The state is just the per-block ``(mean, M2)`` table.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat

WALK_BLOCK = nat.WALK_BLOCK


class RunningEstimate:
    def __init__(self, solver, points, maxSteps: int = 1000, eps: float = 1e-4, seed: int = 0):
        self.solver, self.maxSteps, self.eps, self.seed = solver, int(maxSteps), float(eps), int(seed)
        self.points = torch.as_tensor(points, dtype=torch.float32).reshape(-1, 2).contiguous()
        self.n_walks = 0
        self.steps = 0
        self.blocks = np.zeros((self.points.shape[0], 0, 2))

    def add_walks(self, n: int) -> "RunningEstimate":
        """Run walks [n_walks, n_walks + n).  All but the last call must add a multiple of WALK_BLOCK walks."""
        if self.n_walks % WALK_BLOCK:
            raise ValueError(f"walks can only be appended on a {WALK_BLOCK}-walk boundary (have {self.n_walks})")
        r = self.solver.solve_raw(self.points, int(n), self.maxSteps, self.eps, seed=self.seed, walk_offset=self.n_walks,
                                  want_block_stats=True)
        self.blocks = np.concatenate([self.blocks, r["block_stats"]], axis=1)
        self.n_walks += int(n)
        self.steps += int(r["steps"][0])
        return self

    def estimate(self):
        """(mean (P,), standard error (P,)) over all walks so far, merged by the solver's own fixed-order merge kernel."""
        if self.n_walks == 0:
            raise ValueError("no walks yet")
        mean, m2 = nat.merge_block_stats(self.blocks, self.n_walks, nat.current_device())
        n = float(self.n_walks)
        return mean, np.sqrt(m2 / max(n - 1.0, 1.0) / n)

    # checkpoint / resume ----------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return dict(points=self.points.numpy().copy(), blocks=self.blocks.copy(), n_walks=self.n_walks, steps=self.steps,
                    seed=self.seed, maxSteps=self.maxSteps, eps=self.eps)

    @classmethod
    def from_state_dict(cls, solver, state: dict) -> "RunningEstimate":
        self = cls(solver, state["points"], state["maxSteps"], state["eps"], state["seed"])
        self.blocks, self.n_walks, self.steps = np.array(state["blocks"]), int(state["n_walks"]), int(state["steps"])
        return self
