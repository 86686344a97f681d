"""Multi-GPU driver: one process per GPU (``torch.distributed``, NCCL over NVLink), no data-path collective.

Every (evaluation point, walk) pair is independent (reference ``solvers/WoStSolver.py:182-187`` loops over them
sequentially), and the Philox counters are *global* (point, walk, step) indices, so the work shards freely:

* by evaluation points (electrode positions) when there are at least as many points as ranks — each rank solves a
  contiguous slice with all walks;
* by walk ranges, on boundaries of the deterministic reduction block (``WOST_WALK_BLOCK`` walks), when there are
  fewer points than ranks (e.g. the 9-electrode DCR line) — each rank solves all points for a slice of the walks.

The only communication is the gather of the small per-point statistics (16 bytes per point and block), after which
every rank holds the full result.  Estimates are bit-identical for any number of ranks: per-block statistics are
computed in a fixed order and merged block by block with the same device code (``wost_merge_block_stats``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

WALK_BLOCK = 1024


@dataclass(frozen=True)
class Shard:
    rank: int
    p0: int
    p1: int
    w0: int
    w1: int

    @property
    def n_points(self):
        return self.p1 - self.p0

    @property
    def n_walks(self):
        return self.w1 - self.w0


def _split(n: int, parts: int):
    base, rem = divmod(n, parts)
    edges = [0]
    for r in range(parts):
        edges.append(edges[-1] + base + (1 if r < rem else 0))
    return edges


def shard_plan(n_points: int, n_walks: int, world: int, mode: str = "auto") -> list[Shard]:
    """Work split for ``world`` ranks.  ``mode``: 'points', 'walks' or 'auto' (points if n_points >= world)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if mode == "auto":
        mode = "points" if n_points >= world else "walks"
    if mode == "points":
        e = _split(n_points, world)
        return [Shard(r, e[r], e[r + 1], 0, n_walks) for r in range(world)]
    if mode == "walks":
        nblk = (n_walks + WALK_BLOCK - 1) // WALK_BLOCK
        e = _split(nblk, world)                                        # whole reduction blocks per rank
        return [Shard(r, 0, n_points, min(e[r] * WALK_BLOCK, n_walks), min(e[r + 1] * WALK_BLOCK, n_walks)) for r in range(world)]
    raise ValueError("mode must be 'auto', 'points' or 'walks'")


def _as_tensor(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def _default_merge(block_stats: torch.Tensor, n_walks: int):
    from . import _native as nat

    return nat.merge_block_stats(block_stats, n_walks, block_stats.device.index)


def solve_sharded(solver, points: torch.Tensor, nWalks: int, maxSteps: int = 1000, eps: float = 1e-4, *, seed=None,
                  mode: str = "auto", group=None, merge_fn=None) -> dict:
    """Collective call: every rank passes the same ``points``; returns the full ``mean`` / ``m2`` (fp64, length P),
    the total step count and this rank's shard on every rank."""
    if not dist.is_initialized():
        raise RuntimeError("solve_sharded needs an initialised torch.distributed process group")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pts = torch.as_tensor(points, dtype=torch.float32).reshape(-1, 2)
    P = pts.shape[0]
    # one Philox key for the whole job
    s = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        if seed is None:
            hi, lo = torch.randint(0, 1 << 31, (2,), dtype=torch.int64).tolist()
            seed = (hi << 31) | lo
        s[0] = int(seed) & ((1 << 63) - 1)
    dist.broadcast(s, src=0, group=group)
    seed = int(s.item())

    plan = shard_plan(P, nWalks, world, mode)
    me = plan[rank]
    by_points = all(sh.w0 == 0 and sh.w1 == nWalks for sh in plan)
    nblk_total = (nWalks + WALK_BLOCK - 1) // WALK_BLOCK
    local = None
    if me.n_points > 0 and me.n_walks > 0:
        local = solver.solve_raw(pts[me.p0:me.p1], me.n_walks, maxSteps, eps, seed=seed, point_index_base=me.p0,
                                 walk_offset=me.w0, want_block_stats=not by_points, device_outputs=(device.type == "cuda"))
    steps = torch.zeros(1, dtype=torch.int64, device=device)
    if local is not None:
        steps += _as_tensor(local["steps"], device).to(torch.int64).reshape(-1)[:1]
    dist.all_reduce(steps, group=group)

    if by_points:
        # gather (mean, m2) slices, padded to the largest shard
        pmax = max(sh.n_points for sh in plan)
        buf = torch.zeros(pmax, 2, dtype=torch.float64, device=device)
        if local is not None:
            buf[: me.n_points, 0] = _as_tensor(local["mean"], device)
            buf[: me.n_points, 1] = _as_tensor(local["m2"], device)
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf, group=group)
        full = torch.cat([out[sh.rank][: sh.n_points] for sh in plan], dim=0)
        mean, m2 = full[:, 0].contiguous(), full[:, 1].contiguous()
    else:
        # gather per-block statistics and merge them in block order — identical to the single-GPU reduction
        bmax = max((sh.n_walks + WALK_BLOCK - 1) // WALK_BLOCK for sh in plan)
        buf = torch.zeros(P, bmax, 2, dtype=torch.float64, device=device)
        if local is not None:
            b = _as_tensor(local["block_stats"], device)
            buf[:, : b.shape[1]] = b
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf, group=group)
        blocks = torch.cat([out[sh.rank][:, : (sh.n_walks + WALK_BLOCK - 1) // WALK_BLOCK] for sh in plan], dim=1).contiguous()
        assert blocks.shape[1] == nblk_total
        mean, m2 = (merge_fn or _default_merge)(blocks, nWalks)
        mean, m2 = _as_tensor(mean, device), _as_tensor(m2, device)
    return dict(mean=mean, m2=m2, steps=int(steps.item()), seed=seed, shard=me, by_points=by_points)
