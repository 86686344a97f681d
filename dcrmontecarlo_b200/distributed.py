"""Multi-GPU driver: one process per GPU (``torch.distributed``, NCCL over NVLink), one small collective per solve.

Every (evaluation point, walk) pair is independent (reference ``solvers/WoStSolver.py:182-187`` loops over them
sequentially), and the Philox counters are *global* (point, walk, step) indices, so the work shards freely:

* by evaluation points (electrode positions) when they divide evenly enough — rank r solves points r, r + world, ...
  with all walks (interleaved rather than contiguous slices: walk lengths depend on where a point lies, and neighbouring
  points cost about the same, so every rank gets the same mix);
* by walk ranges, on boundaries of the deterministic reduction block (``WOST_WALK_BLOCK`` walks), when there are fewer
  points than ranks or the points would balance badly (e.g. the 9-electrode DCR line on 8 GPUs) — each rank solves all
  points for a slice of the walks.

The only communication is ONE ``all_gather_into_tensor`` of the small per-point statistics (mean, M2 and the step count;
24 bytes per point, or 16 bytes per point and block), after which every rank holds the full result.  Estimates are
bit-identical for any number of ranks: per-block statistics are computed in a fixed order and merged block by block with
the same device code (``wost_merge_block_stats``).  Nothing in a solve synchronises with the host: the kernel writes its
statistics straight into the gather buffer, the step count travels in the same buffer, and the results are returned as
device tensors (``steps`` included — ``int(res["steps"])`` when the number is needed).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

WALK_BLOCK = 1024


@dataclass(frozen=True)
class Shard:
    """Points ``range(p0, p1, pstride)`` x walks ``[w0, w1)``."""
    rank: int
    p0: int
    p1: int
    w0: int
    w1: int
    pstride: int = 1

    @property
    def n_points(self):
        return len(range(self.p0, self.p1, self.pstride))

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
    """Work split for ``world`` ranks.  ``mode``: 'points', 'walks' or 'auto'.

    'auto' looks for an even split: points if they divide by the number of ranks or nearly do (the busiest rank at most
    2 % above its share: the point-sharded gather is 24 bytes per point, the walk-sharded one 16 bytes per point and
    block plus a merge); else whole reduction blocks of walks if THEY divide (and the gathered block statistics stay
    small); else points when there are >= 16 per rank (one extra point costs at most 6 %), else walks, else points."""
    if world < 1:
        raise ValueError("world must be >= 1")
    nblk = (n_walks + WALK_BLOCK - 1) // WALK_BLOCK
    if mode == "auto":
        small_gather = n_points * ((nblk + world - 1) // world) * 16 <= (8 << 20)
        busiest = (n_points + world - 1) // world
        if n_points >= world and busiest * world <= 1.02 * n_points:
            mode = "points"
        elif nblk >= world and nblk % world == 0 and small_gather:
            mode = "walks"
        elif n_points >= 16 * world:
            mode = "points"
        elif nblk >= world and small_gather:
            mode = "walks"
        else:
            mode = "points" if n_points >= world else "walks"
    if mode == "points":
        return [Shard(r, min(r, n_points), n_points, 0, n_walks, world) for r in range(world)]
    if mode == "walks":
        e = _split(nblk, world)                                        # whole reduction blocks per rank
        return [Shard(r, 0, n_points, min(e[r] * WALK_BLOCK, n_walks), min(e[r + 1] * WALK_BLOCK, n_walks)) for r in range(world)]
    raise ValueError("mode must be 'auto', 'points' or 'walks'")


def _as_tensor(a, device, dtype=None):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)


def _default_merge(block_stats: torch.Tensor, n_walks: int):
    from . import _native as nat

    return nat.merge_block_stats(block_stats, n_walks, block_stats.device.index)


class _Layout:
    """Gather buffers and index maps of one (P, nWalks, world, mode) problem shape, built once per solver."""

    def __init__(self, P, nWalks, world, rank, mode, device):
        self.plan = shard_plan(P, nWalks, world, mode)
        self.me = self.plan[rank]
        self.by_points = all(sh.w0 == 0 and sh.w1 == nWalks for sh in self.plan)
        self.nblk_total = (nWalks + WALK_BLOCK - 1) // WALK_BLOCK
        f64 = dict(dtype=torch.float64, device=device)
        if self.by_points:
            # per rank: row 0 = mean, row 1 = M2 (padded to the largest shard), row 2 = [steps, 0, ...]
            self.pmax = max(max(sh.n_points for sh in self.plan), 1)
            self.send = torch.zeros(3, self.pmax, **f64)
            self.recv = torch.empty(world, 3, self.pmax, **f64)
            where = {}
            for sh in self.plan:
                for k, p in enumerate(range(sh.p0, sh.p1, sh.pstride)):
                    where[p] = sh.rank * 3 * self.pmax + k
            self.idx_mean = torch.tensor([where[p] for p in range(P)], dtype=torch.int64, device=device)
            self.idx_m2 = self.idx_mean + self.pmax
        else:
            # per rank: (P, bmax, 2) block statistics, then one trailing slot with the step count
            self.bmax = max(max((sh.n_walks + WALK_BLOCK - 1) // WALK_BLOCK for sh in self.plan), 1)
            self.n_stats = P * self.bmax * 2
            self.send = torch.zeros(self.n_stats + 1, **f64)
            self.recv = torch.empty(world, self.n_stats + 1, **f64)
            idx = []
            for p in range(P):
                for sh in self.plan:
                    nb = (sh.n_walks + WALK_BLOCK - 1) // WALK_BLOCK
                    base = sh.rank * (self.n_stats + 1) + p * self.bmax * 2
                    idx.extend(range(base, base + 2 * nb))
            self.idx_blocks = torch.tensor(idx, dtype=torch.int64, device=device)
        self.steps_i64 = torch.zeros(1, dtype=torch.int64, device=device)


def solve_sharded(solver, points: torch.Tensor, nWalks: int, maxSteps: int = 1000, eps: float = 1e-4, *, seed=None,
                  mode: str = "auto", group=None, merge_fn=None) -> dict:
    """Collective call: every rank passes the same ``points``; returns the full ``mean`` / ``m2`` (fp64, length P), the
    total step count (a 0-d int64 tensor on the compute device, no host sync) and this rank's shard on every rank.

    Without an initialised process group this is a plain single-GPU solve with the same return value."""
    initialised = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if initialised else (0, 1)
    backend = dist.get_backend(group) if initialised else None
    on_gpu = backend == "nccl" or (not initialised and torch.cuda.is_available())
    device = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    pts = torch.as_tensor(points, dtype=torch.float32).reshape(-1, 2)
    P = pts.shape[0]
    if seed is None:
        # one Philox key for the whole job: rank 0 draws it (the host needs the number, so this path synchronises once;
        # pass `seed=` to avoid it)
        s = torch.zeros(2, dtype=torch.int64, device=device)
        if rank == 0:
            s.copy_(torch.randint(0, 1 << 32, (2,), dtype=torch.int64))
        if world > 1:
            dist.broadcast(s, src=0, group=group)
        lo, hi = s.tolist()
        seed = (hi << 32) | lo
    seed = int(seed) & ((1 << 64) - 1)                                     # all 64 key bits, like the single-GPU path

    cache = solver.__dict__.setdefault("_dist_layouts", {}) if hasattr(solver, "__dict__") else {}
    key = (P, int(nWalks), world, rank, mode, str(device))
    L = cache.get(key)
    if L is None:
        L = cache[key] = _Layout(P, int(nWalks), world, rank, mode, device)
    me = L.me
    have_work = me.n_points > 0 and me.n_walks > 0
    native_out = on_gpu and hasattr(solver, "_device_problem")           # the product solver: write straight into the gather buffer

    if L.by_points:
        if have_work:
            mine = pts[me.p0:me.p1:me.pstride].contiguous()
            if native_out:
                out = dict(mean=L.send[0, : me.n_points], m2=L.send[1, : me.n_points], steps=L.steps_i64)
                solver.solve_raw(mine, me.n_walks, maxSteps, eps, seed=seed, point_index_base=me.p0, point_index_stride=me.pstride,
                                 walk_offset=me.w0, want_block_stats=False, device_outputs=True, out=out)
                L.send[2, 0] = L.steps_i64[0]                                # int64 -> fp64 on the device (exact below 2^53)
            else:
                r = solver.solve_raw(mine, me.n_walks, maxSteps, eps, seed=seed, point_index_base=me.p0, point_index_stride=me.pstride,
                                     walk_offset=me.w0, want_block_stats=False, device_outputs=on_gpu)
                L.send[0, : me.n_points] = _as_tensor(r["mean"], device, torch.float64)
                L.send[1, : me.n_points] = _as_tensor(r["m2"], device, torch.float64)
                L.send[2, 0] = float(_as_tensor(r["steps"], device).reshape(-1)[0])
        else:
            L.send.zero_()
        if world > 1:
            dist.all_gather_into_tensor(L.recv.view(-1), L.send.view(-1), group=group)
        else:
            L.recv[0].copy_(L.send)
        flat = L.recv.reshape(-1)
        mean, m2 = flat.index_select(0, L.idx_mean), flat.index_select(0, L.idx_m2)
        steps = L.recv[:, 2, 0].sum().to(torch.int64)
    else:
        if have_work:
            nb = (me.n_walks + WALK_BLOCK - 1) // WALK_BLOCK
            if native_out and nb == L.bmax:
                out = dict(block_stats=L.send[: L.n_stats].view(P, L.bmax, 2), steps=L.steps_i64)
                solver.solve_raw(pts, me.n_walks, maxSteps, eps, seed=seed, point_index_base=0, walk_offset=me.w0,
                                 want_block_stats=True, device_outputs=True, out=out)
                L.send[L.n_stats] = L.steps_i64[0]
            else:
                r = solver.solve_raw(pts, me.n_walks, maxSteps, eps, seed=seed, point_index_base=0, walk_offset=me.w0,
                                     want_block_stats=True, device_outputs=on_gpu)
                L.send[: L.n_stats].view(P, L.bmax, 2)[:, :nb] = _as_tensor(r["block_stats"], device, torch.float64)
                L.send[L.n_stats] = float(_as_tensor(r["steps"], device).reshape(-1)[0])
        else:
            L.send.zero_()
        if world > 1:
            dist.all_gather_into_tensor(L.recv.view(-1), L.send.view(-1), group=group)
        else:
            L.recv[0].copy_(L.send)
        # per-block statistics in global block order, merged block by block — identical to the single-GPU reduction
        blocks = L.recv.reshape(-1).index_select(0, L.idx_blocks).view(P, L.nblk_total, 2)
        mean, m2 = (merge_fn or _default_merge)(blocks, int(nWalks))
        mean, m2 = _as_tensor(mean, device), _as_tensor(m2, device)
        steps = L.recv[:, L.n_stats].sum().to(torch.int64)
    return dict(mean=mean, m2=m2, steps=steps, seed=seed, shard=me, by_points=L.by_points)


def checksum(t: torch.Tensor) -> str:
    """Order-sensitive checksum of a result vector's BITS (cross-rank-count parity: must not depend on the world size)."""
    import hashlib

    return hashlib.sha256(t.detach().to("cpu").contiguous().numpy().tobytes()).hexdigest()[:16]
