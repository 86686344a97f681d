"""DC-resistivity survey driver: many current-electrode pairs over one line of measurement electrodes.

The reference solves one source pair at a time by editing a script (``tests/testGeophysicalScenario.py:11-33,109-151``;
the notebook variant ``tests/testNotebook.ipynb`` cells 3, 17-21 builds a dipole-dipole line and differences
``V_M - V_N`` between neighbouring electrodes).  Here a survey is a list of source dipoles sharing the geometry and the
conductivity field; each source is one launch of the walk kernel over all measurement electrodes, and sources are
sharded over GPUs (``torch.distributed``) when a process group is initialised.

Current electrodes are Gaussian blobs like the reference's (``norm = I / (2 pi w^2)``, ``exp(-d^2 / (2 w^2))``).
``sink_sign=-1`` gives a physical dipole (+I at A, -I at B, as in the notebook); ``sink_sign=+1`` reproduces
``testGeophysicalScenario.py:29-33``, where the "sink" enters with a positive sign.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from .fields import Field, TermField
from .geometry.Polylines import PolyLines
from .solvers.WoStSolver import WostSolver_2D


@dataclass(frozen=True)
class DipoleSource:
    a: tuple            # (x, y) of the +I electrode
    b: tuple            # (x, y) of the -I electrode
    current: float = 1.0
    width: float = 0.5  # Gaussian width of the injection (reference: sigma = 0.5 m)

    def field(self, sink_sign: float = -1.0) -> TermField:
        norm = self.current / (2.0 * math.pi * self.width ** 2)
        q = 1.0 / (2.0 * self.width ** 2)
        return TermField.gaussian_sum([(norm, tuple(map(float, self.a)), q), (sink_sign * norm, tuple(map(float, self.b)), q)])


def geometric_factor_2d(a, b, m, n) -> float:
    """Geometric factor K of a four-electrode array on a homogeneous 2D half-plane (line electrodes):
    V(r) = -(rho I / pi) ln r  =>  rho_a = K * (V_M - V_N) / I  with  K = pi / (ln(BM/AM) - ln(BN/AN))."""
    d = lambda p, q: math.hypot(p[0] - q[0], p[1] - q[1])                # noqa: E731
    dists = (d(a, m), d(b, m), d(a, n), d(b, n))
    if min(dists) == 0.0:                                                # a potential electrode sits on a current electrode
        return float("nan")
    den = math.log(dists[1] / dists[0]) - math.log(dists[3] / dists[2])
    return math.pi / den if den != 0.0 else float("nan")


class DCRSurvey:
    def __init__(self, dirichletBoundary: PolyLines, neumannBoundary: PolyLines, conductivity: Field,
                 electrodes: torch.Tensor, sources: Sequence[DipoleSource], receivers: Sequence[tuple] | None = None,
                 sink_sign: float = -1.0, compat: str = "reference", **solver_kw):
        """``compat="physical"`` runs the survey with the textbook estimator (reflections that do not leak, delta
        tracking with a spatially varying majorant) instead of the reference's; further keyword arguments go to
        ``WostSolver_2D``."""
        self.electrodes = torch.as_tensor(electrodes, dtype=torch.float32).reshape(-1, 2).contiguous()
        self.sources = list(sources)
        E = self.electrodes.shape[0]
        # receiver dipoles (M, N) as electrode indices; default: neighbouring electrodes (notebook cell 3)
        self.receivers = [(i, i + 1) for i in range(E - 1)] if receivers is None else [tuple(r) for r in receivers]
        self.sink_sign = float(sink_sign)
        # one solver: sigma' and sigma_bar depend on the conductivity only (no absorption in DC resistivity)
        self.solver = WostSolver_2D(dirichletBoundary, None, neumannBoundary, source=None, sigma=None, alpha=conductivity,
                                    compat=compat, **solver_kw)
        self._fields = [s.field(self.sink_sign) for s in self.sources]
        self._streams: list = []

    def run(self, nWalks: int = 1000, maxSteps: int = 500, eps: float = 0.9, seed: int | None = None, streams: int = 8,
            shared_walks: bool = False) -> dict:
        """Potentials at every electrode for every source.  Returns
        ``potentials`` (S, E) float64, ``stderr`` (S, E), ``dV`` (S, R) = V_M - V_N per receiver dipole, ``steps``.
        With an initialised ``torch.distributed`` group the work is sharded over the ranks -- the sources round-robin, or
        with ``shared_walks`` the electrodes -- and the result is gathered on every rank; the same ``seed`` gives the same
        numbers for any number of ranks.

        ``shared_walks=True``: the walk does not depend on the source term, so ONE set of walks per electrode serves
        all sources (``wost_solve_multi_source``) — the cost of a survey becomes almost independent of the number of
        source pairs.  The estimates of different sources are then correlated (same paths), which is what one wants
        for differences between configurations; each is still an unbiased estimate of its own potential."""
        import torch.distributed as dist

        world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
        on_nccl = world > 1 and dist.get_backend() == "nccl"
        if seed is None:
            hi, lo = torch.randint(0, 1 << 31, (2,), dtype=torch.int64).tolist()
            seed = (hi << 31) | lo
            if world > 1:                                                 # one key for the whole job
                t = torch.tensor([seed if rank == 0 else 0], dtype=torch.int64, device="cuda" if on_nccl else "cpu")
                dist.broadcast(t, src=0)
                seed = int(t.item())
        S, E = len(self.sources), self.electrodes.shape[0]
        if shared_walks:
            return self._run_shared(nWalks, maxSteps, eps, seed, world, rank, on_nccl)
        pot, m2 = np.zeros((S, E)), np.zeros((S, E))
        steps = 0
        # One kernel launch per source over all electrodes.  A launch of a few electrodes does not fill the GPU (the
        # persistent grid is sized to the work), so sources are issued round-robin on several streams with
        # device-resident results and collected once at the end.
        mine = list(range(rank, S, world))
        n_streams = max(1, min(int(streams), len(mine)))
        if n_streams > 1:
            # streams are kept: torch's caching allocator pools memory per stream, fresh streams would re-allocate
            while len(self._streams) < n_streams:
                self._streams.append(torch.cuda.Stream())
            pool = self._streams[:n_streams]
            for st in pool:
                st.wait_stream(torch.cuda.current_stream())
        else:
            pool = [torch.cuda.current_stream()]
        el_dev = self.electrodes.cuda()
        pending = []
        for i, s in enumerate(mine):
            self.solver.setSourceTerm(self._fields[s])
            with torch.cuda.stream(pool[i % n_streams]):
                # the source index goes into the Philox key, so every source walks its own paths
                r = self.solver.solve_raw(el_dev, nWalks, maxSteps, eps, seed=(seed + 0x9E3779B97F4A7C15 * (s + 1)) % (1 << 64),
                                          device_outputs=True)
            pending.append((s, r))
        for st in pool:
            st.synchronize()
        for s, r in pending:
            pot[s], m2[s] = r["mean"].cpu().numpy(), r["m2"].cpu().numpy()
            steps += int(r["steps"][0])
        if world > 1:
            dev = "cuda" if on_nccl else "cpu"
            buf = torch.from_numpy(np.stack([pot, m2])).to(dev)
            dist.all_reduce(buf)                                          # disjoint rows: the sum is a gather
            st = torch.tensor([steps], dtype=torch.int64, device=dev)
            dist.all_reduce(st)
            pot, m2, steps = buf[0].cpu().numpy(), buf[1].cpu().numpy(), int(st.item())
        return self._finish(pot, m2, steps, seed, nWalks)

    def _run_shared(self, nWalks, maxSteps, eps, seed, world, rank, on_nccl):
        """Shared walks: ONE set of walks per electrode serves every source, so the ELECTRODES are what is sharded over
        the ranks (each rank: its slice of the electrode line x all sources x all walks, one launch), and one
        all-gather of the (source, electrode) statistics leaves the full result on every rank.  Philox counters carry the
        global electrode index, so the numbers do not depend on the number of ranks."""
        import torch.distributed as dist

        S, E = len(self.sources), self.electrodes.shape[0]
        # rank r takes electrodes r, r + world, ...: the cost of an electrode depends on how many current electrodes its
        # walks pass, which varies along the line, so interleaved subsets balance where contiguous chunks would not
        mine = self.electrodes[rank::world].contiguous()
        n_mine, emax = mine.shape[0], (E + world - 1) // world
        use_cuda = torch.cuda.is_available() and hasattr(self.solver, "_device_problem")
        dev = torch.device("cuda", torch.cuda.current_device()) if (on_nccl or (world == 1 and use_cuda)) else torch.device("cpu")
        send = torch.zeros(2 * S * emax + 1, dtype=torch.float64, device=dev)
        if n_mine > 0:
            r = self.solver.solve_multi_source(mine, self._fields, nWalks, maxSteps, eps, seed=seed, point_index_base=rank,
                                               point_index_stride=world, device_outputs=dev.type == "cuda")
            as_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev)   # noqa: E731
            blk = send[: 2 * S * emax].view(2, S, emax)
            blk[0, :, :n_mine], blk[1, :, :n_mine] = as_t(r["mean"]), as_t(r["m2"])
            send[2 * S * emax] = as_t(r["steps"]).reshape(-1)[0].to(torch.float64)
        if world > 1:
            recv = torch.empty(world, 2 * S * emax + 1, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(recv.view(-1), send, group=None)
        else:
            recv = send.view(1, -1)
        # (world, 2, S, emax) -> electrode e = slot * world + rank
        full = recv[:, : 2 * S * emax].view(world, 2, S, emax).permute(1, 2, 3, 0).reshape(2, S, emax * world)[:, :, :E].cpu().numpy()
        steps = int(recv[:, 2 * S * emax].sum().item())
        return self._finish(full[0], full[1], steps, seed, nWalks)

    def _finish(self, pot, m2, steps, seed, nWalks):
        n = float(nWalks)
        stderr = np.sqrt(m2 / max(n - 1.0, 1.0) / n)
        M = np.array([r[0] for r in self.receivers], dtype=np.int64); N = np.array([r[1] for r in self.receivers], dtype=np.int64)
        return dict(potentials=pot, stderr=stderr, dV=pot[:, M] - pot[:, N], dV_stderr=np.hypot(stderr[:, M], stderr[:, N]),
                    steps=steps, seed=seed)

    def apparent_resistivity(self, dV: np.ndarray) -> np.ndarray:
        """rho_a (S, R) from the receiver voltages with the 2D half-plane geometric factor of each (A, B, M, N)."""
        out = np.zeros_like(dV)
        el = self.electrodes.numpy()
        for s, src in enumerate(self.sources):
            for k, (m, n) in enumerate(self.receivers):
                out[s, k] = geometric_factor_2d(src.a, src.b, el[m], el[n]) * dV[s, k] / src.current
        return out
