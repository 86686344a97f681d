"""``WostSolver_2D`` — the reference's solver class (``solvers/WoStSolver.py:15-353``) on a B200.

Same constructor, setters and ``solve`` signature as the reference; the per-point / per-walk / per-step
Python loop (``_solveUnified``, reference ``:162-316``) is replaced by one persistent CUDA kernel with one
thread per walk (``csrc/wost_lib.cu``), called through the C ABI in ``include/wost.h``.  Setup that runs once
per solver (sigma', sigma_bar — reference ``:66-138``) stays on the host in PyTorch.

The callables ``dirichletBoundaryFunction, source, sigma, alpha`` may be :mod:`fields` objects (evaluated
analytically on the device) or arbitrary Python callables on a ``(2,)`` tensor, which are tabulated on a
lattice over the domain once and interpolated bilinearly by the kernel.

Extra, keyword-only arguments beyond the reference's: ``seed`` (Philox key; default derived from
``torch.initial_seed()`` so ``torch.manual_seed(42)`` still makes a script reproducible), ``return_stats``.
There is no CPU fallback: without the CUDA library or a GPU, ``solve`` raises.
"""
from __future__ import annotations

import numpy as np
import torch

try:
    from ..geometry.Polylines import PolyLines
    from ..utils import torchGradient, torchLaplacian, gridSampleMinMax
    from .. import _native as nat
    from ..fields import Field, TermField, as_field
    from .utils import screened_radius_icdf
except ImportError:  # pragma: no cover - reference-style sys.path layout (solvers.WoStSolver)
    from geometry.Polylines import PolyLines
    from utils import torchGradient, torchLaplacian, gridSampleMinMax
    import _native as nat
    from fields import Field, TermField, as_field
    from solvers.utils import screened_radius_icdf

SP_FULL, SP_RATIO, SP_FIELD = nat.SP_FULL, nat.SP_RATIO, nat.SP_FIELD

def _next_seed() -> int:
    """A fresh Philox key per solve(), drawn from torch's global CPU generator — the stream the reference consumes
    with torch.rand (solvers/WoStSolver.py:226) — so torch.manual_seed(42) replays a script and successive solves differ."""
    hi, lo = torch.randint(0, 1 << 31, (2,), dtype=torch.int64).tolist()
    return (hi << 31) | lo


def _zero_boundary(point):
    return 0.0


class WostSolver_2D:
    """Walk-on-Stars solver for ``-div(alpha grad u) + sigma u = f`` with Dirichlet and zero-Neumann polyline
    boundaries in 2D (reference class docstring, ``solvers/WoStSolver.py:15-20``)."""

    def __init__(self, dirichletBoundary: PolyLines, dirichletBoundaryFunction: callable = None,
                 neumannBoundary: PolyLines = None, source: callable = None, sigma: callable = None,
                 alpha: callable = None, *, field_resolution: int = 257, sigma_prime_resolution: int = 65,
                 sigma_prime_mode: str = "auto", compat: str = "reference", field_tolerance: float | None = 1e-3,
                 majorant_resolution: int = 256, jit: str = "auto"):
        """``sigma_prime_mode``: ``"auto"`` differentiates the coefficients like the reference and falls back to
        ``sigma/alpha`` when that fails; ``"ratio"`` forces the fallback — what the reference ends up with for
        callables that wrap their result in ``torch.tensor(...)`` (tests/testWostVariableCoefficients.py:49,57,
        SURVEY Q12) — and ``"full"`` insists on the differentiated form.

        ``compat``: ``"reference"`` (default) reproduces the reference's estimator, quirks included;
        ``"physical"`` runs textbook Walk on Stars — first hit by ray distance, reflection into the hemisphere facing the
        domain, the closing vertex of a closed loop is a silhouette candidate, termination projects onto the Dirichlet
        boundary, Green's-function source sampling with visibility; variable ``alpha`` / ``sigma`` by delta tracking
        with the screened kernel's proper weights and a true majorant (Neumann walls then need ``d alpha/dn = 0``).  It is
        not part of the reference (whose mixed-boundary walks leak through Neumann walls, Q1/Q2, and whose delta-tracking
        estimator is biased, Q8/Q9/Q13) and is validated against analytic solutions instead."""
        if compat not in nat.COMPAT:
            raise ValueError("compat must be 'reference' or 'physical'")
        if jit not in nat.JIT:
            raise ValueError("jit must be 'auto', 'on' or 'off'")
        # "auto": jobs of >= 2^18 walks run a kernel compiled for this solver's own fields (NVRTC, ~0.3 s once per field
        # set, cached); "on" / "off" force / forbid it.  Bit-identical results either way.
        self.jit = jit
        self.compat = compat
        # physical mode with variable coefficients: cells per side of the spatially varying majorant (a power of two;
        # 0 = one majorant for the whole domain)
        self.majorant_resolution = int(majorant_resolution)
        if self.majorant_resolution and (self.majorant_resolution & (self.majorant_resolution - 1) or self.majorant_resolution > 4096):
            raise ValueError("majorant_resolution must be 0 or a power of two <= 4096")
        self.majorant = None
        self.field_tolerance = field_tolerance       # callables that must be tabulated: refine the table to this relative error
        if sigma_prime_mode not in ("auto", "ratio", "full"):
            raise ValueError("sigma_prime_mode must be 'auto', 'ratio' or 'full'")
        self.sigma_prime_mode = sigma_prime_mode
        self.dirichletBoundary = dirichletBoundary
        self.neumannBoundary = neumannBoundary
        self.field_resolution = int(field_resolution)
        self.sigma_prime_resolution = int(sigma_prime_resolution)

        # bounding box over both boundaries (reference :38-43)
        pts = [dirichletBoundary.points] + ([neumannBoundary.points] if neumannBoundary else [])
        allp = torch.cat([torch.as_tensor(p, dtype=torch.float32).cpu() for p in pts], dim=0)
        self.domain_bounds = [[allp[:, 0].min(), allp[:, 0].max()], [allp[:, 1].min(), allp[:, 1].max()]]

        self.boundaryDirichlet = _zero_boundary if dirichletBoundaryFunction is None else dirichletBoundaryFunction
        self.source = source
        self.use_delta_tracking = False
        self.sp_mode = SP_FULL
        self.last_stats = None
        self._cache: dict = {}
        # bound methods are new objects on every attribute access: bind the sigma' tabulation callable ONCE so that the
        # id()-keyed field cache below hits on the second solve
        self._sp_plain = self._sigma_prime_plain
        self._probe_pts = None

        if sigma is not None or alpha is not None:                      # reference :54-64
            self.sigma = (lambda point: 0.0) if sigma is None else sigma
            self.alpha = (lambda point: 1.0) if alpha is None else alpha
            self._sigma_given, self._alpha_given = sigma is not None, alpha is not None
            self.sigma_prime, self.sigma_bar = self.buildModifiedSigma()
            self.use_delta_tracking = True
            if compat == "physical":
                self.sigma_bar = self._physical_majorant()
                self.use_delta_tracking = self.sigma_bar > 0.0           # constant alpha, no absorption: plain WoSt

    # ------------------------------------------------------------------------------------------------
    # setup (host): sigma' and sigma_bar, reference :66-138
    # ------------------------------------------------------------------------------------------------
    def _wrapped(self):
        def sigma_w(p):
            v = self.sigma(p)
            return v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=torch.float32, requires_grad=True)

        def alpha_w(p):
            v = self.alpha(p)
            if not isinstance(v, torch.Tensor):
                v = torch.tensor(v, dtype=torch.float32, requires_grad=True)
            return torch.clamp(v, min=1e-8)                              # :84-86

        return sigma_w, alpha_w

    def buildModifiedSigma(self):
        """Returns ``(sigma_prime, sigma_bar)``: the delta-tracking absorption
        ``sigma/alpha + (lap(alpha)/alpha - |grad ln alpha|^2 / 2) / 2`` (falling back to ``sigma/alpha`` when the
        callables cannot be differentiated, as the reference does) and its range over a 50x50 lattice on the
        bounding box, replaced by 10.0 when not in (0, 1e3] (reference :66-138)."""
        sigma_w, alpha_w = self._wrapped()
        self._autograd_failed = False

        def sigma_prime(point):
            point = point.clone().requires_grad_(True) if not point.requires_grad else point.clone()
            ratio = sigma_w(point) / alpha_w(point)
            if self.sigma_prime_mode == "ratio":
                return ratio
            try:
                lap = torchLaplacian(alpha_w, point)
                g = torchGradient(lambda p: torch.log(alpha_w(p) + 1e-8), point)
                return ratio + 0.5 * (lap / alpha_w(point) - (g ** 2).sum() / 2.0)
            except Exception:
                if self.sigma_prime_mode == "full":
                    raise
                self._autograd_failed = True
                return ratio

        lo, hi = self._sigma_prime_range(sigma_prime)
        sigma_bar = hi - lo
        if (sigma_bar <= 0) | (sigma_bar > 1e3):                         # :134-136
            sigma_bar = 10.0
        # which device formulation reproduces this closure
        # (plain callables that are closed-form expressions are traced into exact TermFields, see fieldtrace.py)
        analytic = all(not given or isinstance(self._host_field(c), TermField)
                       for c, given in ((self.alpha, self._alpha_given), (self.sigma, self._sigma_given)))
        if self._autograd_failed or self.sigma_prime_mode == "ratio":
            self.sp_mode = SP_RATIO
        elif analytic:
            self.sp_mode = SP_FULL
        else:
            self.sp_mode = SP_FIELD
        return sigma_prime, sigma_bar

    def _sigma_prime_range(self, sigma_prime):
        """min / max of sigma' on the 50x50 lattice (reference :130, utils.py:65-120).  Field coefficients are
        evaluated in one vectorised autograd pass (same elementwise arithmetic), anything else point by point."""
        fields_only = all(isinstance(c, Field) or not given
                          for c, given in ((self.alpha, self._alpha_given), (self.sigma, self._sigma_given)))
        if fields_only:
            vals = self._sigma_prime_lattice_vectorised(50)
            if vals is not None:
                vals = vals[torch.isfinite(vals)]
                if vals.numel() == 0:
                    raise ValueError("Function could not be evaluated at any grid points")
                return vals.min().item(), vals.max().item()
        lo, hi, _, _ = gridSampleMinMax(sigma_prime, self.domain_bounds, grid_resolution=50)
        return lo, hi

    def _physical_majorant(self, n: int = 129) -> float:
        """``compat="physical"``: delta tracking needs a true majorant ``sigma_bar >= |sigma'|`` (the reference takes the
        *range* of sigma' on a 50x50 lattice and 10.0 when that looks odd, ``:130-136``).  Here: 1.05 x the
        largest ``|sigma'|`` on an n x n lattice over the bounding box, with sigma' in its differentiated form."""
        import warnings

        hosts = [(self._host_field(c) if given else None) for c, given in ((self.alpha, self._alpha_given), (self.sigma, self._sigma_given))]
        N = self.majorant_resolution
        if N and all(h is None or isinstance(h, TermField) for h in hosts):
            n = 2 * N + 1                                                 # two lattice intervals per majorant cell
        if all(h is None or isinstance(h, TermField) for h in hosts):
            if self.sigma_prime_mode != "ratio":
                self.sp_mode = SP_FULL if self._alpha_given else SP_RATIO    # closed form on the device, whatever autograd said
            keep = (self.alpha, self.sigma)
            try:
                self.alpha, self.sigma = (hosts[0] if self._alpha_given else self.alpha), (hosts[1] if self._sigma_given else self.sigma)
                vals = self._sigma_prime_lattice_vectorised(n)
            finally:
                self.alpha, self.sigma = keep
        else:
            if self._autograd_failed and self._alpha_given and self.sigma_prime_mode != "ratio":
                warnings.warn("compat='physical': alpha could not be differentiated, sigma' falls back to sigma/alpha "
                              "(exact only where alpha is constant)", RuntimeWarning)
            lo, hi, _, _ = gridSampleMinMax(self.sigma_prime, self.domain_bounds, grid_resolution=min(n, 65))
            vals = torch.tensor([lo, hi])
        if N and vals.numel() == (2 * N + 1) ** 2:
            self.majorant = self._majorant_pyramid(vals.reshape(2 * N + 1, 2 * N + 1), N)
        vals = vals[torch.isfinite(vals)]
        if vals.numel() == 0:
            raise ValueError("sigma' could not be evaluated anywhere on the domain lattice")
        return 1.05 * float(vals.abs().max())

    def _majorant_pyramid(self, lattice: torch.Tensor, N: int) -> dict:
        """Max-pyramid of |sigma'| (include/wost.h, ``wost_solve_params_t.majorant``): cell (i, j) of level 0 takes
        1.05 x the largest |sigma'| on its 3 x 3 lattice nodes (corners, edge midpoints, centre); each further level the
        maxima of 2 x 2 blocks.  Delta tracking stays unbiased if a cell value misses a narrow peak between nodes --
        the weights 1 - sigma'/sigma_bar just leave [0, 1] there."""
        import torch.nn.functional as Fn

        a = torch.nan_to_num(lattice.abs().double(), nan=0.0, posinf=0.0)
        cells = Fn.max_pool2d(a[None, None], kernel_size=3, stride=2)[0, 0] * 1.05          # (N, N), i along x
        levels = [cells]
        while levels[-1].shape[0] > 1:
            levels.append(Fn.max_pool2d(levels[-1][None, None], kernel_size=2)[0, 0])
        (x0, x1), (y0, y1) = self._bounds()
        data = np.concatenate([l.numpy().astype(np.float32).ravel() for l in levels])
        return dict(data=data, levels=len(levels), x0=x0, y0=y0, dx=(x1 - x0) / N, dy=(y1 - y0) / N)

    def _sigma_prime_lattice_vectorised(self, n):
        (x0, x1), (y0, y1) = [[float(a), float(b)] for a, b in self.domain_bounds]
        X, Y = torch.meshgrid(torch.linspace(x0, x1, n), torch.linspace(y0, y1, n), indexing="ij")
        P = torch.stack([X.flatten(), Y.flatten()], dim=1).requires_grad_(True)
        ones = torch.ones(P.shape[0])
        sg = self.sigma(P) if self._sigma_given else torch.zeros(P.shape[0])
        al = torch.clamp(self.alpha(P), min=1e-8) if self._alpha_given else ones
        ratio = sg / al
        if self.sigma_prime_mode == "ratio":
            return ratio.detach()
        try:
            if not self._alpha_given:
                raise RuntimeError("constant alpha has no graph")         # reference: autograd.grad raises -> ratio
            (g,) = torch.autograd.grad(al.sum(), P, create_graph=True)
            lap = torch.zeros(P.shape[0]) + 1e-8
            try:
                for i in range(2):
                    lap = lap + torch.autograd.grad(g[:, i].sum(), P, create_graph=True, retain_graph=True)[0][:, i]
            except Exception:
                pass                                                       # utils.py:60-61: keep what was accumulated
            (lg,) = torch.autograd.grad(torch.log(al + 1e-8).sum(), P, create_graph=True)
            return (ratio + 0.5 * (lap / al - (lg ** 2).sum(dim=1) / 2.0)).detach()
        except Exception:
            if self.sigma_prime_mode == "full":
                raise
            self._autograd_failed = True
            return ratio.detach()

    # ------------------------------------------------------------------------------------------------
    # setters (reference :141-157)
    # ------------------------------------------------------------------------------------------------
    def setBoundaryConditions(self, boundaryDirichlet: callable):
        self._forget(self.boundaryDirichlet)
        self.boundaryDirichlet = boundaryDirichlet

    def setSourceTerm(self, source: callable):
        self._forget(self.source)
        self.source = source

    def invalidate(self):
        """Drop every cached device field / table (after changing state a coefficient callable closes over)."""
        self._cache = {k: v for k, v in self._cache.items() if k[0] == "scene"}

    def _forget(self, obj):
        if obj is not None and not isinstance(obj, Field):
            for k in [k for k in self._cache if k[0] in ("host", "dev") and k[1] == id(obj)]:
                del self._cache[k]

    # ------------------------------------------------------------------------------------------------
    # device objects
    # ------------------------------------------------------------------------------------------------
    def _bounds(self):
        return [[float(a), float(b)] for a, b in self.domain_bounds]

    def _probe_points(self):
        """A few fixed points inside the bounding box at which cached fields are re-checked against their callables."""
        if self._probe_pts is None:
            (x0, x1), (y0, y1) = self._bounds()
            u = torch.tensor([[0.5, 0.5], [0.21, 0.67], [0.83, 0.29], [0.37, 0.11], [0.64, 0.91]], dtype=torch.float32)
            self._probe_pts = torch.stack([x0 + u[:, 0] * (x1 - x0), y0 + u[:, 1] * (y1 - y0)], dim=1)
        return self._probe_pts

    def _still_valid(self, obj, field) -> bool:
        """The reference calls g / f / alpha / sigma live on every step; here a plain callable is traced or tabulated
        once.  If the state it closes over changed since (a source parameter, an electrode position), the cached device
        field would silently be stale: compare it with the callable at the probe points before every solve."""
        if isinstance(obj, Field) or getattr(obj, "__self__", None) is self:
            return True
        try:
            for p in self._probe_points():
                live = obj(p.clone())
                live = float(live.detach()) if isinstance(live, torch.Tensor) else float(live)
                have = float(field(p.reshape(1, 2))[0])
                if not (abs(live - have) <= 1e-3 * max(1.0, abs(live)) or (live != live and have != have)):
                    return False
        except Exception:
            return True                                                    # cannot be probed (e.g. needs grad): keep
        return True

    def _host_field(self, obj, n=None):
        key = ("host", id(obj), n)
        if key in self._cache and not self._still_valid(obj, self._cache[key][1]):
            self._forget(obj)
        if key not in self._cache:
            # sigma' tables (explicit n) evaluate autograd point by point: fixed resolution; everything else is refined
            self._cache[key] = (obj, as_field(obj, bounds=self._bounds(), n=n or self.field_resolution,
                                              tol=None if n else self.field_tolerance))
        return self._cache[key][1]

    def _dev_field(self, obj, device, n=None):
        if obj is None:
            return None
        key = ("dev", id(obj), device, n)
        host = self._host_field(obj, n)                                    # re-validates the cached field (and may drop it)
        if key not in self._cache or self._cache[key][2] is not host:
            self._cache[key] = (obj, nat.DeviceField(host, device), host)
        return self._cache[key][1]

    def _scene(self, device):
        # one scene per device, rebuilt when a boundary's `points` is another tensor or was written in place (version counter)
        dp = self.dirichletBoundary.points
        npts = None if self.neumannBoundary is None else self.neumannBoundary.points
        tok = tuple((id(p), getattr(p, "_version", None)) if isinstance(p, torch.Tensor) else (None if p is None else nat.host_f32(p).tobytes())
                    for p in (dp, npts))
        key = ("scene", device)
        hit = self._cache.get(key)
        if hit is None or hit[0] != tok:
            hit = (tok, nat.Scene(nat.host_f32(dp), None if npts is None else nat.host_f32(npts), device), (dp, npts))   # refs keep the ids unique
            self._cache[key] = hit
        return hit[1]

    def _device_problem(self, device):
        """Scene handle, wost_fields_t and delta-tracking parameters for one device."""
        g = None if self.boundaryDirichlet is _zero_boundary else self._dev_field(self.boundaryDirichlet, device)
        f = self._dev_field(self.source, device)
        alpha = sigma = sp = None
        icdf = None
        if self.use_delta_tracking:
            alpha = self._dev_field(self.alpha, device) if self._alpha_given else None
            sigma = self._dev_field(self.sigma, device) if self._sigma_given else None
            if self.sp_mode == SP_FIELD:
                sp = self._dev_field(self._sp_plain, device, self.sigma_prime_resolution)
            if self.compat != "physical":                                # physical mode samples the radius directly
                key = ("icdf", float(self.sigma_bar), device)
                if key not in self._cache:
                    self._cache[key] = torch.from_numpy(screened_radius_icdf(self.sigma_bar)).to(torch.device("cuda", device))
                icdf = self._cache[key]
        fields = nat.fields_struct(g=g, f=f, alpha=alpha, sigma=sigma, sigma_prime=sp)
        keep = (g, f, alpha, sigma, sp)
        return self._scene(device), fields, icdf, keep

    def _device_majorant(self, device):
        if self.majorant is None or not self.use_delta_tracking or self.compat != "physical":
            return None
        key = ("majorant", device)
        if key not in self._cache:
            self._cache[key] = dict(self.majorant, data=torch.from_numpy(self.majorant["data"]).to(torch.device("cuda", device)))
        return self._cache[key]

    def _sigma_prime_plain(self, point):
        """sigma' as a plain float callable, for tabulation when the coefficients are not analytic fields."""
        return float(self.sigma_prime(point))

    # ------------------------------------------------------------------------------------------------
    # solve (reference :319-353)
    # ------------------------------------------------------------------------------------------------
    def solve_raw(self, solvePoints, nWalks=1000, maxSteps=1000, eps=1e-4, *, seed=None, point_index_base=0,
                  point_index_stride=1, walk_offset=0, want_block_stats=False, want_walk_vals=False, n_trace=0, trace_cap=0,
                  device_outputs=False, device=None, jit=None, out=None):
        """One kernel pass over ``solvePoints`` on one device; returns the raw statistics dict
        (mean, m2, steps, optional block_stats / walk_vals / trace).  Building block of :meth:`solve`
        and of the multi-GPU driver (:mod:`dcrmontecarlo_b200.distributed`)."""
        nat.require_cuda()
        device = nat.current_device() if device is None else int(device)
        scene, fields, icdf, keep = self._device_problem(device)
        if seed is None:
            seed = _next_seed()
        res = nat.solve(scene, fields, solvePoints, int(nWalks), int(maxSteps), float(eps),
                        delta=self.use_delta_tracking, sp_mode=self.sp_mode,
                        sigma_bar=float(self.sigma_bar) if self.use_delta_tracking else 0.0, icdf=icdf, seed=seed,
                        point_index_base=point_index_base, point_index_stride=point_index_stride, walk_offset=walk_offset,
                        want_block_stats=want_block_stats,
                        want_walk_vals=want_walk_vals, n_trace=n_trace, trace_cap=trace_cap, device_outputs=device_outputs,
                        compat=self.compat, majorant=self._device_majorant(device), jit=jit or self.jit, out=out)
        res["seed"] = seed
        return res

    def solve_multi_source(self, solvePoints, sources, nWalks=1000, maxSteps=1000, eps=1e-4, *, seed=None,
                           want_block_stats=False, device_outputs=False, device=None, jit=None, point_index_base=0,
                           point_index_stride=1, walk_offset=0):
        """Shared-walk solve for many source terms (not in the reference, which re-walks per source): the walk does not
        depend on ``f``, so one set of walks gives the estimate for every source in ``sources`` (callables or fields).
        Returns ``mean`` / ``m2`` of shape ``(len(sources), P)``; row ``s`` equals what :meth:`solve_raw` returns with
        ``source = sources[s]`` and the same ``seed``."""
        nat.require_cuda()
        device = nat.current_device() if device is None else int(device)
        scene, fields, icdf, keep = self._device_problem(device)
        if all(isinstance(f, Field) for f in sources):                   # immutable field objects: one lookup for the whole list
            key = ("dev_sources", device, tuple(map(id, sources)))
            hit = self._cache.get(key)
            if hit is None:
                hit = self._cache[key] = (list(sources), [self._dev_field(f, device) for f in sources])
            devs = hit[1]
        else:
            devs = [self._dev_field(f, device) for f in sources]
        if seed is None:
            seed = _next_seed()
        res = nat.solve_multi_source(scene, fields, devs, solvePoints, int(nWalks), int(maxSteps), float(eps),
                                     delta=self.use_delta_tracking, sp_mode=self.sp_mode,
                                     sigma_bar=float(self.sigma_bar) if self.use_delta_tracking else 0.0, icdf=icdf, seed=seed,
                                     point_index_base=point_index_base, point_index_stride=point_index_stride,
                                     walk_offset=walk_offset,
                                     want_block_stats=want_block_stats, device_outputs=device_outputs, compat=self.compat,
                                     majorant=self._device_majorant(device), jit=jit or self.jit)
        res["seed"] = seed
        return res

    def solve(self, solvePoints: torch.Tensor, nWalks=1000, maxSteps=1000, eps=1e-4, return_history=False, *,
              seed=None, return_stats=False):
        """Estimate u at ``solvePoints`` ``(N, 2)`` with ``nWalks`` walks each; returns an ``(N, 1)`` float32 tensor
        (on the device of ``solvePoints``), or ``(tensor, history_dict)`` with ``return_history=True``."""
        pts = torch.as_tensor(solvePoints, dtype=torch.float32)
        P = int(pts.reshape(-1, 2).shape[0])
        n_trace = trace_cap = 0
        if return_history:
            trace_cap = int(min(maxSteps, 4096))
            n_trace = P * int(nWalks)
            if n_trace * (trace_cap + 1) * 32 > (1 << 30):                # (n_trace, trace_cap + 1, 8) float32, host and device
                raise ValueError("return_history would need more than 1 GiB of trace; lower nWalks, maxSteps or the point count")
        res = self.solve_raw(pts, nWalks, maxSteps, eps, seed=seed, want_walk_vals=return_history, n_trace=n_trace, trace_cap=trace_cap)
        mean = torch.from_numpy(res["mean"])
        n = float(nWalks)
        stderr = torch.sqrt(torch.from_numpy(res["m2"]) / max(n - 1.0, 1.0) / n)
        self.last_stats = dict(mean=mean, stderr=stderr, n_walks=int(nWalks), total_steps=int(res["steps"][0]), seed=res["seed"])
        out = mean.to(torch.float32).unsqueeze(1).to(pts.device)
        extras = []
        if return_history:
            if trace_cap < maxSteps and bool((np.asarray(res["trace_len"]) >= trace_cap).any()):
                import warnings

                warnings.warn(f"return_history: walks longer than {trace_cap} steps are truncated in the history "
                              "(their estimates are complete)", RuntimeWarning)
            extras.append(self._history(res, pts.reshape(-1, 2), int(nWalks)))
        if return_stats:
            extras.append(self.last_stats)
        return (out, *extras) if extras else out

    def _history(self, res, pts, W):
        """History dictionary in the reference's schema (:335-349) rebuilt from the device trace buffer: per walk the
        visited points with their cached distances, one 'source' contribution per step (when there is a source term),
        the final 'boundary' contribution, and the running point total (:308)."""
        hist = {}
        trace, tlen, vals = res["trace"], res["trace_len"], res["walk_vals"]
        has_neu, has_src = self.neumannBoundary is not None, self.source is not None
        for p in range(pts.shape[0]):
            running, walks = 0.0, []
            for w in range(W):
                flat = p * W + w
                n = int(tlen[flat])
                rows = trace[flat]
                path = [{"point": torch.tensor(rows[k, :2]), "dirichlet_distance": float(rows[k, 2]),
                         "neumann_distance": float(rows[k, 3]) if has_neu else None} for k in range(n)]
                contributions = []
                if has_src:
                    contributions += [{"step": k, "type": "source", "point": torch.tensor(rows[k, 4:6]), "contribution": float(rows[k, 6])}
                                      for k in range(n)]
                contributions.append({"step": int(rows[n, 4]), "type": "boundary", "point": torch.tensor(rows[n, :2]),
                                      "contribution": float(rows[n, 2])})
                running += float(vals[p, w])
                walks.append({"walk_id": w, "path": path, "contributions": contributions, "total_contribution": running})
            hist[p] = walks
        return hist
