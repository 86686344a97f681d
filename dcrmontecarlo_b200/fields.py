"""Device field descriptors for the callables the reference solver takes.

The reference's ``WostSolver_2D`` receives arbitrary Python callables ``g, f, alpha, sigma`` that it
evaluates at a ``(2,)`` tensor inside the walk loop (reference ``solvers/WoStSolver.py:22,253-256,
277-283,295``).  A CUDA walk kernel cannot call back into Python, so the solver consumes *field
descriptors* instead:

* :class:`TermField` — a constant plus a sum of analytic terms, each a product
  ``A * x^px y^py * exp(-q |x-c|^2) * trig1(w1.x + p1) * trig2(w2.x + p2)`` or a smooth circle
  ``A * sigmoid(-k (|x-c| - R))`` (reference ``utils.py:123-129``), optionally masked by a box or a
  disc.  Covers every coefficient the reference's five scenarios use, with closed-form gradient and
  Laplacian on the device (needed for the delta-tracking ``sigma'``, ``solvers/WoStSolver.py:88-121``).
* :class:`GridField` — a bilinear table over a box, the route for arbitrary callables
  (:meth:`GridField.from_callable`).

Every field is itself a callable on ``(2,)`` (or ``(N, 2)``) float32 tensors using torch ops in the
same order as the device code, so the *same object* can be handed to the reference solver (parity
fixtures are generated that way) and supports autograd.
"""
from __future__ import annotations

import math
from typing import Callable, Iterable, Sequence

import numpy as np
import torch

TERM_PRODUCT, TERM_SIGMOID_CIRCLE = 0, 1
TRIG_NONE, TRIG_SIN, TRIG_COS = 0, 1, 2
FIELD_TERMS, FIELD_GRID = 0, 1
MASK_NONE, MASK_BOX, MASK_DISC = 0, 1, 2

# 16 x 4-byte words; mirrors wost_term_t in include/wost.h
TERM_DTYPE = np.dtype(
    [("kind", "<i4"), ("px", "<i4"), ("py", "<i4"), ("t1", "<i4"), ("t2", "<i4"), ("A", "<f4"),
     ("q", "<f4"), ("cx", "<f4"), ("cy", "<f4"), ("R", "<f4"),
     ("w1x", "<f4"), ("w1y", "<f4"), ("p1", "<f4"), ("w2x", "<f4"), ("w2y", "<f4"), ("p2", "<f4")]
)
_TRIG = {None: TRIG_NONE, "none": TRIG_NONE, "sin": TRIG_SIN, "cos": TRIG_COS}


def _f32(v) -> float:
    return float(np.float32(v))


def make_term(A=1.0, px=0, py=0, q=0.0, center=(0.0, 0.0), trig1=None, trig2=None) -> np.ndarray:
    """One PRODUCT term. ``trigN = (kind, wx, wy, phase)`` with kind 'sin' | 'cos'."""
    t = np.zeros((), dtype=TERM_DTYPE)
    t["kind"], t["px"], t["py"], t["A"], t["q"] = TERM_PRODUCT, int(px), int(py), A, q
    t["cx"], t["cy"] = center
    for name, trig in (("1", trig1), ("2", trig2)):
        if trig is not None:
            kind, wx, wy, ph = trig
            t["t" + name] = _TRIG[kind]
            t["w" + name + "x"], t["w" + name + "y"], t["p" + name] = wx, wy, ph
    return t


def make_circle_term(A, center, R, k=100.0) -> np.ndarray:
    t = np.zeros((), dtype=TERM_DTYPE)
    t["kind"], t["A"], t["q"], t["R"] = TERM_SIGMOID_CIRCLE, A, k, R
    t["cx"], t["cy"] = center
    return t


def _ipow(x: torch.Tensor, p: int) -> torch.Tensor:
    r = torch.ones_like(x)
    for _ in range(p):
        r = r * x
    return r


class Field:
    """Base class: a callable on points plus a plain description the C ABI packs."""

    mask_kind = MASK_NONE
    mask = (0.0, 0.0, 0.0, 0.0)
    outside = 0.0

    def describe(self) -> dict:  # pragma: no cover - abstract
        raise NotImplementedError

    def _raw(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, point: torch.Tensor) -> torch.Tensor:
        point = torch.as_tensor(point)
        x, y = point[..., 0], point[..., 1]
        v = self._raw(x, y)
        if self.mask_kind == MASK_BOX:
            xmin, xmax, ymin, ymax = self.mask
            out = (x < xmin) | (x > xmax) | (y < ymin) | (y > ymax)
            v = torch.where(out, torch.full_like(v, self.outside), v)
        elif self.mask_kind == MASK_DISC:
            cx, cy, r2, _ = self.mask
            ddx, ddy = x - cx, y - cy
            out = ddx * ddx + ddy * ddy > r2
            v = torch.where(out, torch.full_like(v, self.outside), v)
        return v

    # masks ------------------------------------------------------------------------------------
    def _with_mask(self, kind, mask, outside):
        import copy

        f = copy.copy(self)
        f.mask_kind, f.mask, f.outside = kind, tuple(_f32(m) for m in mask), _f32(outside)
        return f

    def masked_box(self, xmin, xmax, ymin, ymax, outside=0.0) -> "Field":
        """Value ``outside`` where x<xmin, x>xmax, y<ymin or y>ymax (reference tests/testWostWithSource.py:51-56)."""
        return self._with_mask(MASK_BOX, (xmin, xmax, ymin, ymax), outside)

    def masked_disc(self, center, R, outside=0.0) -> "Field":
        """Value ``outside`` where |x-c|^2 > R^2 (reference tests/testWostVariableCoefficients.py:80-84)."""
        return self._with_mask(MASK_DISC, (center[0], center[1], float(R) ** 2, 0.0), outside)


class TermField(Field):
    def __init__(self, c0: float = 0.0, terms: Iterable[np.ndarray] = ()):
        self.c0 = _f32(c0)
        terms = list(terms)
        self.terms = np.array(terms, dtype=TERM_DTYPE) if terms else np.zeros(0, dtype=TERM_DTYPE)

    # constructors -----------------------------------------------------------------------------
    @staticmethod
    def constant(c: float) -> "TermField":
        return TermField(c)

    @staticmethod
    def polynomial(coeffs: dict) -> "TermField":
        """``{(i, j): c}`` -> sum c x^i y^j."""
        c0 = coeffs.get((0, 0), 0.0)
        return TermField(c0, [make_term(A=c, px=i, py=j) for (i, j), c in coeffs.items() if (i, j) != (0, 0) and c != 0.0])

    @staticmethod
    def gaussian_sum(blobs: Sequence[tuple], base: float = 0.0) -> "TermField":
        """``[(A, (cx, cy), q)]`` -> base + sum A exp(-q |x-c|^2)."""
        return TermField(base, [make_term(A=A, q=q, center=c) for A, c, q in blobs])

    @staticmethod
    def smooth_circle_sum(base: float, circles: Sequence[tuple], k: float = 100.0) -> "TermField":
        """``[(A, (cx, cy), R)]`` -> base + sum A sigmoid(-k(|x-c|-R)) (reference utils.py:123-129)."""
        return TermField(base, [make_circle_term(A, c, R, k) for A, c, R in circles])

    def __add__(self, other):
        if isinstance(other, (int, float)):
            return TermField(self.c0 + other, list(self.terms))
        if isinstance(other, TermField):
            if self.mask_kind != MASK_NONE or other.mask_kind != MASK_NONE:
                raise ValueError("cannot add masked fields")
            return TermField(self.c0 + other.c0, list(self.terms) + list(other.terms))
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, s):
        if not isinstance(s, (int, float)):
            return NotImplemented
        terms = []
        for t in self.terms:
            t = t.copy()
            t["A"] = t["A"] * s
            terms.append(t)
        f = TermField(self.c0 * s, terms)
        f.mask_kind, f.mask, f.outside = self.mask_kind, self.mask, _f32(self.outside * s)
        return f

    __rmul__ = __mul__

    # evaluation -------------------------------------------------------------------------------
    def _raw(self, x, y):
        v = torch.zeros_like(x) + self.c0
        for t in self.terms:
            A = float(t["A"])
            if int(t["kind"]) == TERM_SIGMOID_CIRCLE:
                ddx, ddy = x - float(t["cx"]), y - float(t["cy"])
                rho = torch.sqrt(ddx * ddx + ddy * ddy)
                v = v + A * torch.sigmoid(-(float(t["q"]) * (rho - float(t["R"]))))
                continue
            tv = torch.zeros_like(x) + A
            if int(t["px"]) or int(t["py"]):
                tv = tv * (_ipow(x, int(t["px"])) * _ipow(y, int(t["py"])))
            if float(t["q"]) != 0.0:
                ddx, ddy = x - float(t["cx"]), y - float(t["cy"])
                tv = tv * torch.exp(-float(t["q"]) * (ddx * ddx + ddy * ddy))
            for n in ("1", "2"):
                kind = int(t["t" + n])
                if kind == TRIG_NONE:
                    continue
                a = float(t["w" + n + "x"]) * x + float(t["w" + n + "y"]) * y + float(t["p" + n])
                tv = tv * (torch.sin(a) if kind == TRIG_SIN else torch.cos(a))
            v = v + tv
        return v

    def describe(self) -> dict:
        return dict(kind=FIELD_TERMS, c0=self.c0, terms=self.terms, mask_kind=self.mask_kind, mask=self.mask,
                    outside=self.outside, grid=None, nx=0, ny=0, x0=0.0, y0=0.0, dx=1.0, dy=1.0)


class GridField(Field):
    """Bilinear table: node (i, j) sits at (x0 + i dx, y0 + j dy); lookups clamp to the table."""

    def __init__(self, values, x0, y0, dx, dy):
        self.values = np.ascontiguousarray(np.asarray(values, dtype=np.float32))
        assert self.values.ndim == 2 and min(self.values.shape) >= 2
        self.nx, self.ny = self.values.shape
        self.x0, self.y0, self.dx, self.dy = _f32(x0), _f32(y0), _f32(dx), _f32(dy)
        self._t = torch.from_numpy(self.values)

    @staticmethod
    def _evaluate(fn: Callable, X: torch.Tensor, Y: torch.Tensor, vectorised):
        """``fn`` on every (X, Y) pair -> (flat float32 values, vectorised?).  One vectorised call is tried first (the
        callable sees ``point[0]``, ``point[1]`` as vectors) and verified on a few points; else the per-point loop.
        No torch.no_grad(): callables may differentiate internally (the solver's sigma' does)."""
        xf, yf = X.flatten(), Y.flatten()
        N = xf.numel()

        def one(i):
            v = fn(torch.stack([xf[i], yf[i]]))
            return float(v.detach()) if isinstance(v, torch.Tensor) else float(v)

        if vectorised is not False:
            try:
                out = torch.as_tensor(fn(torch.stack([xf, yf], dim=0)), dtype=torch.float32).detach()
                if out.shape == (N,):
                    probe = [0, N // 3, N // 2 + 7, N - 1]
                    if vectorised is True or all(abs(one(i) - float(out[i])) <= 1e-5 * (1.0 + abs(float(out[i]))) for i in probe):
                        return out, True
            except Exception:
                pass
        flat = torch.empty(N, dtype=torch.float32)
        for i in range(N):
            flat[i] = one(i)
        return flat, False

    @staticmethod
    def from_callable(fn: Callable, bounds, n: int = 257, margin: float = 0.02, tol: float | None = None,
                      n_max: int = 2049) -> "GridField":
        """Tabulate ``fn(point)`` on an ``n x n`` lattice over ``bounds = [[xmin,xmax],[ymin,ymax]]`` grown by ``margin``
        (relative) — walks may step a hair outside the Dirichlet boundary (reference solvers/WoStSolver.py:206-215).

        Error control: with ``tol`` the bilinear interpolant is compared with ``fn`` at the cell centres (where its error
        peaks) and the lattice is refined (n -> 2n-1, nodes are reused) until the largest error is below ``tol`` times
        the value range, or ``n_max`` is reached (then a warning is issued).  The achieved error is kept in
        ``.interp_error``.  Only attempted for callables that can be evaluated in one vectorised call."""
        (xmin, xmax), (ymin, ymax) = [[float(b[0]), float(b[1])] for b in bounds]
        mx, my = margin * (xmax - xmin), margin * (ymax - ymin)
        vectorised = None
        while True:
            xs = torch.linspace(xmin - mx, xmax + mx, n)
            ys = torch.linspace(ymin - my, ymax + my, n)
            X, Y = torch.meshgrid(xs, ys, indexing="ij")
            flat, vectorised = GridField._evaluate(fn, X, Y, vectorised)
            gf = GridField(flat.reshape(n, n).numpy(), float(xs[0]), float(ys[0]), float(xs[1] - xs[0]), float(ys[1] - ys[0]))
            gf.interp_error = None
            if tol is None or not vectorised:
                return gf
            xc, yc = 0.5 * (xs[:-1] + xs[1:]), 0.5 * (ys[:-1] + ys[1:])
            Xc, Yc = torch.meshgrid(xc, yc, indexing="ij")
            exact, _ = GridField._evaluate(fn, Xc, Yc, True)
            approx = gf(torch.stack([Xc.flatten(), Yc.flatten()], dim=1))
            finite = torch.isfinite(exact)
            scale = float((flat[torch.isfinite(flat)].max() - flat[torch.isfinite(flat)].min()).abs()) + 1e-30
            gf.interp_error = float((approx - exact)[finite].abs().max()) / scale if finite.any() else 0.0
            if gf.interp_error <= tol:
                return gf
            if 2 * n - 1 > n_max:
                import warnings

                warnings.warn(f"tabulated field: interpolation error {gf.interp_error:.2e} of the value range at n={n} "
                              f"(tolerance {tol:.1e}); pass an analytic fields.TermField for an exact device field")
                return gf
            n = 2 * n - 1

    def _raw(self, x, y):
        fx = torch.clamp((x - self.x0) / self.dx, 0.0, float(self.nx - 1))
        fy = torch.clamp((y - self.y0) / self.dy, 0.0, float(self.ny - 1))
        i = torch.clamp(fx.detach().floor().long(), max=self.nx - 2)
        j = torch.clamp(fy.detach().floor().long(), max=self.ny - 2)
        tx, ty = fx - i, fy - j
        g = self._t
        v00, v01, v10, v11 = g[i, j], g[i, j + 1], g[i + 1, j], g[i + 1, j + 1]
        a = v00 + ty * (v01 - v00)
        b = v10 + ty * (v11 - v10)
        return a + tx * (b - a)

    def describe(self) -> dict:
        return dict(kind=FIELD_GRID, c0=0.0, terms=np.zeros(0, dtype=TERM_DTYPE), mask_kind=self.mask_kind,
                    mask=self.mask, outside=self.outside, grid=self.values, nx=self.nx, ny=self.ny,
                    x0=self.x0, y0=self.y0, dx=self.dx, dy=self.dy)


def as_field(obj, bounds=None, n: int = 257, trace: bool = True, tol: float | None = None) -> Field | None:
    """``None`` stays ``None``; numbers become constants; Field passes through; any other callable is first traced
    symbolically into an exact :class:`TermField` (:mod:`fieldtrace`) and, if that is not possible, tabulated."""
    if obj is None or isinstance(obj, Field):
        return obj
    if isinstance(obj, (int, float)):
        return TermField.constant(float(obj))
    if callable(obj):
        if bounds is None:
            raise ValueError("tabulating a callable needs domain bounds")
        if trace:
            try:
                from .fieldtrace import trace_callable
            except ImportError:  # reference-style sys.path layout
                from fieldtrace import trace_callable

            exact = trace_callable(obj, bounds)
            if exact is not None:
                return exact
        return GridField.from_callable(obj, bounds, n=n, tol=tol)
    raise TypeError(f"cannot turn {type(obj)} into a field")


PI32 = _f32(math.pi)
