"""Turn plain Python callables into exact device fields by symbolic tracing.

The reference hands arbitrary callables ``g, f, alpha, sigma`` on a ``(2,)`` tensor to the solver
(``solvers/WoStSolver.py:22``).  Many of them are closed-form expressions in ``point[0]``, ``point[1]`` built from
``+ - * / **`` with constants and ``torch.sin / cos / exp`` (e.g. every callable of
``tests/testWoStCorrectness.py:81-142``).  :func:`trace_callable` runs such a callable once on a symbolic point,
collects the expression as a sum of products

    A * x^i y^j * exp(P2(x, y)) * trig(L1(x, y)) * trig(L2(x, y))

and, if every product fits the device term algebra (``wost_term_t``: isotropic Gaussian, at most two trig factors
with linear arguments), returns the equivalent :class:`~fields.TermField` — evaluated analytically by the kernel, with
closed-form gradient and Laplacian.  Anything it cannot express (``float(...)`` casts, ``if`` on the coordinates,
division by a non-constant, other functions) makes it return ``None`` and the caller falls back to tabulation.  The
result is always checked numerically against the callable before it is accepted.
"""
from __future__ import annotations

import math

import numpy as np
import torch

try:
    from .fields import TermField, make_term
except ImportError:  # reference-style sys.path layout
    from fields import TermField, make_term

_MAX_POW = 16


class _Untraceable(Exception):
    pass


def _poly_mul(a: dict, b: dict) -> dict:
    out: dict = {}
    for (i, j), c in a.items():
        for (k, l), d in b.items():
            out[(i + k, j + l)] = out.get((i + k, j + l), 0.0) + c * d
    return {k: v for k, v in out.items() if v != 0.0}


def _poly_add(a: dict, b: dict, sb: float = 1.0) -> dict:
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0.0) + sb * v
    return {k: v for k, v in out.items() if v != 0.0}


def _freeze(p: dict):
    return tuple(sorted(p.items()))


class Sym:
    """Sum of products  coeff * monomial * exp(poly) * prod trig(linear).  Key of a product: (exp-arg, trig factors),
    value: the polynomial prefactor {(i, j): c}."""

    __slots__ = ("terms",)
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, terms=None):
        self.terms = terms or {}

    # ---- constructors ---------------------------------------------------------------------------------
    @staticmethod
    def const(c: float) -> "Sym":
        return Sym({((), ()): {(0, 0): float(c)}} if c != 0.0 else {})

    @staticmethod
    def var(axis: int) -> "Sym":
        return Sym({((), ()): {((1, 0) if axis == 0 else (0, 1)): 1.0}})

    @staticmethod
    def lift(v) -> "Sym":
        if isinstance(v, Sym):
            return v
        if isinstance(v, torch.Tensor):
            if v.numel() != 1:
                raise _Untraceable("non-scalar tensor constant")
            v = v.item()
        if isinstance(v, (int, float, np.floating, np.integer)):
            return Sym.const(float(v))
        raise _Untraceable(f"cannot lift {type(v)}")

    # ---- queries ------------------------------------------------------------------------------------------
    def as_poly(self) -> dict:
        """The expression as a plain polynomial, or raise."""
        if not self.terms:
            return {}
        if set(self.terms) != {((), ())}:
            raise _Untraceable("not a polynomial")
        return self.terms[((), ())]

    # ---- arithmetic -----------------------------------------------------------------------------------------
    def __add__(self, o):
        o = Sym.lift(o)
        out = {k: dict(v) for k, v in self.terms.items()}
        for k, v in o.terms.items():
            out[k] = _poly_add(out.get(k, {}), v)
        return Sym({k: v for k, v in out.items() if v})

    __radd__ = __add__

    def __neg__(self):
        return Sym({k: {m: -c for m, c in v.items()} for k, v in self.terms.items()})

    def __pos__(self):
        return self

    def __sub__(self, o):
        return self + (-Sym.lift(o))

    def __rsub__(self, o):
        return Sym.lift(o) + (-self)

    def __mul__(self, o):
        o = Sym.lift(o)
        out: dict = {}
        for (ea, ta), pa in self.terms.items():
            for (eb, tb), pb in o.terms.items():
                e = _freeze(_poly_add(dict(ea), dict(eb)))
                t = tuple(sorted(ta + tb))
                if len(t) > 2:
                    raise _Untraceable("more than two trig factors")
                key = (e, t)
                out[key] = _poly_add(out.get(key, {}), _poly_mul(pa, pb))
        return Sym({k: v for k, v in out.items() if v})

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Sym.lift(o)
        p = o.as_poly()
        if set(p) - {(0, 0)} or not p:
            raise _Untraceable("division by a non-constant")
        return self * (1.0 / p[(0, 0)])

    def __rtruediv__(self, o):
        raise _Untraceable("division by an expression")

    def __pow__(self, n):
        if isinstance(n, torch.Tensor):
            n = n.item()
        if not (isinstance(n, (int, float)) and float(n).is_integer() and 0 <= n <= _MAX_POW):
            raise _Untraceable("non-integer power")
        out = Sym.const(1.0)
        for _ in range(int(n)):
            out = out * self
        return out

    # anything that needs a concrete number makes the callable untraceable
    def _no(self, *a, **k):
        raise _Untraceable("needs a concrete value")

    __float__ = __int__ = __bool__ = __lt__ = __le__ = __gt__ = __ge__ = __eq__ = __ne__ = _no
    __hash__ = None

    def exp(self):
        p = self.as_poly()
        c = p.get((0, 0), 0.0)
        arg = {k: v for k, v in p.items() if k != (0, 0)}
        if any(i + j > 2 for (i, j) in arg):
            raise _Untraceable("exp of a polynomial of degree > 2")
        return Sym({(_freeze(arg), ()): {(0, 0): math.exp(c)}})

    def _trig(self, kind: str):
        p = self.as_poly()
        if any(i + j > 1 for (i, j) in p):
            raise _Untraceable("trig of a non-linear argument")
        fac = (kind, p.get((1, 0), 0.0), p.get((0, 1), 0.0), p.get((0, 0), 0.0))
        if fac[1] == 0.0 and fac[2] == 0.0:
            return Sym.const(math.sin(fac[3]) if kind == "sin" else math.cos(fac[3]))
        return Sym({((), (fac,)): {(0, 0): 1.0}})

    def sin(self):
        return self._trig("sin")

    def cos(self):
        return self._trig("cos")

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, "__name__", "")
        a = [Sym.lift(x) if isinstance(x, (Sym, int, float, torch.Tensor)) else x for x in args]
        if name == "sin":
            return a[0].sin()
        if name == "cos":
            return a[0].cos()
        if name == "exp":
            return a[0].exp()
        if name in ("add", "__add__", "__radd__"):
            return a[0] + a[1]
        if name in ("sub", "__sub__"):
            return a[0] - a[1]
        if name in ("__rsub__", "rsub"):
            return a[1] - a[0]
        if name in ("mul", "__mul__", "__rmul__"):
            return a[0] * a[1]
        if name in ("div", "true_divide", "__truediv__"):
            return a[0] / a[1]
        if name in ("pow", "__pow__"):
            return a[0] ** args[1]
        if name in ("neg", "__neg__"):
            return -a[0]
        if name == "square":
            return a[0] * a[0]
        raise _Untraceable(f"torch.{name} is not in the term algebra")


class _SymPoint:
    """What the callable sees as ``point``: indexing gives the symbolic coordinates."""

    def __getitem__(self, idx):
        if isinstance(idx, torch.Tensor):
            idx = int(idx)
        if idx in (0, -2):
            return Sym.var(0)
        if idx in (1, -1):
            return Sym.var(1)
        raise _Untraceable("point index out of range")

    def __iter__(self):
        return iter((Sym.var(0), Sym.var(1)))

    def __len__(self):
        return 2

    def __getattr__(self, name):
        raise _Untraceable(f"point.{name} is not traceable")


def _to_termfield(expr: Sym) -> TermField:
    c0, terms = 0.0, []
    for (earg, trigs), poly in expr.terms.items():
        q, cx, cy, scale = 0.0, 0.0, 0.0, 1.0
        if earg:
            e = dict(earg)
            qx, qy, mixed = -e.get((2, 0), 0.0), -e.get((0, 2), 0.0), e.get((1, 1), 0.0)
            if mixed != 0.0 or qx <= 0.0 or abs(qx - qy) > 1e-12 * max(abs(qx), abs(qy)):
                raise _Untraceable("exp argument is not an isotropic, decaying quadratic")
            q = qx
            cx, cy = e.get((1, 0), 0.0) / (2 * q), e.get((0, 1), 0.0) / (2 * q)
            scale = math.exp(q * (cx * cx + cy * cy))                     # completing the square
        t1 = trigs[0] if len(trigs) > 0 else None
        t2 = trigs[1] if len(trigs) > 1 else None
        for (i, j), c in poly.items():
            if i > _MAX_POW or j > _MAX_POW:
                raise _Untraceable("monomial power too large")
            if not earg and not trigs and (i, j) == (0, 0):
                c0 += c
                continue
            terms.append(make_term(A=c * scale, px=i, py=j, q=q, center=(cx, cy), trig1=t1, trig2=t2))
    return TermField(c0, terms)


def trace_callable(fn, bounds, n_check: int = 24, rtol: float = 2e-5):
    """``fn`` as an exact :class:`TermField`, or ``None`` if it is outside the term algebra or fails the numerical
    check against the callable on ``n_check`` random points of ``bounds = [[xmin, xmax], [ymin, ymax]]``."""
    try:
        out = fn(_SymPoint())
        field = _to_termfield(Sym.lift(out))
    except _Untraceable:
        return None
    except Exception:
        return None
    (x0, x1), (y0, y1) = [[float(b[0]), float(b[1])] for b in bounds]
    g = torch.Generator().manual_seed(12345)
    pts = torch.rand(n_check, 2, generator=g) * torch.tensor([x1 - x0, y1 - y0]) + torch.tensor([x0, y0])
    try:
        with torch.no_grad():
            ref = torch.tensor([float(fn(p)) for p in pts], dtype=torch.float64)
            got = field(pts).double()
    except Exception:
        return None
    scale = float(ref.abs().max()) + 1e-30
    if not torch.all((got - ref).abs() <= rtol * (ref.abs() + scale)):
        return None
    return field
