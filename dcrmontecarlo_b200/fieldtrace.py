"""Turn plain Python callables into exact device fields by symbolic tracing.

The reference hands arbitrary callables ``g, f, alpha, sigma`` on a ``(2,)`` tensor to the solver
(``solvers/WoStSolver.py:22``).  Most of the ones in its scripts are closed-form expressions in ``point[0]``,
``point[1]``: polynomials, ``torch.exp`` of an isotropic quadratic, ``torch.sin / cos`` of a linear form,
``torch_smooth_circle`` anomalies, wrapped in ``float(...)`` / ``torch.tensor(...)`` and guarded by an
``if`` that zeroes them outside a box or a disc (``tests/testWoStCorrectness.py:81-142``,
``tests/testWostWithSource.py:46-58``, ``tests/testWostVariableCoefficients.py:42-86``,
``tests/testGeophysicalScenario.py:11-55``).  :func:`trace_callable` runs such a callable on a symbolic point —
once per branch of its ``if`` — collects the expression as a sum of products

    A * x^i y^j * exp(P2(x, y)) * trig(L1(x, y)) * trig(L2(x, y))        or        A * sigmoid(a |x - c| + b)

and, if everything fits the device term algebra (``wost_term_t`` + box / disc mask), returns the equivalent
:class:`~fields.TermField`: evaluated analytically by the kernel, with closed-form gradient and Laplacian, instead of a
bilinear table (which cannot resolve, e.g., a 1 cm conductivity rim or a 0.5 m electrode on a 200 m domain).
Anything it cannot express makes it return ``None`` and the caller falls back to tabulation.  The result is always
checked numerically against the callable before it is accepted.
"""
from __future__ import annotations

import builtins
import math

import numpy as np
import threading

import torch

try:
    from .fields import TermField, make_circle_term, make_term
except ImportError:  # reference-style sys.path layout
    from fields import TermField, make_circle_term, make_term

_MAX_POW = 16
_BIG = 3.0e38


class _Untraceable(Exception):
    pass


def _poly_mul(a: dict, b: dict) -> dict:
    out: dict = {}
    for (i, j), c in a.items():
        for (k, l), d in b.items():
            out[(i + k, j + l)] = out.get((i + k, j + l), 0.0) + c * d
    return {k: v for k, v in out.items() if v != 0.0}


def _poly_add(a: dict, b: dict) -> dict:
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0.0) + v
    return {k: v for k, v in out.items() if v != 0.0}


def _freeze(p: dict):
    return tuple(sorted(p.items()))


def _number(v):
    if isinstance(v, torch.Tensor):
        if v.numel() != 1:
            raise _Untraceable("non-scalar tensor constant")
        return float(v.item())
    if isinstance(v, (int, float, np.floating, np.integer)):
        return float(v)
    raise _Untraceable(f"not a number: {type(v)}")


_ctx = {"path": None, "conds": None}        # branch enumeration state while a trace is running


class Cond:
    """``poly > 0`` — the outcome of comparing an expression with a constant; its truth value is dictated by the
    branch path being explored."""

    def __init__(self, poly: dict):
        self.poly = poly

    def __bool__(self):
        path, conds = _ctx["path"], _ctx["conds"]
        if path is None:
            raise _Untraceable("comparison outside a trace")
        i = len(conds)
        if i >= 8:
            raise _Untraceable("too many branches")
        conds.append(self)
        return path[i] if i < len(path) else False

    def _no(self, *a):
        raise _Untraceable("boolean algebra on conditions")

    __or__ = __and__ = __invert__ = _no


class Sym:
    """Sum of products.  Key of a product: (exp-arg, trig factors, circle factor); value: polynomial prefactor."""

    __slots__ = ("terms",)
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators
    __array_priority__ = 1000
    _K0 = ((), (), None)

    def __init__(self, terms=None):
        self.terms = terms or {}

    @staticmethod
    def const(c: float) -> "Sym":
        return Sym({Sym._K0: {(0, 0): float(c)}} if c != 0.0 else {})

    @staticmethod
    def var(axis: int) -> "Sym":
        return Sym({Sym._K0: {((1, 0) if axis == 0 else (0, 1)): 1.0}})

    @staticmethod
    def lift(v) -> "Sym":
        if isinstance(v, Sym):
            return v
        if isinstance(v, Affine):
            raise _Untraceable("a distance can only enter through sigmoid()")
        return Sym.const(_number(v))

    def as_poly(self) -> dict:
        if not self.terms:
            return {}
        if set(self.terms) != {Sym._K0}:
            raise _Untraceable("not a polynomial")
        return self.terms[Sym._K0]

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
        for (ea, ta, ca), pa in self.terms.items():
            for (eb, tb, cb), pb in o.terms.items():
                if ca is not None and cb is not None:
                    raise _Untraceable("product of two circle terms")
                e = _freeze(_poly_add(dict(ea), dict(eb)))
                t = tuple(sorted(ta + tb))
                if len(t) > 2:
                    raise _Untraceable("more than two trig factors")
                key = (e, t, ca if ca is not None else cb)
                out[key] = _poly_add(out.get(key, {}), _poly_mul(pa, pb))
        return Sym({k: v for k, v in out.items() if v})

    __rmul__ = __mul__

    def __truediv__(self, o):
        p = Sym.lift(o).as_poly()
        if set(p) - {(0, 0)} or not p:
            raise _Untraceable("division by a non-constant")
        return self * (1.0 / p[(0, 0)])

    def __rtruediv__(self, o):
        raise _Untraceable("division by an expression")

    def __pow__(self, n):
        n = _number(n)
        if n == 0.5:
            return self.sqrt()
        if not (n.is_integer() and 0 <= n <= _MAX_POW):
            raise _Untraceable("non-integer power")
        out = Sym.const(1.0)
        for _ in range(int(n)):
            out = out * self
        return out

    # ---- comparisons become branch conditions ---------------------------------------------------------------------
    def _cmp(self, o, sign):
        d = (self - Sym.lift(o)).as_poly()
        return Cond({k: sign * v for k, v in d.items()})

    def __gt__(self, o):
        return self._cmp(o, +1.0)

    __ge__ = __gt__

    def __lt__(self, o):
        return self._cmp(o, -1.0)

    __le__ = __lt__

    def _no(self, *a, **k):
        raise _Untraceable("needs a concrete value")

    __float__ = __int__ = __bool__ = __eq__ = __ne__ = _no
    __hash__ = None

    # ---- functions ---------------------------------------------------------------------------------------------
    def exp(self):
        p = self.as_poly()
        c = p.get((0, 0), 0.0)
        arg = {k: v for k, v in p.items() if k != (0, 0)}
        if any(i + j > 2 for (i, j) in arg):
            raise _Untraceable("exp of a polynomial of degree > 2")
        return Sym({(_freeze(arg), (), None): {(0, 0): math.exp(c)}})

    def _trig(self, kind: str):
        p = self.as_poly()
        if any(i + j > 1 for (i, j) in p):
            raise _Untraceable("trig of a non-linear argument")
        fac = (kind, p.get((1, 0), 0.0), p.get((0, 1), 0.0), p.get((0, 0), 0.0))
        if fac[1] == 0.0 and fac[2] == 0.0:
            return Sym.const(math.sin(fac[3]) if kind == "sin" else math.cos(fac[3]))
        return Sym({((), (fac,), None): {(0, 0): 1.0}})

    def sin(self):
        return self._trig("sin")

    def cos(self):
        return self._trig("cos")

    def sqrt(self):
        """sqrt(a ((x-cx)^2 + (y-cy)^2)) -> sqrt(a) times the distance to (cx, cy)."""
        p = self.as_poly()
        a = p.get((2, 0), 0.0)
        if a <= 0.0 or p.get((0, 2), 0.0) != a or p.get((1, 1), 0.0) != 0.0 or any(i + j > 2 for (i, j) in p):
            raise _Untraceable("sqrt of something that is not a squared distance")
        cx, cy = -p.get((1, 0), 0.0) / (2 * a), -p.get((0, 1), 0.0) / (2 * a)
        if abs(p.get((0, 0), 0.0) - a * (cx * cx + cy * cy)) > 1e-9 * (abs(a) * (cx * cx + cy * cy) + 1.0):
            raise _Untraceable("sqrt of a quadratic with an offset")
        return Affine(cx, cy, math.sqrt(a), 0.0)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _dispatch(func, args)


class Affine:
    """``a * |x - c| + b``: what a distance may become before it enters ``sigmoid`` (reference utils.py:123-129)."""

    __slots__ = ("cx", "cy", "a", "b")
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, cx, cy, a, b):
        self.cx, self.cy, self.a, self.b = float(cx), float(cy), float(a), float(b)

    def __add__(self, o):
        return Affine(self.cx, self.cy, self.a, self.b + _number(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Affine(self.cx, self.cy, self.a, self.b - _number(o))

    def __rsub__(self, o):
        return Affine(self.cx, self.cy, -self.a, _number(o) - self.b)

    def __neg__(self):
        return Affine(self.cx, self.cy, -self.a, -self.b)

    def __mul__(self, o):
        o = _number(o)
        return Affine(self.cx, self.cy, self.a * o, self.b * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self * (1.0 / _number(o))

    def sigmoid(self):
        # sigmoid(a rho + b) = sigmoid(-k (rho - R)) with k = -a, R = b / k
        if self.a == 0.0:
            return Sym.const(1.0 / (1.0 + math.exp(-self.b)))
        k = -self.a
        if k < 0.0:                                   # rising step: 1 - sigmoid(-|k| (rho - R))
            return Sym.const(1.0) - Affine(self.cx, self.cy, -self.a, -self.b).sigmoid()
        return Sym({((), (), (k, self.cx, self.cy, self.b / k)): {(0, 0): 1.0}})

    def _no(self, *a, **k):
        raise _Untraceable("a distance can only enter through sigmoid()")

    __float__ = __bool__ = __lt__ = __gt__ = __le__ = __ge__ = __pow__ = _no

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _dispatch(func, args)


class _Vec2:
    """``point - center``: only its norm is meaningful."""

    def __init__(self, dx: Sym, dy: Sym):
        self.dx, self.dy = dx, dy

    def norm(self, *a, **k):
        return (self.dx * self.dx + self.dy * self.dy).sqrt()

    def __getitem__(self, i):
        return (self.dx, self.dy)[int(i)]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if getattr(func, "__name__", "") == "norm":
            return args[0].norm()
        raise _Untraceable("vector op outside the algebra")


def _dispatch(func, args):
    name = getattr(func, "__name__", "")
    a0 = args[0] if args else None
    if name in ("sin", "cos", "exp", "sqrt", "sigmoid"):
        if isinstance(a0, (Sym, Affine)) and hasattr(a0, name):
            return getattr(a0, name)()
        raise _Untraceable(f"torch.{name} of {type(a0).__name__}")
    symbolic = (Sym, Affine)
    if name in ("add", "__add__", "__radd__"):
        return args[0] + args[1] if isinstance(args[0], symbolic) else args[1] + args[0]
    if name in ("sub", "__sub__"):
        return args[0] - args[1] if isinstance(args[0], symbolic) else args[1].__rsub__(args[0])
    if name in ("__rsub__", "rsub"):
        return args[0].__rsub__(args[1])
    if name in ("mul", "__mul__", "__rmul__"):
        return args[0] * args[1] if isinstance(args[0], symbolic) else args[1] * args[0]
    if name in ("div", "true_divide", "__truediv__"):
        if isinstance(args[0], symbolic):
            return args[0] / args[1]
        raise _Untraceable("division by an expression")
    if name in ("pow", "__pow__"):
        return args[0] ** args[1]
    if name in ("neg", "__neg__"):
        return -args[0]
    if name == "square":
        return args[0] * args[0]
    raise _Untraceable(f"torch.{name} is not in the term algebra")


class _SymPoint:
    """What the callable sees as ``point``: indexing gives the symbolic coordinates, ``point - c`` a vector."""

    __array_ufunc__ = None
    __array_priority__ = 1000

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

    def __sub__(self, c):
        c = torch.as_tensor(c, dtype=torch.float64).reshape(-1)
        if c.numel() != 2:
            raise _Untraceable("point - something that is not a 2-vector")
        return _Vec2(Sym.var(0) - float(c[0]), Sym.var(1) - float(c[1]))

    def norm(self, *a, **k):
        return (Sym.var(0) * Sym.var(0) + Sym.var(1) * Sym.var(1)).sqrt()

    def __getattr__(self, name):
        raise _Untraceable(f"point.{name} is not traceable")

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, "__name__", "")
        if name in ("sub", "__sub__") and isinstance(args[0], _SymPoint):
            return args[0] - args[1]
        if name == "norm" and isinstance(args[0], _SymPoint):
            return args[0].norm()
        raise _Untraceable("point op outside the algebra")


# ---- expression -> TermField -----------------------------------------------------------------------------------
def _to_termfield(expr: Sym) -> TermField:
    c0, terms = 0.0, []
    for (earg, trigs, circ), poly in expr.terms.items():
        if circ is not None:
            if earg or trigs or set(poly) - {(0, 0)}:
                raise _Untraceable("circle term multiplied by something non-constant")
            k, cx, cy, R = circ
            terms.append(make_circle_term(poly[(0, 0)], (cx, cy), R, k))
            continue
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


def _mask_from_conditions(conds):
    """Conditions that all lead to the 'outside' value when true -> ('box', xmin, xmax, ymin, ymax) or ('disc', c, R)."""
    box = [-_BIG, _BIG, -_BIG, _BIG]
    disc = None
    for c in conds:
        p = c.poly                                                       # outside  <=>  p > 0
        keys = set(p) - {(0, 0)}
        c0 = p.get((0, 0), 0.0)
        if keys == {(1, 0)}:
            a = p[(1, 0)]
            if a > 0:
                box[1] = min(box[1], -c0 / a)                            # a x + c0 > 0  <=>  x > -c0/a
            else:
                box[0] = max(box[0], -c0 / a)                            # x < -c0/a
        elif keys == {(0, 1)}:
            a = p[(0, 1)]
            if a > 0:
                box[3] = min(box[3], -c0 / a)
            else:
                box[2] = max(box[2], -c0 / a)
        elif (2, 0) in p and p.get((0, 2)) == p[(2, 0)] and p[(2, 0)] > 0 and p.get((1, 1), 0.0) == 0.0 and all(i + j <= 2 for i, j in p):
            a = p[(2, 0)]
            cx, cy = -p.get((1, 0), 0.0) / (2 * a), -p.get((0, 1), 0.0) / (2 * a)
            R2 = cx * cx + cy * cy - c0 / a                              # |x-c|^2 > R2
            if R2 <= 0 or disc is not None:
                raise _Untraceable("unsupported disc condition")
            disc = ((cx, cy), math.sqrt(R2))
        else:
            raise _Untraceable("branch condition is neither a half-plane nor a disc")
    if disc is not None:
        if box != [-_BIG, _BIG, -_BIG, _BIG]:
            raise _Untraceable("box and disc conditions mixed")
        return ("disc",) + disc
    return ("box", *box)


def _run(fn, path):
    _ctx["path"], _ctx["conds"] = list(path), []
    try:
        out = fn(_SymPoint())
        return out, list(_ctx["conds"])
    finally:
        _ctx["path"], _ctx["conds"] = None, None


def _passthrough_float(v):
    return v if isinstance(v, (Sym, Affine)) else builtins.float(v)


def _symbolic(fn) -> TermField:
    """Trace ``fn`` including one level of ``if <outside>: return <const>`` guards."""
    inside, conds = _run(fn, [])
    field = _to_termfield(Sym.lift(inside))
    if not conds:
        return field
    outside = None
    for i in range(len(conds)):
        out_i, conds_i = _run(fn, [False] * i + [True])
        if len(conds_i) != i + 1:
            raise _Untraceable("branch structure changes between runs")
        p = Sym.lift(out_i).as_poly()
        if set(p) - {(0, 0)}:
            raise _Untraceable("the guarded branch does not return a constant")
        v = p.get((0, 0), 0.0)
        if outside is not None and v != outside:
            raise _Untraceable("guards return different constants")
        outside = v
    mask = _mask_from_conditions(conds)
    if mask[0] == "disc":
        return field.masked_disc(mask[1], mask[2], outside=outside)
    return field.masked_box(*mask[1:], outside=outside)


_TRACE_LOCK = threading.RLock()


def trace_callable(fn, bounds, n_check: int = 32, rtol: float = 2e-5):
    """``fn`` as an exact :class:`TermField`, or ``None`` if it is outside the term algebra or fails the numerical
    check against the callable on ``n_check`` random points of ``bounds = [[xmin, xmax], [ymin, ymax]]``."""
    # `float(expr)` and `torch.tensor(expr)` wrappers (tests/testWostWithSource.py:48, testWostVariableCoefficients.py:49)
    # must let the symbolic value through while tracing: shadow `float` in the callable's globals, patch torch.tensor.
    g = getattr(fn, "__globals__", None)
    # the patching below touches process-wide names (torch.tensor, the callable's module globals): one tracer at a time,
    # and other threads only ever see the pass-through wrappers, which behave like the originals for ordinary values
    with _TRACE_LOCK:
        had_float = g is not None and "float" in g
        old_float = g.get("float") if had_float else None
        old_tensor = torch.tensor
        try:
            if g is not None:
                g["float"] = _passthrough_float
            torch.tensor = lambda v, *a, **k: v if isinstance(v, (Sym, Affine)) else old_tensor(v, *a, **k)
            try:
                field = _symbolic(fn)
            except Exception:                   # _Untraceable, or whatever the callable does with a symbolic point
                return None
        finally:
            torch.tensor = old_tensor
            if g is not None:
                if had_float:
                    g["float"] = old_float
                else:
                    g.pop("float", None)
    (x0, x1), (y0, y1) = [[float(b[0]), float(b[1])] for b in bounds]
    gen = torch.Generator().manual_seed(12345)
    mx, my = 0.05 * (x1 - x0), 0.05 * (y1 - y0)                          # a little beyond the box: masks are tested too
    pts = torch.rand(n_check, 2, generator=gen) * torch.tensor([x1 - x0 + 2 * mx, y1 - y0 + 2 * my]) + torch.tensor([x0 - mx, y0 - my])
    try:
        ref = torch.tensor([float(torch.as_tensor(fn(p)).detach()) for p in pts], dtype=torch.float64)
        got = field(pts).double()
    except Exception:
        return None
    scale = float(ref.abs().max()) + 1e-30
    if not torch.all((got - ref).abs() <= rtol * (ref.abs() + scale)):
        return None
    return field
