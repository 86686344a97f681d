"""ctypes front-end of the CPU oracle (``oracle/wost_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libwost_oracle.so"

RNG_MT, RNG_PHILOX = 0, 1
SP_FULL, SP_RATIO, SP_FIELD = 0, 1, 2


def build(force: bool = False) -> Path:
    src = [HERE / "wost_oracle.c", HERE / "wost_oracle.h", HERE.parent / "include" / "wost_math.h"]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in src):
        subprocess.check_call(["make", "-s", "-C", str(HERE), "libwost_oracle.so"] + (["-B"] if force else []))
    return LIB_PATH


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("px", C.c_int32), ("py", C.c_int32), ("t1", C.c_int32), ("t2", C.c_int32),
                ("A", C.c_float), ("q", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("R", C.c_float),
                ("w1x", C.c_float), ("w1y", C.c_float), ("p1", C.c_float),
                ("w2x", C.c_float), ("w2y", C.c_float), ("p2", C.c_float)]


class _Field(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_terms", C.c_int32), ("c0", C.c_float), ("mask_kind", C.c_int32),
                ("mask", C.c_float * 4), ("outside", C.c_float), ("nx", C.c_int32), ("ny", C.c_int32),
                ("x0", C.c_float), ("y0", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
                ("terms", C.c_void_p), ("grid", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("dir_pts", C.c_void_p), ("n_dir", C.c_int32), ("neu_pts", C.c_void_p), ("n_neu", C.c_int32),
                ("f", C.POINTER(_Field)), ("alpha", C.POINTER(_Field)), ("sigma", C.POINTER(_Field)),
                ("sigma_prime", C.POINTER(_Field)), ("g", C.POINTER(_Field)),
                ("delta", C.c_int32), ("sp_mode", C.c_int32), ("sigma_bar", C.c_float),
                ("n_walks", C.c_int64), ("max_steps", C.c_int32), ("eps", C.c_float),
                ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("seed_numpy", C.c_uint64),
                ("point_index_base", C.c_int64), ("walk_offset", C.c_int64),
                ("icdf", C.c_void_p), ("icdf_len", C.c_int32), ("n_threads", C.c_int32),
                ("sincos_fn", C.c_void_p), ("atan2_fn", C.c_void_p), ("compat_mode", C.c_int32), ("phys_nudge", C.c_float),
                ("maj_levels", C.c_int32), ("majorant", C.c_void_p), ("maj_x0", C.c_float), ("maj_y0", C.c_float),
                ("maj_dx", C.c_float), ("maj_dy", C.c_float)]


SINCOS_FN = C.CFUNCTYPE(None, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float))
ATAN2_FN = C.CFUNCTYPE(C.c_float, C.c_float, C.c_float)


def torch_trig_callbacks():
    """cos/sin/atan2 evaluated by torch's CPU kernels exactly as the reference calls them
    (solvers/WoStSolver.py:228-231): 1-element float32 tensors."""
    import torch

    def _sincos(theta, c, s):
        t = torch.tensor([theta], dtype=torch.float32)
        c[0] = torch.cos(t).item()
        s[0] = torch.sin(t).item()

    def _atan2(y, x):
        return torch.atan2(torch.tensor(y, dtype=torch.float32), torch.tensor(x, dtype=torch.float32)).item()

    return SINCOS_FN(_sincos), ATAN2_FN(_atan2)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_distance.restype = C.c_float
        L.orc_distance.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.orc_silhouette_distance.restype = C.c_float
        L.orc_silhouette_distance.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.orc_is_silhouette.restype = C.c_int
        L.orc_is_silhouette.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.orc_ray_intersection.restype = None
        L.orc_ray_intersection.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_intersect_polylines.restype = C.c_int
        L.orc_intersect_polylines.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 5 + [C.c_void_p] * 3
        L.orc_distance_batch.restype = None
        L.orc_distance_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_silhouette_distance_batch.restype = None
        L.orc_silhouette_distance_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_intersect_batch.restype = None
        L.orc_intersect_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        L.orc_field_eval_batch.restype = None
        L.orc_field_eval_batch.argtypes = [C.POINTER(_Field), C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        L.orc_sigma_prime.restype = C.c_float
        L.orc_sigma_prime.argtypes = [C.POINTER(_Params), C.c_float, C.c_float]
        for n in ("orc_i0", "orc_k0"):
            getattr(L, n).restype = C.c_double
            getattr(L, n).argtypes = [C.c_double]
        L.orc_screened_greens_norm.restype = C.c_double
        L.orc_screened_greens_norm.argtypes = [C.c_double, C.c_double]
        L.orc_screened_greens.restype = C.c_double
        L.orc_screened_greens.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_greens_cache_fill.restype = None
        L.orc_greens_cache_fill.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        L.orc_screened_cache_fill.restype = None
        L.orc_screened_cache_fill.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_void_p]
        L.orc_screened_icdf.restype = None
        L.orc_screened_icdf.argtypes = [C.c_double, C.c_int, C.c_void_p]
        L.orc_solve.restype = C.c_int
        L.orc_solve.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _lib = L
    return _lib


def _f32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PackedField:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, desc: dict):
        self.terms = np.ascontiguousarray(desc["terms"])
        assert self.terms.dtype.itemsize == C.sizeof(_Term)
        self.grid = None if desc["grid"] is None else _f32(desc["grid"])
        s = _Field()
        s.kind, s.n_terms, s.c0, s.mask_kind = int(desc["kind"]), len(self.terms), float(desc["c0"]), int(desc["mask_kind"])
        for i in range(4):
            s.mask[i] = float(desc["mask"][i])
        s.outside = float(desc["outside"])
        s.nx, s.ny, s.x0, s.y0, s.dx, s.dy = int(desc["nx"]), int(desc["ny"]), desc["x0"], desc["y0"], desc["dx"], desc["dy"]
        s.terms = self.terms.ctypes.data if len(self.terms) else None
        s.grid = self.grid.ctypes.data if self.grid is not None else None
        self.struct = s

    def ref(self):
        return C.pointer(self.struct)


def pack_field(field):
    if field is None:
        return None
    return PackedField(field.describe() if hasattr(field, "describe") else field)


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def distance(pts, q):
    pts, q = _f32(pts), _f32(q).reshape(-1, 2)
    out = np.empty(len(q), np.float32)
    lib().orc_distance_batch(_ptr(pts), len(pts), _ptr(q), len(q), _ptr(out))
    return out


def silhouette_distance(pts, q):
    pts, q = _f32(pts), _f32(q).reshape(-1, 2)
    out = np.empty(len(q), np.float32)
    lib().orc_silhouette_distance_batch(_ptr(pts), len(pts), _ptr(q), len(q), _ptr(out))
    return out


def is_silhouette(pts, p):
    pts, p = _f32(pts), _f32(p)
    mask = np.zeros(max(len(pts) - 2, 0), np.uint8)
    lib().orc_is_silhouette(_ptr(pts), len(pts), float(p[0]), float(p[1]), _ptr(mask))
    return mask.astype(bool)


def ray_intersection(pts, p, d):
    pts, p, d = _f32(pts), _f32(p), _f32(d)
    out = np.empty(len(pts) - 1, np.float32)
    lib().orc_ray_intersection(_ptr(pts), len(pts), float(p[0]), float(p[1]), float(d[0]), float(d[1]), _ptr(out))
    return out


def intersect(pts, q, d, r):
    """Batched intersect_polylines: returns (pt (B,2), normal (B,2), found (B,), seg (B,))."""
    pts, q, d = _f32(pts), _f32(q).reshape(-1, 2), _f32(d).reshape(-1, 2)
    r = np.ascontiguousarray(np.broadcast_to(_f32(r), (len(q),)))
    pt, nr = np.empty((len(q), 2), np.float32), np.empty((len(q), 2), np.float32)
    found, seg = np.empty(len(q), np.uint8), np.empty(len(q), np.int32)
    lib().orc_intersect_batch(_ptr(pts), len(pts), _ptr(q), _ptr(d), _ptr(r), len(q), _ptr(pt), _ptr(nr), _ptr(found), _ptr(seg))
    return pt, nr, found.astype(bool), seg


def field_eval(field, q, derivs=False):
    pf, q = pack_field(field), _f32(q).reshape(-1, 2)
    v = np.empty(len(q), np.float32)
    if not derivs:
        lib().orc_field_eval_batch(pf.ref(), _ptr(q), len(q), _ptr(v), None, None, None)
        return v
    gx, gy, lap = (np.empty(len(q), np.float32) for _ in range(3))
    lib().orc_field_eval_batch(pf.ref(), _ptr(q), len(q), _ptr(v), _ptr(gx), _ptr(gy), _ptr(lap))
    return v, gx, gy, lap


def i0(z):
    return lib().orc_i0(float(z))


def k0(z):
    return lib().orc_k0(float(z))


def screened_greens_norm(R, sb):
    return lib().orc_screened_greens_norm(float(R), float(sb))


def screened_greens(r, R, sb):
    return lib().orc_screened_greens(float(r), float(R), float(sb))


def greens_cache(seed_numpy, n=10000):
    out = np.empty(n, np.float64)
    lib().orc_greens_cache_fill(int(seed_numpy), n, _ptr(out))
    return out


def screened_cache(seed_numpy, sigma_bar, n=10000):
    out = np.empty(n, np.float64)
    lib().orc_screened_cache_fill(int(seed_numpy), float(sigma_bar), n, _ptr(out))
    return out


def screened_icdf(sigma_bar, n=1024):
    out = np.empty(n, np.float32)
    lib().orc_screened_icdf(float(sigma_bar), n, _ptr(out))
    return out


def philox(c, k):
    out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(*[int(x) for x in c], *[int(x) for x in k], _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# the solver
# ---------------------------------------------------------------------------------------------
class Problem:
    """Scene + fields + walk parameters, packed for ``orc_solve``."""

    def __init__(self, dirichlet, neumann=None, g=None, f=None, alpha=None, sigma=None, sigma_prime=None,
                 sigma_bar=0.0, sp_mode=SP_FULL, delta=None, majorant=None):
        self.majorant = majorant          # physical mode: dict(data, levels, x0, y0, dx, dy), see include/wost.h
        self.dir = _f32(dirichlet)
        self.neu = None if neumann is None else _f32(neumann)
        self.fields = {k: pack_field(v) for k, v in dict(g=g, f=f, alpha=alpha, sigma=sigma, sigma_prime=sigma_prime).items()}
        self.sigma_bar = float(sigma_bar)
        self.sp_mode = SP_FIELD if sigma_prime is not None else sp_mode
        self.delta = (alpha is not None or sigma is not None) if delta is None else bool(delta)
        self._icdf = None

    @classmethod
    def from_scenario(cls, s, sigma_bar=None, majorant=None):
        """(pass ``compat=s.compat`` to :meth:`solve` for the physical-mode scenes)"""
        sb = sigma_bar if sigma_bar is not None else (s.sigma_bar or 0.0)
        return cls(s.dirichlet, s.neumann, g=s.g, f=s.f, alpha=s.alpha, sigma=s.sigma, sigma_bar=sb, sp_mode=s.sp_mode,
                   majorant=majorant)

    def params(self, n_walks, max_steps, eps, rng_mode, seed, seed_numpy=0, point_index_base=0, walk_offset=0,
               icdf=None, n_threads=0, torch_trig=False, compat="reference"):
        p = _Params()
        p.dir_pts, p.n_dir = self.dir.ctypes.data, len(self.dir)
        p.neu_pts, p.n_neu = (self.neu.ctypes.data, len(self.neu)) if self.neu is not None else (None, 0)
        for k, pf in self.fields.items():
            if pf is not None:
                setattr(p, k, pf.ref())
        p.delta, p.sp_mode, p.sigma_bar = int(self.delta), int(self.sp_mode), self.sigma_bar
        p.n_walks, p.max_steps, p.eps = int(n_walks), int(max_steps), float(eps)
        p.rng_mode, p.seed, p.seed_numpy = int(rng_mode), int(seed), int(seed_numpy)
        p.point_index_base, p.walk_offset = int(point_index_base), int(walk_offset)
        if self.delta and rng_mode == RNG_PHILOX:
            if icdf is None:
                if self._icdf is None:
                    self._icdf = screened_icdf(self.sigma_bar, 1024)
                icdf = self._icdf
            self._icdf_live = _f32(icdf)
            p.icdf, p.icdf_len = self._icdf_live.ctypes.data, len(self._icdf_live)
        p.n_threads = int(n_threads)
        p.compat_mode = {"reference": 0, "physical": 1}[compat]
        if self.majorant is not None and compat == "physical":
            m = self.majorant
            self._maj_live = _f32(m["data"])
            p.majorant, p.maj_levels = self._maj_live.ctypes.data, int(m["levels"])
            p.maj_x0, p.maj_y0, p.maj_dx, p.maj_dy = float(m["x0"]), float(m["y0"]), float(m["dx"]), float(m["dy"])
        scale = float(np.abs(self.dir).max()) if self.neu is None else float(max(np.abs(self.dir).max(), np.abs(self.neu).max()))
        p.phys_nudge = np.float32(1e-5 * scale)
        if torch_trig:
            self._trig = torch_trig_callbacks()
            p.sincos_fn, p.atan2_fn = C.cast(self._trig[0], C.c_void_p), C.cast(self._trig[1], C.c_void_p)
        return p

    def sigma_prime(self, x, y):
        p = self.params(1, 1, 1e-4, RNG_MT, 0)
        return lib().orc_sigma_prime(C.byref(p), float(x), float(y))

    def solve(self, points, n_walks, max_steps=1000, eps=1e-4, rng_mode=RNG_PHILOX, seed=42, seed_numpy=42,
              point_index_base=0, walk_offset=0, icdf=None, n_threads=0, walk_vals=False, walk_steps=False,
              n_trace=0, trace_cap=0, torch_trig=False, compat="reference"):
        pts = _f32(points).reshape(-1, 2)
        P = len(pts)
        p = self.params(n_walks, max_steps, eps, rng_mode, seed, seed_numpy, point_index_base, walk_offset, icdf, n_threads, torch_trig, compat)
        mean, m2 = np.zeros(P, np.float64), np.zeros(P, np.float64)
        vals = np.zeros((P, n_walks), np.float32) if walk_vals else None
        wst = np.zeros((P, n_walks), np.int32) if walk_steps else None
        steps = C.c_int64(0)
        trace = np.zeros((n_trace, trace_cap, 4), np.float32) if n_trace else None
        tlen = np.zeros(n_trace, np.int32) if n_trace else None
        rc = lib().orc_solve(C.byref(p), _ptr(pts), P, _ptr(mean), _ptr(m2), _ptr(vals), C.addressof(steps), _ptr(wst),
                             n_trace, trace_cap, _ptr(trace), _ptr(tlen))
        assert rc == 0
        out = dict(mean=mean, m2=m2, steps=steps.value, n=n_walks)
        out["stderr"] = np.sqrt(m2 / max(n_walks - 1, 1) / n_walks)
        if walk_vals:
            out["walk_vals"] = vals
        if walk_steps:
            out["walk_steps"] = wst
        if n_trace:
            out["trace"], out["trace_len"] = trace, tlen
        return out
