"""ctypes binding of libwost.so (include/wost.h).  PyTorch only hands over pointers and streams.

There is no CPU fallback: if the library is missing, or no CUDA device is present when a compute call
is made, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from pathlib import Path

import numpy as np
import torch

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("WOST_LIB", PKG / "libwost.so"))   # WOST_LIB: alternative build (A/B experiments)

WALK_BLOCK = 1024
SP_FULL, SP_RATIO, SP_FIELD = 0, 1, 2
COMPAT = {"reference": 0, "physical": 1}


class WostError(RuntimeError):
    pass


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("px", C.c_int32), ("py", C.c_int32), ("t1", C.c_int32), ("t2", C.c_int32),
                ("A", C.c_float), ("q", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("R", C.c_float),
                ("w1x", C.c_float), ("w1y", C.c_float), ("p1", C.c_float),
                ("w2x", C.c_float), ("w2y", C.c_float), ("p2", C.c_float)]


class FieldDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_terms", C.c_int32), ("c0", C.c_float), ("mask_kind", C.c_int32),
                ("mask", C.c_float * 4), ("outside", C.c_float), ("nx", C.c_int32), ("ny", C.c_int32),
                ("x0", C.c_float), ("y0", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
                ("terms", C.c_void_p), ("grid", C.c_void_p)]


class Fields(C.Structure):
    _fields_ = [("g", C.c_void_p), ("f", C.c_void_p), ("alpha", C.c_void_p), ("sigma", C.c_void_p), ("sigma_prime", C.c_void_p)]


class SolveParams(C.Structure):
    _fields_ = [("n_walks", C.c_int64), ("max_steps", C.c_int32), ("eps", C.c_float), ("delta_tracking", C.c_int32),
                ("sp_mode", C.c_int32), ("sigma_bar", C.c_float), ("screened_icdf", C.c_void_p), ("icdf_len", C.c_int32),
                ("seed", C.c_uint64), ("point_index_base", C.c_int64), ("walk_offset", C.c_int64), ("compat_mode", C.c_int32),
                ("majorant_levels", C.c_int32), ("majorant", C.c_void_p),
                ("majorant_x0", C.c_float), ("majorant_y0", C.c_float), ("majorant_dx", C.c_float), ("majorant_dy", C.c_float),
                ("jit", C.c_int32), ("point_index_stride", C.c_int64)]


EXPORTS = {
    # name: (restype, argtypes)
    "wost_version": (C.c_int, []),
    "wost_last_error": (C.c_char_p, []),
    "wost_device_count": (C.c_int, []),
    "wost_scene_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "wost_scene_destroy": (C.c_int, [C.c_void_p]),
    "wost_scene_trim": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "wost_selftest_division": (C.c_int, [C.c_int32, C.c_int64, C.c_uint64, C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "wost_field_create": (C.c_int, [C.POINTER(FieldDesc), C.c_int32, C.POINTER(C.c_void_p)]),
    "wost_field_destroy": (C.c_int, [C.c_void_p]),
    "wost_field_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_sigma_prime_eval": (C.c_int, [C.POINTER(Fields), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "wost_solve": (C.c_int, [C.c_void_p, C.POINTER(Fields), C.POINTER(SolveParams), C.c_void_p, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_solve_multi_source": (C.c_int, [C.c_void_p, C.POINTER(Fields), C.POINTER(C.c_void_p), C.c_int32, C.POINTER(SolveParams),
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_merge_block_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_geom_distance": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_geom_silhouette": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_geom_ray": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "wost_geom_intersect": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wost_fp32_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "wost_jit_stats": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "wost_jit_last_note": (C.c_char_p, []),
    "wost_jit_offline": (C.c_int, [C.POINTER(C.POINTER(FieldDesc))] + [C.c_int32] * 11 + [C.c_char_p, C.c_char_p]),
}
JIT = {"auto": 0, "on": 1, "off": 2}

_lib = None


def lib():
    """Load libwost.so; fails loudly if it has not been built (python -m dcrmontecarlo_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise WostError(f"{LIB_PATH} is missing: build the CUDA extension with `python -m dcrmontecarlo_b200.build` "
                            "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise WostError(f"libwost error {rc}: {lib().wost_last_error().decode(errors='replace')}")


def require_cuda() -> int:
    n = lib().wost_device_count()
    if n <= 0:
        raise WostError("no CUDA device available: the Walk-on-Stars kernels need a GPU (there is no CPU fallback)")
    return n


def current_device() -> int:
    require_cuda()
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def current_stream(device: int):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream) if torch.cuda.is_available() else C.c_void_p(0)


def ptr(t):
    """data pointer of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def host_f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Scene:
    """wost_scene_t handle for a (Dirichlet, Neumann) polyline pair on one device."""

    def __init__(self, dirichlet, neumann=None, device: int | None = None):
        self.device = current_device() if device is None else int(device)
        d = host_f32(dirichlet).reshape(-1, 2)
        n = None if neumann is None else host_f32(neumann).reshape(-1, 2)
        h = C.c_void_p(0)
        check(lib().wost_scene_create(ptr(d), len(d), ptr(n), 0 if n is None else len(n), self.device, C.byref(h)))
        self.handle = h
        self.n_dirichlet, self.n_neumann = len(d), 0 if n is None else len(n)
        self._fin = weakref.finalize(self, lib().wost_scene_destroy, h)

    def trim(self) -> int:
        """Return the scene's cached device scratch to the driver; returns the number of bytes released."""
        n = C.c_int64(0)
        check(lib().wost_scene_trim(self.handle, C.byref(n)))
        return n.value


def field_desc(field):
    """wost_field_desc_t of a fields.Field, plus the arrays it points into (keep them alive while the struct is used)."""
    desc = field.describe()
    terms = np.ascontiguousarray(desc["terms"])
    assert terms.dtype.itemsize == C.sizeof(_Term), "term layout mismatch with wost_term_t"
    grid = None if desc["grid"] is None else host_f32(desc["grid"])
    d = FieldDesc()
    d.kind, d.n_terms, d.c0, d.mask_kind = int(desc["kind"]), len(terms), float(desc["c0"]), int(desc["mask_kind"])
    for i in range(4):
        d.mask[i] = float(desc["mask"][i])
    d.outside = float(desc["outside"])
    d.nx, d.ny, d.x0, d.y0, d.dx, d.dy = int(desc["nx"]), int(desc["ny"]), desc["x0"], desc["y0"], desc["dx"], desc["dy"]
    d.terms = terms.ctypes.data if len(terms) else None
    d.grid = grid.ctypes.data if grid is not None else None
    return d, terms, grid


class DeviceField:
    """wost_field_t handle built from a fields.Field description."""

    def __init__(self, field, device: int):
        d, self._terms, self._grid = field_desc(field)
        h = C.c_void_p(0)
        check(lib().wost_field_create(C.byref(d), int(device), C.byref(h)))
        self.handle, self.device = h, int(device)
        self._fin = weakref.finalize(self, lib().wost_field_destroy, h)

    def eval(self, pts, derivs: bool = False):
        p = host_f32(pts).reshape(-1, 2)
        B = len(p)
        v = np.empty(B, np.float32)
        if not derivs:
            check(lib().wost_field_eval(self.handle, ptr(p), B, ptr(v), None, None, None, current_stream(self.device)))
            return v
        gx, gy, lap = (np.empty(B, np.float32) for _ in range(3))
        check(lib().wost_field_eval(self.handle, ptr(p), B, ptr(v), ptr(gx), ptr(gy), ptr(lap), current_stream(self.device)))
        return v, gx, gy, lap


def fields_struct(g=None, f=None, alpha=None, sigma=None, sigma_prime=None) -> Fields:
    F = Fields()
    for k, v in dict(g=g, f=f, alpha=alpha, sigma=sigma, sigma_prime=sigma_prime).items():
        setattr(F, k, v.handle if v is not None else None)
    return F


def sigma_prime_eval(fields: Fields, sp_mode: int, pts, device: int):
    p = host_f32(pts).reshape(-1, 2)
    out = np.empty(len(p), np.float32)
    check(lib().wost_sigma_prime_eval(C.byref(fields), int(sp_mode), ptr(p), len(p), ptr(out), current_stream(device)))
    return out


def _set_majorant(prm, majorant):
    """majorant: None or dict(data=float32 pyramid (numpy or CUDA tensor), levels, x0, y0, dx, dy) -- see include/wost.h."""
    if majorant is None:
        return None
    data = majorant["data"]
    keep = data if (isinstance(data, torch.Tensor) and data.is_cuda) else host_f32(data)
    prm.majorant, prm.majorant_levels = ptr(keep).value, int(majorant["levels"])
    prm.majorant_x0, prm.majorant_y0 = float(majorant["x0"]), float(majorant["y0"])
    prm.majorant_dx, prm.majorant_dy = float(majorant["dx"]), float(majorant["dy"])
    return keep


def solve(scene: Scene, fields: Fields, pts, n_walks: int, max_steps: int, eps: float, *, delta: bool = False,
          sp_mode: int = SP_FULL, sigma_bar: float = 0.0, icdf=None, seed: int = 0, point_index_base: int = 0,
          walk_offset: int = 0, point_index_stride: int = 1, want_block_stats: bool = False, want_walk_vals: bool = False, n_trace: int = 0,
          trace_cap: int = 0, device_outputs: bool = False, compat: str = "reference", majorant=None, jit: str = "auto",
          out: dict | None = None):
    """One wost_solve call.  ``pts`` may be a host array/tensor or a CUDA tensor on the scene's device.
    With ``device_outputs`` the results stay on the device as torch tensors (stream-ordered, no sync); ``out`` may then
    hold preallocated contiguous CUDA tensors ``mean`` / ``m2`` (P,) float64, ``steps`` (1,) int64, ``block_stats``
    (P, nblk, 2) float64 to write into (e.g. slices of a gather buffer)."""
    dev = scene.device
    if isinstance(pts, torch.Tensor) and pts.is_cuda:
        p = pts.detach().to(torch.float32).contiguous().reshape(-1, 2)
    else:
        p = host_f32(pts).reshape(-1, 2)
    P = int(p.shape[0])
    nblk = (n_walks + WALK_BLOCK - 1) // WALK_BLOCK
    prm = SolveParams()
    prm.n_walks, prm.max_steps, prm.eps = int(n_walks), int(max_steps), float(eps)
    prm.delta_tracking, prm.sp_mode, prm.sigma_bar = int(bool(delta)), int(sp_mode), float(sigma_bar)
    icdf_keep = None
    if delta and icdf is not None:
        icdf_keep = icdf if (isinstance(icdf, torch.Tensor) and icdf.is_cuda) else host_f32(icdf)
        prm.screened_icdf, prm.icdf_len = ptr(icdf_keep).value, int(icdf_keep.shape[0])
    prm.seed, prm.point_index_base, prm.walk_offset = int(seed) & (2 ** 64 - 1), int(point_index_base), int(walk_offset)
    prm.compat_mode = COMPAT[compat]
    prm.jit, prm.point_index_stride = JIT[jit], int(point_index_stride)
    maj_keep = _set_majorant(prm, majorant)

    if device_outputs:
        tdev = torch.device("cuda", dev)
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=tdev)
        o = out or {}
        mean = o["mean"] if "mean" in o else mk((P,), torch.float64)
        m2 = o["m2"] if "m2" in o else mk((P,), torch.float64)
        blk = (o["block_stats"] if "block_stats" in o else mk((P, nblk, 2), torch.float64)) if want_block_stats else None
        vals = mk((P, n_walks), torch.float32) if want_walk_vals else None
        steps = o["steps"] if "steps" in o else mk((1,), torch.int64)
        for t, n in ((mean, P), (m2, P), (steps, 1)) + (((blk, P * nblk * 2),) if blk is not None else ()):
            assert t.is_cuda and t.is_contiguous() and t.numel() == n and t.element_size() == 8, "preallocated outputs: contiguous 8-byte CUDA tensors"
        trace = mk((n_trace, trace_cap + 1, 8), torch.float32) if n_trace else None
        tlen = mk((n_trace,), torch.int32) if n_trace else None
    else:
        mean, m2 = np.empty(P, np.float64), np.empty(P, np.float64)
        blk = np.empty((P, nblk, 2), np.float64) if want_block_stats else None
        vals = np.empty((P, n_walks), np.float32) if want_walk_vals else None
        steps = np.zeros(1, np.uint64)
        trace = np.empty((n_trace, trace_cap + 1, 8), np.float32) if n_trace else None
        tlen = np.empty(n_trace, np.int32) if n_trace else None
    check(lib().wost_solve(scene.handle, C.byref(fields), C.byref(prm), ptr(p), P, ptr(mean), ptr(m2), ptr(blk), ptr(vals),
                           ptr(steps), int(n_trace), int(trace_cap), ptr(trace), ptr(tlen), current_stream(dev)))
    out = dict(mean=mean, m2=m2, steps=steps, n=n_walks)
    if want_block_stats:
        out["block_stats"] = blk
    if want_walk_vals:
        out["walk_vals"] = vals
    if n_trace:
        out["trace"], out["trace_len"] = trace, tlen
    return out


def solve_multi_source(scene: Scene, fields: Fields, sources, pts, n_walks: int, max_steps: int, eps: float, *,
                       delta: bool = False, sp_mode: int = SP_FULL, sigma_bar: float = 0.0, icdf=None, seed: int = 0,
                       point_index_base: int = 0, walk_offset: int = 0, point_index_stride: int = 1, want_block_stats: bool = False,
                       device_outputs: bool = False, compat: str = "reference", majorant=None, jit: str = "auto"):
    """One wost_solve_multi_source call: shared walks, one estimate per (source, point).  ``sources`` is a list of
    DeviceField.  Returns mean / m2 of shape (S, P)."""
    dev = scene.device
    if isinstance(pts, torch.Tensor) and pts.is_cuda:
        p = pts.detach().to(torch.float32).contiguous().reshape(-1, 2)
    else:
        p = host_f32(pts).reshape(-1, 2)
    P, S = int(p.shape[0]), len(sources)
    nblk = (n_walks + WALK_BLOCK - 1) // WALK_BLOCK
    prm = SolveParams()
    prm.n_walks, prm.max_steps, prm.eps = int(n_walks), int(max_steps), float(eps)
    prm.delta_tracking, prm.sp_mode, prm.sigma_bar = int(bool(delta)), int(sp_mode), float(sigma_bar)
    icdf_keep = None
    if delta and icdf is not None:
        icdf_keep = icdf if (isinstance(icdf, torch.Tensor) and icdf.is_cuda) else host_f32(icdf)
        prm.screened_icdf, prm.icdf_len = ptr(icdf_keep).value, int(icdf_keep.shape[0])
    prm.seed, prm.point_index_base, prm.walk_offset = int(seed) & (2 ** 64 - 1), int(point_index_base), int(walk_offset)
    prm.compat_mode = COMPAT[compat]
    prm.jit, prm.point_index_stride = JIT[jit], int(point_index_stride)
    maj_keep = _set_majorant(prm, majorant)
    handles = (C.c_void_p * S)(*[s.handle for s in sources])
    if device_outputs:
        tdev = torch.device("cuda", dev)
        mean, m2 = torch.empty((S, P), dtype=torch.float64, device=tdev), torch.empty((S, P), dtype=torch.float64, device=tdev)
        blk = torch.empty((S, P, nblk, 2), dtype=torch.float64, device=tdev) if want_block_stats else None
        steps = torch.empty((1,), dtype=torch.int64, device=tdev)
    else:
        mean, m2 = np.empty((S, P), np.float64), np.empty((S, P), np.float64)
        blk = np.empty((S, P, nblk, 2), np.float64) if want_block_stats else None
        steps = np.zeros(1, np.uint64)
    check(lib().wost_solve_multi_source(scene.handle, C.byref(fields), handles, S, C.byref(prm), ptr(p), P, ptr(mean), ptr(m2),
                                        ptr(blk), ptr(steps), current_stream(dev)))
    out = dict(mean=mean, m2=m2, steps=steps, n=n_walks)
    if want_block_stats:
        out["block_stats"] = blk
    return out


def merge_block_stats(block_stats, n_walks: int, device: int):
    """Fixed-order merge of (P, nblk, 2) per-block (mean, M2) -> (mean[P], m2[P]) with the solver's own device code."""
    if isinstance(block_stats, torch.Tensor) and block_stats.is_cuda:
        b = block_stats.contiguous()
        P = b.shape[0]
        mean = torch.empty(P, dtype=torch.float64, device=b.device)
        m2 = torch.empty(P, dtype=torch.float64, device=b.device)
    else:
        b = np.ascontiguousarray(np.asarray(block_stats, np.float64))
        P = b.shape[0]
        mean, m2 = np.empty(P, np.float64), np.empty(P, np.float64)
    check(lib().wost_merge_block_stats(ptr(b), P, int(n_walks), int(device), ptr(mean), ptr(m2), current_stream(device)))
    return mean, m2


def jit_stats():
    """(kernels compiled by NVRTC, cache hits, solves that ran a specialised kernel)"""
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib().wost_jit_stats(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def jit_last_note() -> str:
    return lib().wost_jit_last_note().decode(errors="replace")


def jit_offline(fields: dict, *, neu, src, delta, trace=False, phys=False, big=False, multi=False, sp_mode=0, min_blocks=4,
                n_dseg=-1, n_nseg=-1, arch="sm_100a", prefix="wost_walk_jit"):
    """Developer diagnostic (no GPU needed): compile the specialised kernel for ``fields`` (dict g / f / alpha / sigma /
    sigma_prime -> fields.Field or None) and write ``prefix``.cu / .cubin."""
    keep, arr = [], (C.POINTER(FieldDesc) * 5)()
    for i, k in enumerate(("g", "f", "alpha", "sigma", "sigma_prime")):
        if fields.get(k) is not None:
            d = field_desc(fields[k]); keep.append(d)
            arr[i] = C.pointer(d[0])
    check(lib().wost_jit_offline(arr, int(neu), int(src), int(delta), int(trace), int(phys), int(big), int(multi), int(sp_mode),
                                 int(min_blocks), int(n_dseg), int(n_nseg), arch.encode(), str(prefix).encode()))
    return lib().wost_last_error().decode()


def fp32_peak(device: int = 0):
    tf, mhz = C.c_double(0), C.c_double(0)
    check(lib().wost_fp32_peak(int(device), C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value
