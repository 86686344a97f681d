"""Generate the golden fixtures under tests/golden/ by importing the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only); the fixtures it writes are
committed so that the oracle can be pinned anywhere (`/root/reference` does not exist on the GPU box).

    python oracle/gen_golden.py [--only geometry|samplers|sigma|walks] [--ref /root/reference]

What is recorded (all produced by the reference's own code paths):
  geometry.npz   the 5 known-answer tests of geometry/PolylinesSimple.py:309-357 plus randomized
                 queries of every primitive on four scenes
  samplers.npz   the Green's / screened-Green's radius caches for np.random.seed(42)
                 (solvers/utils.py:138-151,181-195), scipy i0/k0 samples, screenedGreensNorm2D values
  sigma.npz      sigma' (solvers/WoStSolver.py:88-127) and sigma_bar (:130-136) for the
                 delta-tracking scenarios
  walks_<cfg>.npz  WostSolver_2D.solve(..., return_history=True) with torch.manual_seed(42),
                 np.random.seed(42): per-walk totals, step counts, the (P,1) estimate and the
                 full paths of the first walks — the oracle replays these in ORC_RNG_MT mode
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


def import_reference(ref_root: str):
    """The reference imports matplotlib at module scope (utils.py:7-8); it is absent here, stub it."""
    stub = Path(tempfile.mkdtemp(prefix="mplstub_")) / "matplotlib"
    stub.mkdir()
    (stub / "__init__.py").write_text("")
    (stub / "pyplot.py").write_text("Figure = object\n")
    (stub / "patches.py").write_text("Circle = object\n")
    sys.path.insert(0, str(stub.parent))
    sys.path.insert(0, ref_root)
    warnings.filterwarnings("ignore")
    os.environ["TQDM_DISABLE"] = "1"
    from geometry.PolylinesSimple import PolyLinesSimple  # noqa
    from solvers.WoStSolver import WostSolver_2D  # noqa
    import solvers.utils as sutils  # noqa

    return PolyLinesSimple, WostSolver_2D, sutils


def gen_geometry(PolyLinesSimple):
    out = {}
    # --- the reference's own known-answer tests (PolylinesSimple.py:309-357) -----------------
    sq = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    tent = torch.tensor([[0.0, 0.0], [1.0, 1.0], [2.0, 0.0]])
    out["kat_square"], out["kat_tent"] = sq.numpy(), tent.numpy()
    out["kat_distance"] = PolyLinesSimple(sq).distance(torch.tensor([0.5, 0.5])).numpy()
    out["kat_is_silhouette"] = PolyLinesSimple(tent).isSilhouette(torch.tensor([1.5, 0.6])).numpy()
    out["kat_silhouette_distance"] = PolyLinesSimple(tent).silhouetteDistance(torch.tensor([1.5, 0.6])).numpy()
    out["kat_ray"] = PolyLinesSimple(sq).rayIntersection(torch.tensor([0.5, 0.5]), torch.tensor([1.0, 0.0])).numpy()
    pt, nr, found = PolyLinesSimple(sq).intersectPolylines(torch.tensor([0.5, 0.5]), torch.tensor([1.0, 0.0]), 2.0)
    out["kat_intersect_pt"], out["kat_intersect_nrm"], out["kat_intersect_found"] = pt.numpy(), nr.numpy(), np.array(found)

    # --- randomized differential vectors -------------------------------------------------------
    sys.path.insert(0, str(ROOT))
    from dcrmontecarlo_b200 import scenarios as sc

    g = torch.Generator().manual_seed(1234)
    x = torch.arange(0, 12.0, 0.25)
    topo = torch.stack((x, 0.6 * torch.sin(0.9 * x) + 0.2 * torch.cos(2.3 * x)), dim=-1)   # funcToPolyline-style open line
    scenes = {"square2": (sc.square(2.0), 2.2), "circle05": (sc.circle(0.5, 32), 2.0), "tent": (tent, 2.5),
              "topo": (topo, 12.0), "edge": (torch.tensor([[-100.0, 100.0], [100.0, 100.0]]), 110.0)}
    B = 1500
    for name, (pts, lim) in scenes.items():
        poly = PolyLinesSimple(pts)
        q = (torch.rand(B, 2, generator=g) * 2 - 1) * lim
        if name == "topo":
            q[:, 0] = q[:, 0].abs(); q[:, 1] = q[:, 1] * 0.2
        th = torch.rand(B, generator=g) * 2 * np.pi
        d = torch.stack([torch.cos(th), torch.sin(th)], dim=1)
        r = torch.rand(B, generator=g) * lim * 0.75 + 1e-3
        # a third of the queries start ON the polyline (like a reflected walker does)
        k = torch.randint(0, len(pts) - 1, (B,), generator=g)
        w = torch.rand(B, generator=g)
        on = pts[k] * (1 - w[:, None]) + pts[k + 1] * w[:, None]
        q[::3] = on[::3]
        nseg = len(pts) - 1
        dist, sild = np.empty(B, np.float32), np.empty(B, np.float32)
        silm = np.zeros((B, max(len(pts) - 2, 0)), bool)
        ray = np.empty((B, nseg), np.float32)
        ipt, inr = np.empty((B, 2), np.float32), np.empty((B, 2), np.float32)
        ifound, iseg = np.empty(B, bool), np.full(B, -1, np.int32)
        for i in range(B):
            dist[i] = poly.distance(q[i]).item()
            sild[i] = poly.silhouetteDistance(q[i]).item()
            silm[i] = poly.isSilhouette(q[i]).numpy()
            ray[i] = poly.rayIntersection(q[i], d[i]).numpy()
            p_, n_, f_ = poly.intersectPolylines(q[i], d[i], r[i].item())
            ipt[i], inr[i], ifound[i] = p_.numpy(), n_.numpy(), bool(f_)
            if f_:
                # the segment the reference picked: first index attaining the minimum (PolylinesSimple.py:177-178)
                unit = d[i] / torch.norm(d[i])
                s = poly.rayIntersection(q[i] + 1e-6 * unit, unit)
                iseg[i] = int(torch.where(torch.isfinite(s) & (s == s[torch.isfinite(s)].min()))[0][0])
        out.update({f"{name}_pts": pts.numpy(), f"{name}_q": q.numpy(), f"{name}_d": d.numpy(), f"{name}_r": r.numpy(),
                    f"{name}_distance": dist, f"{name}_sil_distance": sild, f"{name}_sil_mask": silm, f"{name}_ray": ray,
                    f"{name}_ipt": ipt, f"{name}_inrm": inr, f"{name}_ifound": ifound, f"{name}_iseg": iseg})
    np.savez_compressed(GOLDEN / "geometry.npz", **out)
    print("geometry.npz:", {k: v.shape for k, v in out.items() if k.endswith("_q")})


def gen_samplers(sutils):
    from scipy.special import i0, k0

    out = {}
    np.random.seed(42)
    s = sutils.GreensDistribution2D(10000); s._refill_cache()
    out["greens_cache_seed42"] = np.array(s.cache, np.float64)
    for sb in (2.40625, 10.0):
        np.random.seed(42)
        s = sutils.ScreenedGreensDistribution2D(sb, 4000); s._refill_cache()
        out[f"screened_cache_seed42_sb{sb}"] = np.array(s.cache, np.float64)
    z = np.concatenate([np.logspace(-6, 0, 40), np.linspace(1.0, 35.0, 60)])
    out["bessel_z"], out["bessel_i0"], out["bessel_k0"] = z, i0(z), k0(z)
    R = np.logspace(-5, 2, 50)
    out["norm_R"] = R
    for sb in (2.40625, 3.2175, 10.0):
        out[f"norm_sb{sb}"] = np.array([sutils.screenedGreensNorm2D(float(r), sb) for r in R])
        rr = np.linspace(1e-4, 0.999, 64)
        out[f"greens_r_sb{sb}"] = rr
        out[f"greens_sb{sb}"] = np.array([float(sutils.screenedGreens2D(torch.zeros(2), torch.tensor([r, 0.0]), 1.0, sb)) for r in rr])
    np.savez_compressed(GOLDEN / "samplers.npz", **out)
    print("samplers.npz written")


def _detached(field):
    """Mimic the reference test's callables that wrap results in torch.tensor(...)."""
    return lambda p: field(p).detach().clone()


def build_reference_solver(WostSolver_2D, PolyLinesSimple, s, detach_coeffs=False):
    alpha, sigma = s.alpha, s.sigma
    if detach_coeffs:
        alpha, sigma = (_detached(alpha) if alpha is not None else None), (_detached(sigma) if sigma is not None else None)
    with contextlib.redirect_stdout(io.StringIO()):
        solver = WostSolver_2D(
            dirichletBoundary=PolyLinesSimple(s.dirichlet), dirichletBoundaryFunction=s.g,
            neumannBoundary=PolyLinesSimple(s.neumann) if s.neumann is not None else None,
            source=s.f, sigma=sigma, alpha=alpha)
    return solver


def gen_sigma(PolyLinesSimple, WostSolver_2D):
    from dcrmontecarlo_b200 import scenarios as sc

    out = {}
    g = torch.Generator().manual_seed(7)
    for key, detach in (("cfg1b", False), ("cfg4", True), ("cfg5", False)):
        s = sc.ALL[key]()
        solver = build_reference_solver(WostSolver_2D, PolyLinesSimple, s, detach)
        (xmin, xmax), (ymin, ymax) = [[float(a), float(b)] for a, b in solver.domain_bounds]
        q = torch.rand(300, 2, generator=g) * torch.tensor([xmax - xmin, ymax - ymin]) + torch.tensor([xmin, ymin])
        with contextlib.redirect_stdout(io.StringIO()):
            sp = np.array([float(solver.sigma_prime(p)) for p in q], np.float32)
        out[f"{key}_q"], out[f"{key}_sigma_prime"], out[f"{key}_sigma_bar"] = q.numpy(), sp, np.float64(solver.sigma_bar)
        print(key, "sigma_bar", solver.sigma_bar)
    np.savez_compressed(GOLDEN / "sigma.npz", **out)


def run_reference_walks(solver, pts, n_walks, max_steps, eps, seed=42, n_trace=48, trace_cap=96):
    torch.manual_seed(seed); np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        est, hist = solver.solve(pts, nWalks=n_walks, maxSteps=max_steps, eps=eps, return_history=True)
    P = len(pts)
    vals, steps = np.zeros((P, n_walks), np.float64), np.zeros((P, n_walks), np.int32)
    trace = np.full((n_trace, trace_cap, 4), np.nan, np.float32); tlen = np.zeros(n_trace, np.int32)
    for pi in range(P):
        prev = 0.0
        for w, wh in enumerate(hist[pi]):
            vals[pi, w] = wh["total_contribution"] - prev; prev = wh["total_contribution"]   # WoStSolver.py:308
            steps[pi, w] = len(wh["path"])
            flat = pi * n_walks + w
            if flat < n_trace:
                tlen[flat] = min(len(wh["path"]), trace_cap)
                for k, st in enumerate(wh["path"][:trace_cap]):
                    dn = st["neumann_distance"]
                    trace[flat, k] = (st["point"][0].item(), st["point"][1].item(), st["dirichlet_distance"],
                                      np.inf if dn is None else dn)
    return dict(estimate=est.detach().numpy().astype(np.float32), walk_vals=vals, walk_steps=steps, trace=trace, trace_len=tlen)


def gen_walks(PolyLinesSimple, WostSolver_2D, only=None):
    from dcrmontecarlo_b200 import scenarios as sc

    # (scenario key, evaluation-point subset, walks) sized so the whole script stays within minutes
    plan = {"cfg1a": (slice(None), 150), "cfg1b": (slice(None), 40), "cfg2": (slice(0, 404, 9), 100),
            "cfg3": (slice(0, 404, 9), 100), "cfg4": (slice(0, 648, 18), 25), "cfg5": (slice(0, 9, 2), 150)}
    for key, (sub, W) in plan.items():
        if only and key not in only:
            continue
        s = sc.ALL[key]()
        solver = build_reference_solver(WostSolver_2D, PolyLinesSimple, s, detach_coeffs=(key == "cfg4"))
        pts = s.points[sub].contiguous()
        r = run_reference_walks(solver, pts, W, s.max_steps, s.eps)
        r.update(points=pts.numpy(), n_walks=np.int64(W), max_steps=np.int64(s.max_steps), eps=np.float64(s.eps),
                 sigma_bar=np.float64(getattr(solver, "sigma_bar", 0.0)), seed=np.int64(42))
        np.savez_compressed(GOLDEN / f"walks_{key}.npz", **r)
        print(key, "P", len(pts), "W", W, "steps/walk", r["walk_steps"].mean(), "mean est", float(r["estimate"].mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    sys.path.insert(0, str(ROOT))
    PolyLinesSimple, WostSolver_2D, sutils = import_reference(a.ref)
    only = set(a.only) if a.only else None
    walk_keys = {k for k in (only or ()) if k.startswith("cfg")}
    if not only or "geometry" in only:
        gen_geometry(PolyLinesSimple)
    if not only or "samplers" in only:
        gen_samplers(sutils)
    if not only or "sigma" in only:
        gen_sigma(PolyLinesSimple, WostSolver_2D)
    if not only or "walks" in only or walk_keys:
        gen_walks(PolyLinesSimple, WostSolver_2D, walk_keys or None)


if __name__ == "__main__":
    main()
