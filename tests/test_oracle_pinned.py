"""Pins the CPU oracle (oracle/wost_oracle.c) to the reference.

Fixtures under tests/golden/ were produced by importing the UNMODIFIED reference
(oracle/gen_golden.py).  Geometry primitives must agree bit for bit; the walk estimator is replayed
with the reference's own RNG streams (torch + numpy mt19937) and must take the very same walks.
"""
import numpy as np
import pytest

from dcrmontecarlo_b200 import scenarios as sc
from oracle import wost_oracle as orc

SCENES = ["square2", "circle05", "tent", "topo", "edge"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- the reference's own known-answer tests (geometry/PolylinesSimple.py:309-357) -------------------
def test_kat_distance(golden):
    G = golden["geometry"]
    d = orc.distance(G["kat_square"], [0.5, 0.5])[0]
    assert abs(d - 0.5) <= 1e-6 and d == G["kat_distance"]


def test_kat_silhouette(golden):
    G = golden["geometry"]
    assert orc.is_silhouette(G["kat_tent"], [1.5, 0.6]).tolist() == [True] == G["kat_is_silhouette"].tolist()
    d = orc.silhouette_distance(G["kat_tent"], [1.5, 0.6])[0]
    assert abs(d - np.hypot(0.5, 0.4)) <= 1e-6 and d == G["kat_silhouette_distance"]


def test_kat_ray(golden):
    G = golden["geometry"]
    s = orc.ray_intersection(G["kat_square"], [0.5, 0.5], [1.0, 0.0])
    assert np.array_equal(s, np.array([np.inf, 0.5, np.inf, np.inf], np.float32)) and np.array_equal(s, G["kat_ray"])


def test_kat_intersect(golden):
    G = golden["geometry"]
    pt, nr, found, seg = orc.intersect(G["kat_square"], [0.5, 0.5], [1.0, 0.0], 2.0)
    assert np.allclose(pt[0], [1.0, 0.5], atol=1e-6) and np.allclose(nr[0], [-1.0, 0.0], atol=1e-6) and found[0] and seg[0] == 1
    assert np.array_equal(bits(pt[0]), bits(G["kat_intersect_pt"])) and np.array_equal(nr[0], G["kat_intersect_nrm"])


# ---- randomized differential vectors: bit-exact ---------------------------------------------------
@pytest.mark.parametrize("scene", SCENES)
def test_geometry_bit_exact(golden, scene):
    G = golden["geometry"]
    pts, q, d, r = (G[f"{scene}_{k}"] for k in ("pts", "q", "d", "r"))
    assert np.array_equal(bits(orc.distance(pts, q)), bits(G[f"{scene}_distance"]))
    assert np.array_equal(bits(orc.silhouette_distance(pts, q)), bits(G[f"{scene}_sil_distance"]))
    for i in range(0, len(q), 7):
        assert np.array_equal(orc.is_silhouette(pts, q[i]), G[f"{scene}_sil_mask"][i])
        assert np.array_equal(bits(orc.ray_intersection(pts, q[i], d[i])), bits(G[f"{scene}_ray"][i]))
    pt, nr, found, seg = orc.intersect(pts, q, d, r)
    assert np.array_equal(found, G[f"{scene}_ifound"])
    assert np.array_equal(seg, G[f"{scene}_iseg"])                     # hit segment indices: exact
    assert np.array_equal(bits(pt), bits(G[f"{scene}_ipt"]))
    assert np.allclose(nr, G[f"{scene}_inrm"], atol=1e-7)


def test_zero_length_and_empty_cases():
    two = np.array([[0.0, 0.0], [1.0, 0.0]], np.float32)
    assert orc.silhouette_distance(two, [0.3, 0.4])[0] == np.inf       # 2-point polyline: no silhouette (Q4)
    assert orc.is_silhouette(two, [0.3, 0.4]).shape == (0,)
    pt, nr, found, seg = orc.intersect(two, [0.5, 1.0], [0.0, 0.0], 1.0)  # zero direction (:150-154)
    assert not found[0] and np.array_equal(pt[0], [0.5, 1.0]) and np.array_equal(nr[0], [1.0, 0.0])


# ---- Green's helpers and samplers -----------------------------------------------------------------
def test_bessel_against_scipy(golden):
    S = golden["samplers"]
    i0 = np.array([orc.i0(z) for z in S["bessel_z"]]); k0 = np.array([orc.k0(z) for z in S["bessel_z"]])
    assert np.allclose(i0, S["bessel_i0"], rtol=1e-13)
    assert np.allclose(k0, S["bessel_k0"], rtol=1e-12)


@pytest.mark.parametrize("sb", [2.40625, 3.2175, 10.0])
def test_greens_norm_and_function(golden, sb):
    S = golden["samplers"]
    n = np.array([orc.screened_greens_norm(R, sb) for R in S["norm_R"]])
    assert np.allclose(n, S[f"norm_sb{sb}"], rtol=1e-9, atol=1e-15)  # 1 - 1/I0 cancels in fp64 for tiny R
    g = np.array([orc.screened_greens(float(np.float32(r)), 1.0, sb) for r in S[f"greens_r_sb{sb}"]])
    assert np.allclose(g, S[f"greens_sb{sb}"], rtol=2e-6)             # the reference rounds r, I0, K0 to fp32


def test_sampler_caches_replay_numpy_stream(golden):
    S = golden["samplers"]
    assert np.array_equal(orc.greens_cache(42, 10000), S["greens_cache_seed42"])
    for sb in (2.40625, 10.0):
        ref = S[f"screened_cache_seed42_sb{sb}"]
        assert np.array_equal(orc.screened_cache(42, sb, len(ref)), ref)


def test_direct_samplers_match_cache_distributions(golden):
    """The cache-free samplers (product of uniforms; inverse-CDF table) draw from the distributions the
    reference's rejection caches realise: two-sample Kolmogorov-Smirnov."""
    from scipy.stats import ks_2samp

    S = golden["samplers"]
    rng = np.random.default_rng(0)
    direct = np.maximum(rng.random(40000) * rng.random(40000), 1e-6)
    assert ks_2samp(direct, S["greens_cache_seed42"]).pvalue > 1e-3
    assert abs(direct.mean() - 0.25) < 5e-3 and abs((direct ** 2).mean() - 1.0 / 9.0) < 5e-3   # SURVEY Q8 moments
    for sb in (2.40625, 10.0):
        tab = orc.screened_icdf(sb, 1024)
        assert np.all(np.diff(tab) >= 0) and tab[0] >= 1e-6 and tab[-1] <= 1.0 + 1e-6
        pos = rng.random(40000) * (len(tab) - 1); i = np.minimum(pos.astype(int), len(tab) - 2)
        draw = tab[i] + (pos - i) * (tab[i + 1] - tab[i])
        assert ks_2samp(draw, S[f"screened_cache_seed42_sb{sb}"]).pvalue > 1e-3


# ---- sigma' ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["cfg1b", "cfg4", "cfg5"])
def test_sigma_prime_closed_form(golden, key):
    G = golden["sigma"]
    s = sc.ALL[key]()
    prob = orc.Problem.from_scenario(s, sigma_bar=float(G[f"{key}_sigma_bar"]))
    got = np.array([prob.sigma_prime(x, y) for x, y in G[f"{key}_q"]], np.float32)
    ref = G[f"{key}_sigma_prime"]
    # cfg5's sigma' spans 12 orders of magnitude around the anomaly rims: compare relative to the local scale
    assert np.allclose(got, ref, rtol=2e-3, atol=2e-5 * max(1.0, float(np.abs(ref).max())))
    assert np.median(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-5


# ---- the walk: RNG replay --------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["cfg1a", "cfg1b", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_walk_replay(golden, key):
    """Same seeds, same streams => the oracle takes exactly the walks the reference took: identical step
    counts for every walk, identical paths, identical estimates (fp32 rounding of exp/sin in the fields aside)."""
    W = golden[f"walks_{key}"]
    s = sc.ALL[key]()
    prob = orc.Problem.from_scenario(s, sigma_bar=float(W["sigma_bar"]))
    nt, cap = W["trace"].shape[:2]
    r = prob.solve(W["points"], int(W["n_walks"]), int(W["max_steps"]), float(W["eps"]), rng_mode=orc.RNG_MT,
                   seed=int(W["seed"]), seed_numpy=int(W["seed"]), walk_vals=True, walk_steps=True,
                   n_trace=nt, trace_cap=cap, torch_trig=True)
    assert np.array_equal(r["walk_steps"], W["walk_steps"])
    assert r["steps"] == int(W["walk_steps"].sum())
    for i in range(nt):
        n = int(W["trace_len"][i])
        assert r["trace_len"][i] == n
        assert np.allclose(r["trace"][i, :n], W["trace"][i, :n], rtol=1e-6, atol=1e-7, equal_nan=True)
    # fixture per-walk values are differences of the reference's running fp32 total (WoStSolver.py:308)
    scale = 1.0 + np.abs(np.cumsum(W["walk_vals"], axis=1))
    assert np.all(np.abs(r["walk_vals"] - W["walk_vals"]) <= 2e-6 * scale + 1e-6 * np.abs(W["walk_vals"]))
    assert np.allclose(r["mean"], W["estimate"][:, 0], rtol=2e-6, atol=1e-7 * float(np.abs(W["estimate"]).max() + 1e-30))


def test_walk_replay_without_torch_trig_stays_close(golden):
    """With libm's cosf/sinf instead of torch's the walks differ by ulps that grow ~2x per step, so only the
    statistics are compared: the estimate stays within 3 combined standard errors."""
    W = golden["walks_cfg1a"]
    prob = orc.Problem.from_scenario(sc.cfg1a())
    r = prob.solve(W["points"], int(W["n_walks"]), int(W["max_steps"]), float(W["eps"]), rng_mode=orc.RNG_MT, seed=42, seed_numpy=42)
    se = r["stderr"]
    assert np.all(np.abs(r["mean"] - W["estimate"][:, 0]) <= 3 * np.sqrt(2) * se + 1e-6)


# ---- Philox mode (the stream the CUDA kernel uses) vs the replayed reference: statistics ------------
@pytest.mark.parametrize("key", ["cfg1a", "cfg1b", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_philox_mode_agrees_with_reference_statistically(golden, key):
    W = golden[f"walks_{key}"]
    s = sc.ALL[key]()
    prob = orc.Problem.from_scenario(s, sigma_bar=float(W["sigma_bar"]))
    nw = 4000 if key != "cfg5" else 1500
    r = prob.solve(W["points"], nw, int(W["max_steps"]), float(W["eps"]), rng_mode=orc.RNG_PHILOX, seed=7)
    n_ref = int(W["n_walks"])
    se_ref = W["walk_vals"].std(axis=1, ddof=1) / np.sqrt(n_ref)
    z = (r["mean"] - W["walk_vals"].mean(axis=1)) / np.sqrt(se_ref ** 2 + r["stderr"] ** 2 + 1e-30)
    assert np.mean(np.abs(z) <= 3.0) >= 0.95, z
    assert abs(np.mean(z)) < 4.0 / np.sqrt(len(z)) + 0.35                 # no systematic offset
    # walk lengths follow the same law
    assert abs(r["steps"] / (nw * len(W["points"])) / W["walk_steps"].mean() - 1.0) < 0.08


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert orc.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_analytic_solutions():
    for key, nw, tol in (("cfg1a", 20000, 3.5), ("cfg3", 20000, 3.5)):
        s = sc.ALL[key]()
        r = orc.Problem.from_scenario(s).solve(s.points[::5], nw, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=3)
        exact = s.analytic(s.points[::5]).numpy()
        z = (r["mean"] - exact) / (r["stderr"] + 1e-4)                   # 1e-4: the eps-shell bias of the estimator
        assert np.mean(np.abs(z) <= tol) >= 0.95, z


# ---- "physical" mode (textbook WoSt, not in the reference): analytic mixed-boundary solutions ---------------------
@pytest.mark.parametrize("key", ["phys_laplace", "phys_poisson", "phys_cylinder"])
def test_physical_mode_matches_analytic_mixed_boundary_solutions(key):
    s = sc.PHYSICAL[key]() if key != "phys_cylinder" else sc.phys_cylinder(64)
    nw = 30000
    r = orc.Problem.from_scenario(s).solve(s.points, nw, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=11, compat="physical")
    exact = s.analytic(s.points).numpy()
    slack = 2e-4 if key != "phys_cylinder" else 2e-3                    # eps shell; polygonal cylinder
    z = np.abs(r["mean"] - exact) / (r["stderr"] + slack)
    assert np.all(z <= 3.5), z
    # the reference's estimator does NOT solve these problems (walks leak through the Neumann wall)
    ref = orc.Problem.from_scenario(s).solve(s.points, nw, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=11, compat="reference")
    zr = np.abs(ref["mean"] - exact) / (ref["stderr"] + slack)
    assert not np.all(np.isfinite(zr)) or np.max(zr) > 5.0


def test_physical_delta_weights_against_scipy():
    """K1 quadrature, G_screened / G_laplace and the wall-hit weight 2 pi Q(t) I0(c) of the physical variable-coefficient walk."""
    import ctypes as C
    from scipy import special as sp

    L = orc.lib()
    for fn, n in ((L.orc_k1, 1), (L.orc_phys_green_ratio, 3), (L.orc_phys_wall_weight, 3)):
        fn.restype, fn.argtypes = C.c_double, [C.c_double] * n
    for z in (1e-5, 1e-2, 0.5, 1.0, 2.0, 7.0):
        assert abs(L.orc_k1(z) / sp.k1(z) - 1.0) < 1e-12
    rng = np.random.default_rng(0)
    for r, sb in ((0.5, 3.0), (0.05, 1.7), (1.0, 1.0)):
        c, s = r * np.sqrt(sb), np.sqrt(sb)
        rho = r * np.sqrt(rng.random(2000) * rng.random(2000))
        ratio = np.array([L.orc_phys_green_ratio(x, r, sb) for x in rho])
        exact = (sp.k0(rho * s) - sp.k0(c) / sp.i0(c) * sp.i0(rho * s)) / np.log(r / rho)
        assert np.allclose(ratio, exact, rtol=1e-9)
        assert np.all(ratio <= 1.0 + 1e-12) and np.all(ratio >= 1.0 / sp.i0(c) - 1e-9)      # screening only removes mass
        t = r * rng.random(200)
        w = np.array([L.orc_phys_wall_weight(x, r, sb) for x in t])
        assert np.allclose(w, t * s * (sp.k1(t * s) * sp.i0(c) + sp.k0(c) * sp.i1(t * s)), rtol=1e-9)
        assert np.all(w >= 1.0 - 1e-12) and np.all(w <= sp.i0(c) + 1e-12)
        assert abs(L.orc_phys_wall_weight(r, r, sb) - 1.0) < 1e-12                           # Wronskian: the sphere weighs 1
    # E_{rho ~ Laplace density}[ratio] = |G_screened| / |G_laplace|
    r, sb = 0.5, 3.0
    c = r * np.sqrt(sb)
    rho = r * np.sqrt(rng.random(200000) * rng.random(200000))
    m = np.mean([L.orc_phys_green_ratio(x, r, sb) for x in rho[:20000]])
    assert abs(m - (1 - 1 / sp.i0(c)) / (c * c / 4)) < 2e-3


@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("key", sorted(sc.PHYSICAL_VARCOEF))
def test_physical_mode_with_variable_coefficients_converges_to_analytic(key, local):
    """Delta tracking by the book converges where the reference's estimator has a bias floor (cfg 1b: RMSE 0.028) --
    with one majorant for the domain and with the spatially varying one (max-pyramid of |sigma'|)."""
    s = sc.PHYSICAL_VARCOEF[key]()
    solver = s.make_solver()                                             # host setup only: sigma', the majorant
    assert solver.use_delta_tracking and solver.sigma_bar > 0 and solver.majorant["levels"] == 9
    assert abs(solver.majorant["data"][-1] - solver.sigma_bar) < 1e-5 * solver.sigma_bar   # top of the pyramid = global majorant
    nw = 30000
    r = orc.Problem.from_scenario(s, sigma_bar=solver.sigma_bar, majorant=solver.majorant if local else None).solve(
        s.points, nw, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=11, compat="physical")
    exact = s.analytic(s.points).numpy()
    z = np.abs(r["mean"] - exact) / (r["stderr"] + 3e-4)
    assert np.all(z <= 3.6), z
    assert np.sqrt(np.mean((r["mean"] - exact) ** 2)) < 8e-3
    if key == "phys_varcoef_dirichlet":
        ref = sc.cfg1b()
        rr = orc.Problem.from_scenario(ref).solve(ref.points, nw, ref.max_steps, ref.eps, rng_mode=orc.RNG_PHILOX, seed=11)
        assert np.sqrt(np.mean((rr["mean"] - exact) ** 2)) > 0.02        # the reference's bias floor


def test_majorant_pyramid_bounds_sigma_prime_over_balls():
    """The value read for a ball dominates |sigma'| at points inside the ball, and the step radius satisfies r^2 M <= 1."""
    import ctypes as C

    s = sc.phys_dcr_halfspace()
    solver = s.make_solver()
    prob = orc.Problem.from_scenario(s, sigma_bar=solver.sigma_bar, majorant=solver.majorant)
    p = prob.params(1, 1, 1e-2, orc.RNG_PHILOX, 0, compat="physical")
    L = orc.lib()
    L.orc_majorant_over_ball.restype = C.c_float
    L.orc_majorant_over_ball.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_majorant_radius.restype = C.c_float
    L.orc_majorant_radius.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.orc_sigma_prime.restype = C.c_float
    L.orc_sigma_prime.argtypes = [C.c_void_p, C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    for _ in range(300):
        x, y = rng.uniform(-95, 95), rng.uniform(-95, -1)
        r = float(10 ** rng.uniform(-2, 1.9))
        M = L.orc_majorant_over_ball(C.byref(p), x, y, r)
        th, rho = rng.uniform(0, 2 * np.pi, 40), r * np.sqrt(rng.random(40))
        inside = [abs(L.orc_sigma_prime(C.byref(p), float(x + a * np.cos(t)), float(y + a * np.sin(t)))) for t, a in zip(th, rho)]
        assert M >= max(inside) * 0.999, (x, y, r, M, max(inside))
        Mr = C.c_float(0)
        rr = L.orc_majorant_radius(C.byref(p), x, y, r, 5e-3, C.byref(Mr))
        assert rr <= r * 1.000001 and (rr * rr * Mr.value <= 1.000001 or rr <= 5e-3)
    # far from the body sigma' vanishes: the step is not capped there, while the global majorant would cap it at 6.5 m
    Mr = C.c_float(0)
    assert L.orc_majorant_radius(C.byref(p), -80.0, -10.0, 10.0, 5e-3, C.byref(Mr)) == 10.0
    assert 1.0 / np.sqrt(solver.sigma_bar) < 7.0


def test_physical_dcr_halfspace_matches_finite_differences():
    """Half-space with a smooth conductive body, current dipole, insulating surface: no analytic solution, so the
    estimator (oracle side) is compared with a finite-difference solve of the same boundary value problem."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import fd_reference as fd

    s = sc.phys_dcr_halfspace()
    xs, ys, U = fd.solve_rectangle(-100, 100, -100, 0, 1.0, s.alpha, s.f)
    ref = fd.interpolate(xs, ys, U, s.points.numpy())
    solver = s.make_solver()
    r = orc.Problem.from_scenario(s, sigma_bar=solver.sigma_bar, majorant=solver.majorant).solve(
        s.points, 6000, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=2, compat="physical")
    z = np.abs(r["mean"] - ref) / (r["stderr"] + 0.01 * np.abs(ref).max())
    assert np.all(z <= 3.5), z
    assert 60 < r["steps"] / 6000 / len(s.points) < 200                 # ~110 steps per walk with the local majorant
