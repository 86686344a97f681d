"""CPU-side tests: the C-ABI library loads and exports what include/wost.h declares, host logic (fields, scenarios,
sigma', sigma_bar, helper functions) matches the reference's values, and compute calls fail loudly without a GPU."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from dcrmontecarlo_b200 import _native as nat
from dcrmontecarlo_b200 import scenarios as sc
from dcrmontecarlo_b200 import utils as U
from dcrmontecarlo_b200.fields import GridField, TermField, as_field
from dcrmontecarlo_b200.geometry.Polylines import PolyLines
from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple
from dcrmontecarlo_b200.solvers import utils as SU
from dcrmontecarlo_b200.solvers.WoStSolver import WostSolver_2D

ROOT = Path(__file__).resolve().parents[1]
HAS_GPU = torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "wost.h").read_text()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(wost_\w+)\s*\(", header, flags=re.M))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    L = nat.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.wost_version() == 201
    assert L.wost_device_count() >= 0


def test_struct_layouts_match_header_sizes():
    import ctypes as C

    assert C.sizeof(nat._Term) == 64
    assert C.sizeof(nat.FieldDesc) == 80                                    # 60 bytes of scalars, padded to 64, two pointers
    assert C.sizeof(nat.Fields) == 40
    assert C.sizeof(nat.SolveParams) == 120


@pytest.mark.skipif(HAS_GPU, reason="checks the behaviour WITHOUT a GPU")
def test_compute_calls_fail_loudly_without_gpu():
    """No CPU fallback anywhere in the product path."""
    s = sc.cfg1a()
    solver = s.make_solver()
    with pytest.raises(nat.WostError, match="no CUDA device"):
        solver.solve(s.points, nWalks=4)
    with pytest.raises(nat.WostError, match="no CUDA device"):
        PolyLinesSimple(s.dirichlet).distance(torch.tensor([0.1, 0.2]))
    with pytest.raises(nat.WostError):
        nat.DeviceField(s.g, 0)
    import ctypes as C

    h = C.c_void_p(0)
    pts = np.zeros((3, 2), np.float32); pts[1, 0] = 1; pts[2, 1] = 1
    assert nat.lib().wost_scene_create(nat.ptr(pts), 3, None, 0, 0, C.byref(h)) == -2     # WOST_ERR_CUDA
    assert b"no CUDA device" in nat.lib().wost_last_error()
    assert nat.lib().wost_scene_create(None, 0, None, 0, 0, C.byref(h)) == -1             # WOST_ERR_INVALID comes first


# ---- helper functions of the reference's utils.py (its embedded tests, utils.py:133-233) ----------------
def test_torch_gradient_laplacian_gridsample():
    assert torch.allclose(U.torchGradient(lambda x: x ** 2, torch.tensor([3.0], requires_grad=True)), torch.tensor([6.0]))
    q2 = lambda p: p[0] ** 2 + p[1] ** 2                                                  # noqa: E731
    assert torch.allclose(U.torchGradient(q2, torch.tensor([2.0, 3.0], requires_grad=True)), torch.tensor([4.0, 6.0]))
    assert torch.allclose(U.torchGradient(lambda x: 3 * x + 2, torch.tensor([5.0], requires_grad=True)), torch.tensor([3.0]))
    assert torch.allclose(U.torchLaplacian(q2, torch.tensor([1.0, 2.0], requires_grad=True)), torch.tensor(4.0), atol=1e-6)
    q3 = lambda p: p[0] ** 2 + p[1] ** 2 + p[2] ** 2                                      # noqa: E731
    assert torch.allclose(U.torchLaplacian(q3, torch.tensor([1.0, 2.0, 3.0], requires_grad=True)), torch.tensor(6.0), atol=1e-6)
    q4 = lambda p: p[0] ** 4 + p[1] ** 4                                                  # noqa: E731
    assert torch.allclose(U.torchLaplacian(q4, torch.tensor([2.0, 1.0], requires_grad=True)), torch.tensor(60.0), atol=1e-5)
    # linear function: the second differentiation fails and the 1e-8 regulariser is what remains (utils.py:54-61)
    assert U.torchLaplacian(lambda p: 2 * p[0] + p[1], torch.tensor([1.0, 2.0])).item() == pytest.approx(1e-8)
    lo, hi, plo, phi = U.gridSampleMinMax(q2, [[-2.0, 2.0], [-2.0, 2.0]], 50)
    assert abs(lo) < 0.1 and abs(plo[0]) < 0.1 and abs(plo[1]) < 0.1 and hi == pytest.approx(8.0)
    lo, hi, plo, phi = U.gridSampleMinMax(lambda p: -p[0] ** 2 + 4, [[-3.0, 3.0]], 100)
    assert abs(hi - 4.0) < 0.1 and abs(phi[0]) < 0.1
    with pytest.raises(ValueError):
        U.gridSampleMinMax(lambda p: torch.tensor(float("nan")), [[0.0, 1.0]], 5)
    c = U.torch_smooth_circle(torch.tensor([0.0, 0.0]), torch.tensor([0.0, 0.0]), 1.0)
    assert c.item() == pytest.approx(1.0) and U.torch_smooth_circle(torch.tensor([3.0, 0.0]), torch.tensor([0.0, 0.0]), 1.0).item() < 1e-30


def test_polylines_interface():
    pts = sc.square(1.0)
    base = PolyLines(pts)
    assert len(base) == 5 and torch.equal(base[1], pts[1])
    for call in (lambda: base.distance(pts[0]), lambda: base.isSilhouette(pts[0]), lambda: base.silhouetteDistance(pts[0]),
                 lambda: base.rayIntersection(pts[0], pts[1]), lambda: base.intersectPolylines(pts[0], pts[1], 1.0)):
        with pytest.raises(NotImplementedError):
            call()
    line = PolyLinesSimple.funcToPolyline(lambda x: 0.5 * x, 3.0, 10.0, 0.5)              # x_min ignored (Q14)
    assert line.points[0, 0] == 0.0 and len(line) == 20 and isinstance(line, PolyLines)
    poly = PolyLinesSimple(pts)
    a, b = torch.tensor([1.0, 2.0]), torch.tensor([[3.0, 4.0], [0.0, 1.0]])
    assert torch.equal(poly.crossProduct2D(a, b), torch.tensor([1.0 * 4 - 2 * 3, 1.0]))


# ---- fields -------------------------------------------------------------------------------------------
def test_fields_evaluate_like_the_reference_callables():
    q = (torch.rand(500, 2, generator=torch.Generator().manual_seed(0)) * 2 - 1) * 1.6
    x, y = q[:, 0], q[:, 1]
    s4 = sc.cfg4()
    assert torch.allclose(s4.alpha(q), 0.5 + 1.5 * torch.exp(-2.0 * (x ** 2 + y ** 2)), atol=1e-6)
    assert torch.allclose(s4.sigma(q), 0.3 + 0.7 * (1 + torch.sin(2 * np.pi * x) * torch.cos(2 * np.pi * y)), atol=2e-6)
    assert torch.allclose(s4.g(q), torch.sin(np.pi * x) * torch.sin(np.pi * y), atol=1e-6)
    f_ref = torch.where(x ** 2 + y ** 2 > 1.5 ** 2, torch.zeros_like(x), torch.exp(-(x ** 2 + y ** 2)) * torch.sin(np.pi * x) * torch.cos(np.pi * y))
    assert torch.allclose(s4.f(q), f_ref, atol=1e-6)
    s1 = sc.cfg1b()
    u = (1 - x ** 2) * (1 - y ** 2)
    D = 2 + 0.5 * x + 0.5 * y
    f1 = -(D * (-2 * (2 - x ** 2 - y ** 2)) + (-x * (1 - y ** 2) - y * (1 - x ** 2))) + (2 + x * y) * u   # testWoStCorrectness.py:124-140
    assert torch.allclose(s1.f(q), f1, atol=2e-5) and torch.allclose(s1.g(q), u, atol=1e-6)
    s5 = sc.cfg5()
    q5 = q * 60
    circ = lambda c, r: torch.sigmoid(-100 * ((q5 - torch.tensor(c)).norm(dim=1) - r))    # noqa: E731
    assert torch.allclose(s5.alpha(q5), 100 - 90 * circ([-20.0, -30.0], 10.0) + 900 * circ([25.0, -40.0], 10.0), rtol=1e-5)
    nrm = 1 / (2 * np.pi * 0.25)
    f5 = nrm * torch.exp(-((q5[:, 0] + 10) ** 2 + q5[:, 1] ** 2) / 0.5) + nrm * torch.exp(-((q5[:, 0] - 10) ** 2 + q5[:, 1] ** 2) / 0.5)
    assert torch.allclose(s5.f(q5), f5, atol=1e-7)
    s3 = sc.cfg3()
    assert s3.f(torch.tensor([0.0, 0.0])).item() == -4.0 and s3.f(torch.tensor([2.5, 0.0])).item() == 0.0
    # single-point calls return 0-d tensors and support autograd like the reference's callables
    p = torch.tensor([0.3, -0.2], requires_grad=True)
    v = s4.alpha(p)
    assert v.dim() == 0
    (gr,) = torch.autograd.grad(v, p)
    assert torch.allclose(gr, -6.0 * torch.exp(-2 * (p ** 2).sum()) * p.detach(), atol=1e-6)


def test_field_algebra_and_tabulation():
    a = TermField.polynomial({(0, 0): 1.0, (1, 0): 2.0}) + TermField.polynomial({(0, 1): 3.0}) + 0.5
    p = torch.tensor([0.5, -1.0])
    assert a(p).item() == pytest.approx(1.0 + 1.0 - 3.0 + 0.5)
    assert (2.0 * a)(p).item() == pytest.approx(2 * a(p).item())
    with pytest.raises(ValueError):
        a.masked_box(0, 1, 0, 1) + a
    assert as_field(None) is None and as_field(3.0)(p).item() == 3.0 and as_field(a) is a
    with pytest.raises(ValueError):
        as_field(lambda q: q[0])
    g = GridField.from_callable(lambda q: q[0] * q[1], [[-1.0, 1.0], [-1.0, 1.0]], n=33, margin=0.0)
    assert g.values.shape == (33, 33) and g(torch.tensor([0.5, 0.5])).item() == pytest.approx(0.25, abs=2e-3)
    assert g(torch.tensor([5.0, 5.0])).item() == pytest.approx(1.0)                        # clamped to the table
    d = g.describe()
    assert d["kind"] == 1 and d["nx"] == 33 and d["grid"].dtype == np.float32


# ---- solver construction (host): sigma', sigma_bar, reference quirks ----------------------------------------
@pytest.mark.parametrize("key", ["cfg1b", "cfg4", "cfg5"])
def test_sigma_bar_and_sigma_prime_match_reference(golden, key):
    G = golden["sigma"]
    solver = sc.ALL[key]().make_solver()
    assert solver.use_delta_tracking
    assert solver.sigma_bar == pytest.approx(float(G[f"{key}_sigma_bar"]), rel=1e-6)
    got = np.array([float(solver.sigma_prime(torch.from_numpy(q))) for q in G[f"{key}_q"][:40]])
    ref = G[f"{key}_sigma_prime"][:40]
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5 * max(1.0, float(np.abs(ref).max())))
    assert solver.sp_mode == (nat.SP_RATIO if key == "cfg4" else nat.SP_FULL)


def test_solver_attributes_and_plain_callables():
    s = sc.cfg2()
    solver = s.make_solver()
    assert not solver.use_delta_tracking and solver.source is None
    (x0, x1), (y0, y1) = solver.domain_bounds
    assert (float(x0), float(x1), float(y0), float(y1)) == (-2.0, 2.0, -2.0, 2.0)
    assert WostSolver_2D(PolyLinesSimple(s.dirichlet)).boundaryDirichlet(torch.zeros(2)) == 0.0   # default g (:45-46)
    f = lambda p: 1.0                                                                     # noqa: E731
    solver.setSourceTerm(f); solver.setBoundaryConditions(f)
    assert solver.source is f and solver.boundaryDirichlet is f
    # sigma-only with a plain python alpha default and python-float callables construct fine
    so = WostSolver_2D(PolyLinesSimple(s.dirichlet), sigma=lambda p: 1.0 + 0.0 * p[0])
    assert so.use_delta_tracking and so.sigma_bar == 10.0 and so.sp_mode == nat.SP_RATIO     # constant sigma': range 0 -> fallback 10 (Q13)
    # alpha = 2 + x, sigma = 1 at (0.3, 0.2): closed form 0.3875236 (SURVEY §8c probe)
    sa = WostSolver_2D(PolyLinesSimple(s.dirichlet), alpha=lambda p: 2.0 + p[0], sigma=lambda p: 1.0 + 0.0 * p[0])
    assert float(sa.sigma_prime(torch.tensor([0.3, 0.2]))) == pytest.approx(0.3875236, abs=1e-6)
    assert sa.sp_mode == nat.SP_FULL                 # closed-form callables are traced into exact device fields
    su = WostSolver_2D(PolyLinesSimple(s.dirichlet), alpha=lambda p: 2.0 + torch.sqrt(p[0] ** 2 + 1.0), sigma=lambda p: 1.0 + 0.0 * p[0])
    assert su.sp_mode == nat.SP_FIELD                # differentiable but outside the term algebra: sigma' is tabulated
    with pytest.raises(ValueError):
        WostSolver_2D(PolyLinesSimple(s.dirichlet), sigma_prime_mode="bogus")


def test_screened_table_and_green_helpers(golden):
    S = golden["samplers"]
    for sb in (2.40625, 10.0):
        tab = SU.screened_radius_icdf(sb)
        assert tab.dtype == np.float32 and len(tab) == 1024 and np.all(np.diff(tab) >= 0)
        from scipy.stats import ks_2samp
        rng = np.random.default_rng(1)
        pos = rng.random(40000) * 1023; i = np.minimum(pos.astype(int), 1022)
        assert ks_2samp(tab[i] + (pos - i) * (tab[i + 1] - tab[i]), S[f"screened_cache_seed42_sb{sb}"]).pvalue > 1e-3
        assert np.allclose([SU.screenedGreensNorm2D(R, sb) for R in S["norm_R"]], S[f"norm_sb{sb}"], rtol=1e-12)
        got = [SU.screenedGreens2D(torch.zeros(2), torch.tensor([r, 0.0]), 1.0, sb) for r in S[f"greens_r_sb{sb}"]]
        assert np.allclose(got, S[f"greens_sb{sb}"], rtol=2e-6)
    assert SU.greensFunctionNorm2D(2.0) == 1.0 and SU.greensFunction2D(torch.zeros(2), torch.zeros(2), 1.0) == 0.0
    np.random.seed(0)
    g = SU.GreensDistribution2D(); sg = SU.ScreenedGreensDistribution2D(2.40625)
    draws = np.array([g.sample(None, 2.0) for _ in range(4000)])
    assert abs(draws.mean() / 2.0 - 0.25) < 0.02 and 0 < sg.sample(None, 0.5) <= 0.5
    assert g.pdf(0.5, None, 1.0) == pytest.approx(-np.log(0.5) * 4) and g.pdf(2.0, None, 1.0) == 0.0 and sg.pdf(0.3, None, 1.0) > 0


def test_scenarios_match_survey_sizes():
    sizes = {k: len(f().points) for k, f in sc.ALL.items()}
    assert sizes == {"cfg1a": 16, "cfg1b": 16, "cfg2": 404, "cfg3": 404, "cfg4": 648, "cfg5": 9}
    assert len(sc.cfg2().neumann) == 33 and len(sc.cfg5().neumann) == 2
    assert len(sc.cfg5(50).points) == 50 and len(sc.cfg5(175).points) == 175
    big = sc.cfg2_throughput(4096, 8)
    assert big.points.shape == (4096, 2) and float(torch.norm(big.points, dim=1).min()) > 0.6
    assert sc.scale_scene(64, 128, 4).dirichlet.shape == (65, 2)


def test_mis_helpers_importable_and_consistent():
    np.random.seed(3)
    mis = SU.MultipleImportanceSampler2D([SU.GreensDistribution2D(), SU.UniformDistribution2D()], [3.0, 1.0])
    assert np.allclose(mis.weights, [0.75, 0.25])
    r, i, w = mis.sample(None, 2.0)
    assert 0 < r <= 2.0 and i in (0, 1) and 0.0 <= w <= 1.0
    tot = sum(mis._compute_mis_weight(0.7, None, 2.0, k) for k in range(2))
    assert tot == pytest.approx(1.0)
    assert 0 < SU.sampleGreensFunction2D(None, 1.0) <= 1.0 and 0 < SU.sampleScreenedGreensFunction2D(None, 1.0, 2.0) <= 1.0
    assert SU.UniformDistribution2D().pdf(0.5, None, 2.0) == 0.5 and SU.UniformDistribution2D().pdf(3.0, None, 2.0) == 0.0


def test_survey_helpers():
    from dcrmontecarlo_b200.survey import DipoleSource, geometric_factor_2d

    src = DipoleSource((-10.0, 0.0), (10.0, 0.0), current=2.0, width=0.5)
    f = src.field()
    nrm = 2.0 / (2 * np.pi * 0.25)
    assert f(torch.tensor([-10.0, 0.0])).item() == pytest.approx(nrm, rel=1e-6)
    assert f(torch.tensor([10.0, 0.0])).item() == pytest.approx(-nrm, rel=1e-6)
    assert src.field(sink_sign=+1.0)(torch.tensor([10.0, 0.0])).item() == pytest.approx(nrm, rel=1e-6)   # reference quirk
    # homogeneous half-plane: V = -(rho I/pi) ln r per electrode  =>  K dV / I recovers rho
    a, b, m, n = (-3.0, 0.0), (3.0, 0.0), (-1.0, 0.0), (1.5, 0.0)
    rho, I = 40.0, 1.0
    V = lambda p: -(rho * I / np.pi) * (np.log(np.hypot(p[0] - a[0], p[1])) - np.log(np.hypot(p[0] - b[0], p[1])))   # noqa: E731
    assert geometric_factor_2d(a, b, m, n) * (V(m) - V(n)) / I == pytest.approx(rho)


@pytest.mark.parametrize("key", ["cfg1a", "cfg1b", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_specialised_kernel_compiles_without_a_device(key, tmp_path):
    """wost_jit_offline: the per-solver kernel (NVRTC, wost_jit.inc) compiles for every reference scenario on a CPU-only
    box; the generated source holds the fields as constants and the cubin is sm_100a code without local-memory spills."""
    import subprocess

    s = sc.ALL[key]()
    prefix = tmp_path / f"jit_{key}"
    note = nat.jit_offline(dict(g=s.g, f=s.f, alpha=s.alpha, sigma=s.sigma), neu=s.neumann is not None, src=s.f is not None,
                           delta=s.delta, sp_mode=s.sp_mode if s.delta else 0, prefix=prefix,
                           n_dseg=len(s.dirichlet) - 1, n_nseg=(len(s.neumann) - 1) if s.neumann is not None else 0)
    assert note.startswith("compiled in")
    src = prefix.with_suffix(".cu").read_text()
    assert "struct JitFP" in src and "wost_walk_jit" in src and "term_value_t" in src
    res = subprocess.run(["cuobjdump", "-res-usage", str(prefix.with_suffix(".cubin"))], capture_output=True, text=True).stdout
    assert "wost_walk_jit" in res
    m = re.search(r"REG:(\d+) STACK:(\d+)", res)
    assert m and int(m.group(1)) <= 64 and int(m.group(2)) == 0, res


def test_cached_fields_follow_closure_state_and_setters():
    """a traced / tabulated callable is cached per solver; if the state it closes over changes between solves the
    cached field must not be used silently (the reference calls its callables live, solvers/WoStSolver.py:253-256)."""
    from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple
    from dcrmontecarlo_b200.solvers.WoStSolver import WostSolver_2D

    state = {"k": 1.0}
    g = lambda p: state["k"] * p[0]                                              # noqa: E731
    solver = WostSolver_2D(PolyLinesSimple(sc.square(1.0)), g)
    q = torch.tensor([[0.5, 0.25]])
    assert float(solver._host_field(g)(q)[0]) == pytest.approx(0.5)
    assert solver._host_field(g) is solver._host_field(g)                        # cached while the callable is unchanged
    state["k"] = 5.0
    assert float(solver._host_field(g)(q)[0]) == pytest.approx(2.5)              # re-traced, not stale
    # setters drop the cache entries of the callable they replace
    n = len(solver._cache)
    solver.setBoundaryConditions(lambda p: 2.0 * p[1])
    assert len(solver._cache) < n
    solver.invalidate()
    assert all(k[0] == "scene" for k in solver._cache)


def test_sigma_prime_table_is_cached_across_solves():
    """in SP_FIELD mode the sigma' table was keyed on id(bound method) -- a new object on every access -- so
    every solve re-tabulated it point by point.  The callable is bound once now."""
    from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple
    from dcrmontecarlo_b200.solvers.WoStSolver import WostSolver_2D

    s = sc.cfg1b()
    alpha = lambda p: 2.0 + torch.sqrt(p[0] ** 2 + 1.0) * 0.5 + 0.25 * p[1]      # noqa: E731  (outside the term algebra)
    solver = WostSolver_2D(PolyLinesSimple(s.dirichlet), s.g, None, source=s.f, alpha=alpha, sigma=s.sigma,
                           field_resolution=33, sigma_prime_resolution=9)
    assert solver.sp_mode == nat.SP_FIELD
    assert solver._sp_plain is solver._sp_plain
    a = solver._host_field(solver._sp_plain, solver.sigma_prime_resolution)
    n = len(solver._cache)
    b = solver._host_field(solver._sp_plain, solver.sigma_prime_resolution)
    assert a is b and len(solver._cache) == n
