"""Symbolic tracing of plain callables into exact device fields (dcrmontecarlo_b200/fieldtrace.py)."""
import numpy as np
import torch

from dcrmontecarlo_b200.fields import GridField, TermField, as_field
from dcrmontecarlo_b200.fieldtrace import trace_callable

B = [[-1.5, 1.5], [-1.5, 1.5]]
Q = (torch.rand(400, 2, generator=torch.Generator().manual_seed(0)) * 2 - 1) * 1.5


def check(fn, expect_terms=None):
    f = trace_callable(fn, B)
    assert isinstance(f, TermField)
    ref = torch.stack([torch.as_tensor(fn(p), dtype=torch.float32) for p in Q])
    assert torch.allclose(f(Q), ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max() + 1e-9))
    if expect_terms is not None:
        assert len(f.terms) == expect_terms
    return f


def test_reference_correctness_script_callables_trace_exactly():
    """Every callable of the reference's tests/testWoStCorrectness.py:81-142, written exactly as there."""
    def diffusion_coefficient(point):
        return 2.0 + 0.5 * point[0] + 0.5 * point[1]

    def absorption_coefficient(point):
        return point[0] * point[1] + 2

    def boundary_condition(point):
        x, y = point[0], point[1]
        return (1 - x**2) * (1 - y**2)

    def source_term(point):
        x, y = point[0], point[1]
        u = (1 - x**2) * (1 - y**2)
        laplacian_u = -2 * (2 - x**2 - y**2)
        D = 2 + 0.5*x + 0.5*y
        gradD_dot_gradu = -x*(1 - y**2) - y*(1 - x**2)
        div_D_grad_u = D * laplacian_u + gradD_dot_gradu
        alpha = 2 + x * y
        return -div_D_grad_u + alpha * u

    assert check(diffusion_coefficient, 2).c0 == 2.0
    check(absorption_coefficient, 1)
    check(boundary_condition, 3)
    f = check(source_term)
    from dcrmontecarlo_b200 import scenarios as sc
    assert torch.allclose(f(Q), sc.cfg1b().f(Q), atol=2e-5)               # same polynomial as the hand-expanded scenario


def test_trig_exp_and_torch_functions():
    check(lambda p: torch.sin(torch.pi * p[0]) * torch.sin(torch.pi * p[1]), 1)
    check(lambda p: 0.3 + 0.7 * (1 + torch.sin(2 * np.pi * p[0]) * torch.cos(2 * np.pi * p[1])), 1)
    check(lambda p: 0.5 + 1.5 * torch.exp(-2.0 * (p[0] ** 2 + p[1] ** 2)), 1)
    check(lambda p: torch.exp(-((p[0] + 0.3) ** 2 + (p[1] - 0.2) ** 2) / (2 * 0.5 ** 2)) / (2 * torch.pi * 0.25), 1)   # off-centre Gaussian
    check(lambda p: torch.exp(-(p[0] ** 2 + p[1] ** 2)) * torch.sin(np.pi * p[0]) * torch.cos(np.pi * p[1]) * p[0] ** 2, 1)
    check(lambda p: (p[0] - p[1]) ** 3 / 4.0 - torch.cos(p[0] + 2 * p[1] + 0.5), 5)
    f = check(lambda p: 3.5 + 0 * p[0], 0)
    assert f.c0 == 3.5


def test_reference_scripts_callables_trace_exactly():
    """The callables of the reference's other three scripts, written as there: float()/torch.tensor() wrappers, `if`
    guards for outside-the-domain, torch_smooth_circle anomalies."""
    from dcrmontecarlo_b200 import scenarios as sc
    from dcrmontecarlo_b200.fields import MASK_BOX, MASK_DISC
    from dcrmontecarlo_b200.utils import torch_smooth_circle

    # tests/testWostWithSource.py:42-58
    def dirichlet_bc(point):
        x, y = point[0], point[1]
        return float(x**2 + y**2)

    def source_term(point):
        if point[0] < -2.0 or point[0] > 2.0 or point[1] < -2.0 or point[1] > 2.0:
            return 0.0
        return -4.0

    check(dirichlet_bc, 2)
    f = trace_callable(source_term, [[-2.0, 2.0], [-2.0, 2.0]])
    assert f.mask_kind == MASK_BOX and f.mask == (-2.0, 2.0, -2.0, 2.0) and f.c0 == -4.0 and f.outside == 0.0
    assert f(torch.tensor([1.9, -1.9])).item() == -4.0 and f(torch.tensor([2.1, 0.0])).item() == 0.0

    # tests/testWostVariableCoefficients.py:42-86
    def diffusion_coefficient(point):
        x, y = point[0], point[1]
        r_squared = x**2 + y**2
        return torch.tensor(0.5 + 1.5 * torch.exp(-2.0 * r_squared))

    def absorption_coefficient(point):
        x, y = point[0], point[1]
        return torch.tensor(0.3 + 0.7 * (1 + torch.sin(2*np.pi*x) * torch.cos(2*np.pi*y)))

    def dirichlet_bc4(point):
        x, y = point[0], point[1]
        return float(torch.sin(np.pi * x) * torch.sin(np.pi * y))

    def source_term4(point):
        x, y = point[0], point[1]
        r_squared = x**2 + y**2
        if r_squared > 1.5**2:
            return 0.0
        return float(torch.exp(-r_squared) * torch.sin(np.pi * x) * torch.cos(np.pi * y))

    s4 = sc.cfg4()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert torch.allclose(check(diffusion_coefficient, 1)(Q), s4.alpha(Q), atol=1e-6)
        assert torch.allclose(check(absorption_coefficient, 1)(Q), s4.sigma(Q), atol=2e-6)
    check(dirichlet_bc4, 1)
    f4 = trace_callable(source_term4, B)
    assert f4.mask_kind == MASK_DISC and f4.mask[:3] == (0.0, 0.0, 2.25)
    q4 = Q * 1.2
    assert torch.allclose(f4(q4), s4.f(q4), atol=1e-6)

    # tests/testGeophysicalScenario.py:11-55
    def dcr_current_source(point):
        x, y = point[0], point[1]
        current_amplitude = 1.0
        sigma = 0.5
        pos_dist2 = (x + 10.0)**2 + y**2
        neg_dist2 = (x - 10.0)**2 + y**2
        norm = current_amplitude / (2 * torch.pi * sigma**2)
        positive_source = norm * torch.exp(-pos_dist2 / (2 * sigma**2))
        negative_sink = -norm * torch.exp(-neg_dist2 / (2 * sigma**2))
        return float(positive_source - negative_sink)

    def conductivity_field(point):
        background_conductivity = 1e2
        anomaly_center1 = torch.tensor([-20, -30])
        anomaly_center2 = torch.tensor([25, -40])
        anomaly1 = (1e1 - background_conductivity) * torch_smooth_circle(point, anomaly_center1, 10)
        anomaly2 = (1e3 - background_conductivity) * torch_smooth_circle(point, anomaly_center2, 10)
        return background_conductivity + anomaly1 + anomaly2

    s5 = sc.cfg5()
    B5 = [[-100.0, 100.0], [-100.0, 100.0]]
    q5 = (torch.rand(3000, 2, generator=torch.Generator().manual_seed(4)) * 2 - 1) * 60
    q5[:200] = torch.tensor([-20.0, -30.0]) + 10.0 * torch.nn.functional.normalize(torch.randn(200, 2, generator=torch.Generator().manual_seed(5)), dim=1) \
        * (1 + 0.004 * torch.randn(200, 1, generator=torch.Generator().manual_seed(6)))          # on the 1 cm rim of an anomaly
    fs, fa = trace_callable(dcr_current_source, B5), trace_callable(conductivity_field, B5)
    assert len(fs.terms) == 2 and len(fa.terms) == 2 and fa.c0 == 100.0
    assert torch.allclose(fs(q5), s5.f(q5), atol=1e-7) and torch.allclose(fa(q5), s5.alpha(q5), rtol=1e-5)   # same terms as the scenario
    ref = torch.stack([conductivity_field(p) for p in q5[:300]])
    # the same closed form on the 1 cm rim (where a table cannot follow); fp32 rounding of |x-c| times k = 100 leaves ~1e-4
    assert torch.allclose(fa(q5[:300]), ref, rtol=2e-3)


def test_untraceable_callables_fall_back_to_tabulation():
    cases = [
        lambda p: 1.0 / (1.0 + p[0] ** 2),                                 # rational
        lambda p: torch.exp(-p[0] ** 2),                                   # anisotropic Gaussian
        lambda p: torch.sin(p[0] * p[1]),                                  # non-linear trig argument
        lambda p: torch.sqrt(p[0] ** 2 + 1.0),                             # sqrt of something that is not a squared distance
        lambda p: torch.tanh(p[0]),                                        # function outside the algebra
        lambda p: (p - torch.tensor([0.1, 0.2])).norm(),                   # a bare distance (only sigmoid of it is a device term)
        lambda p: p[0] if p[0] > 0 else p[1],                              # branch that is not a constant-outside guard
        lambda p: float(abs(p[0])),
    ]
    for fn in cases:
        assert trace_callable(fn, B) is None
        g = as_field(fn, bounds=B, n=17)
        assert isinstance(g, GridField)
    assert isinstance(as_field(lambda p: p[0] ** 2 - p[1] ** 2, bounds=B), TermField)
    assert isinstance(as_field(lambda p: p[0] ** 2 - p[1] ** 2, bounds=B, n=9, trace=False), GridField)
    assert float is __builtins__["float"] if isinstance(__builtins__, dict) else True      # tracing leaves `float` alone
    assert "float" not in globals()


def test_wrong_trace_is_rejected_by_the_numerical_check():
    state = {"n": 0}

    def impure(p):                                                         # first call (the trace) differs from later calls
        state["n"] += 1
        return p[0] * (1.0 if state["n"] == 1 else 2.0)

    assert trace_callable(impure, B) is None


def test_tabulation_error_control():
    import warnings

    smooth = lambda p: 1.0 / (1.0 + p[0] ** 2 + p[1] ** 2)                      # noqa: E731  (rational: not traceable)
    coarse = GridField.from_callable(smooth, B, n=9)
    fine = GridField.from_callable(smooth, B, n=9, tol=1e-4)
    assert coarse.interp_error is None and coarse.nx == 9
    assert fine.nx > 9 and fine.interp_error <= 1e-4
    exact = torch.stack([smooth(p) for p in Q[:200]])
    assert (fine(Q[:200]) - exact).abs().max() < 2e-4 < (coarse(Q[:200]) - exact).abs().max()
    steep = lambda p: torch.tanh(400.0 * p[0])                                  # noqa: E731  (cannot be resolved by n_max)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        g = GridField.from_callable(steep, B, n=17, tol=1e-3, n_max=129)
    assert g.nx == 129 and g.interp_error > 1e-3 and any("interpolation error" in str(x.message) for x in w)
    branchy = lambda p: 1.0 if float(p[0]) > 0 else -1.0                        # noqa: E731  (loop path: no refinement)
    assert GridField.from_callable(branchy, B, n=9, tol=1e-6).nx == 9
