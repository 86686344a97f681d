"""Symbolic tracing of plain callables into exact device fields (dcrmontecarlo_b200/fieldtrace.py)."""
import numpy as np
import pytest
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


def test_untraceable_callables_fall_back_to_tabulation():
    cases = [
        lambda p: float(p[0] ** 2 + p[1] ** 2),                             # float() cast (tests/testWostWithSource.py:48)
        lambda p: 0.0 if p[0] < -2.0 else -4.0,                            # branch on a coordinate (:51-56)
        lambda p: 1.0 / (1.0 + p[0] ** 2),                                 # rational
        lambda p: torch.exp(-p[0] ** 2),                                   # anisotropic Gaussian
        lambda p: torch.sin(p[0] * p[1]),                                  # non-linear trig argument
        lambda p: torch.sqrt(p[0] ** 2 + 1.0),                             # function outside the algebra
        lambda p: (p - torch.tensor([0.1, 0.2])).norm(),                   # tensor methods on the point
    ]
    for fn in cases:
        assert trace_callable(fn, B) is None
        g = as_field(fn, bounds=B, n=17)
        assert isinstance(g, GridField)
    assert isinstance(as_field(lambda p: p[0] ** 2 - p[1] ** 2, bounds=B), TermField)
    assert isinstance(as_field(lambda p: p[0] ** 2 - p[1] ** 2, bounds=B, n=9, trace=False), GridField)


def test_wrong_trace_is_rejected_by_the_numerical_check():
    state = {"n": 0}

    def impure(p):                                                         # first call (the trace) differs from later calls
        state["n"] += 1
        return p[0] * (1.0 if state["n"] == 1 else 2.0)

    assert trace_callable(impure, B) is None
