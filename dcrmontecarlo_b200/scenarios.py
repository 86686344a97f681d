"""The reference's five scenarios (plus scale-up scenes) expressed with device fields.

Each scenario mirrors one of the reference's driver scripts (cited per function) so that tests,
``bench.py`` and the golden-fixture generator all build exactly the same inputs.  Numbering follows
SURVEY.md §8(d): 1a Laplace/Dirichlet, 1b ``testWoStCorrectness`` as shipped, 2 mixed
Dirichlet/Neumann, 3 ``testWostWithSource``, 4 ``testWostVariableCoefficients``,
5 ``testGeophysicalScenario``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from .fields import Field, TermField, make_term

SP_FULL, SP_RATIO, SP_FIELD = 0, 1, 2


@dataclass
class Scenario:
    name: str
    dirichlet: torch.Tensor                 # (N,2) float32 vertices
    neumann: Optional[torch.Tensor]
    points: torch.Tensor                    # (P,2) float32 evaluation points
    g: Optional[Field] = None
    f: Optional[Field] = None
    alpha: Optional[Field] = None
    sigma: Optional[Field] = None
    n_walks: int = 100
    max_steps: int = 1000
    eps: float = 1e-4
    sp_mode: int = SP_FULL                  # how sigma' is formed when delta tracking is on
    sigma_bar: Optional[float] = None       # value the reference ctor computes, for checks
    analytic: Optional[Callable] = None     # exact solution on (P,2) points, if one exists
    notes: str = ""

    @property
    def delta(self) -> bool:
        return self.alpha is not None or self.sigma is not None

    compat: str = "reference"               # "physical" for the textbook-WoSt validation scenes below

    def make_solver(self, **kw):
        """The product solver (``solvers.WoStSolver.WostSolver_2D``) for this scenario."""
        from .geometry.PolylinesSimple import PolyLinesSimple
        from .solvers.WoStSolver import WostSolver_2D

        mode = "ratio" if (self.delta and self.sp_mode == SP_RATIO) else "auto"
        return WostSolver_2D(PolyLinesSimple(self.dirichlet), self.g,
                             PolyLinesSimple(self.neumann) if self.neumann is not None else None,
                             self.f, self.sigma, self.alpha, sigma_prime_mode=mode, compat=self.compat, **kw)


def square(half: float) -> torch.Tensor:
    """Closed CCW square (reference tests/testWoStCorrectness.py:10-20)."""
    return torch.tensor([[-half, -half], [half, -half], [half, half], [-half, half], [-half, -half]], dtype=torch.float32)


def circle(radius: float, n: int = 32) -> torch.Tensor:
    """Closed n-gon, theta = linspace(0, 2pi, n+1) (reference tests/testWostWithSource.py:28-35)."""
    theta = torch.linspace(0, 2 * np.pi, n + 1)
    return torch.stack([radius * torch.cos(theta), radius * torch.sin(theta)], dim=1)


def grid_points(lim: float, n: int, hole: float = 0.0) -> torch.Tensor:
    x = torch.linspace(-lim, lim, n)
    X, Y = torch.meshgrid(x, x, indexing="ij")
    pts = torch.stack([X.flatten(), Y.flatten()], dim=1)
    if hole > 0.0:
        pts = pts[torch.norm(pts, dim=1) > hole]
    return pts.contiguous()


# --- tiny polynomial algebra over {(i,j): c} dicts (to expand manufactured sources) ---------------
def _padd(*ps):
    out: dict = {}
    for p in ps:
        for k, c in p.items():
            out[k] = out.get(k, 0.0) + c
    return {k: c for k, c in out.items() if c != 0.0}


def _pmul(a, b):
    out: dict = {}
    for (i, j), c in a.items():
        for (k, l), d in b.items():
            out[(i + k, j + l)] = out.get((i + k, j + l), 0.0) + c * d
    return {k: c for k, c in out.items() if c != 0.0}


def _pscale(a, s):
    return {k: c * s for k, c in a.items()}


def cfg1a() -> Scenario:
    """Laplace, pure Dirichlet, g = x^2 - y^2 (harmonic) on the square of tests/testWoStCorrectness.py:10-20."""
    return Scenario(
        name="cfg1a_laplace_dirichlet", dirichlet=square(1.0), neumann=None, points=grid_points(0.7, 4),
        g=TermField.polynomial({(2, 0): 1.0, (0, 2): -1.0}), n_walks=150, max_steps=800, eps=1e-4,
        analytic=lambda p: p[:, 0] ** 2 - p[:, 1] ** 2,
    )


def cfg1b() -> Scenario:
    """tests/testWoStCorrectness.py:81-196 as shipped: u=(1-x^2)(1-y^2), D=2+.5x+.5y, sigma=2+xy, delta tracking."""
    X, Y, ONE = {(1, 0): 1.0}, {(0, 1): 1.0}, {(0, 0): 1.0}
    x2, y2 = _pmul(X, X), _pmul(Y, Y)
    u = _pmul(_padd(ONE, _pscale(x2, -1)), _padd(ONE, _pscale(y2, -1)))
    D = {(0, 0): 2.0, (1, 0): 0.5, (0, 1): 0.5}
    absorb = {(0, 0): 2.0, (1, 1): 1.0}
    # f = 2 D (2 - x^2 - y^2) + x(1-y^2) + y(1-x^2) + (2+xy) u      (:124-140)
    f = _padd(
        _pscale(_pmul(D, _padd({(0, 0): 2.0}, _pscale(x2, -1), _pscale(y2, -1))), 2.0),
        _pmul(X, _padd(ONE, _pscale(y2, -1))),
        _pmul(Y, _padd(ONE, _pscale(x2, -1))),
        _pmul(absorb, u),
    )
    return Scenario(
        name="cfg1b_correctness_as_shipped", dirichlet=square(1.0), neumann=None, points=grid_points(0.7, 4),
        g=TermField.polynomial(u), f=TermField.polynomial(f), alpha=TermField.polynomial(D),
        sigma=TermField.polynomial(absorb), n_walks=150, max_steps=800, eps=1e-4, sp_mode=SP_FULL,
        sigma_bar=2.40625, analytic=lambda p: (1 - p[:, 0] ** 2) * (1 - p[:, 1] ** 2),
    )


def cfg2() -> Scenario:
    """Mixed Dirichlet square +-2 / Neumann 32-gon r=0.5 (geometry of tests/testWostWithSource.py:19-40), Laplace, g = x."""
    return Scenario(
        name="cfg2_mixed_dirichlet_neumann", dirichlet=square(2.0), neumann=circle(0.5, 32),
        points=grid_points(1.8, 21, hole=0.6), g=TermField.polynomial({(1, 0): 1.0}),
        n_walks=150, max_steps=500, eps=1e-4,
    )


def cfg3() -> Scenario:
    """tests/testWostWithSource.py:42-110: Poisson, f = -4 inside the box, g = x^2 + y^2, Dirichlet only."""
    return Scenario(
        name="cfg3_poisson_source", dirichlet=square(2.0), neumann=None, points=grid_points(1.8, 21, hole=0.6),
        g=TermField.polynomial({(2, 0): 1.0, (0, 2): 1.0}),
        f=TermField.constant(-4.0).masked_box(-2.0, 2.0, -2.0, 2.0, outside=0.0),
        n_walks=150, max_steps=500, eps=1e-4, analytic=lambda p: p[:, 0] ** 2 + p[:, 1] ** 2,
    )


def cfg4() -> Scenario:
    """tests/testWostVariableCoefficients.py:12-105,218-233: mixed boundary, variable alpha/sigma, source."""
    two_pi, pi = float(np.float32(2 * np.pi)), float(np.float32(np.pi))
    alpha = TermField(0.5, [make_term(A=1.5, q=2.0)])                                     # :42-49
    sigma = TermField(1.0, [make_term(A=0.7, trig1=("sin", two_pi, 0.0, 0.0), trig2=("cos", 0.0, two_pi, 0.0))])  # :51-57
    g = TermField(0.0, [make_term(A=1.0, trig1=("sin", pi, 0.0, 0.0), trig2=("sin", 0.0, pi, 0.0))])               # :67-72
    f = TermField(0.0, [make_term(A=1.0, q=1.0, trig1=("sin", pi, 0.0, 0.0), trig2=("cos", 0.0, pi, 0.0))]) \
        .masked_disc((0.0, 0.0), 1.5, outside=0.0)                                        # :74-84
    return Scenario(
        name="cfg4_variable_coefficients", dirichlet=square(1.5), neumann=circle(0.4, 32),
        points=grid_points(1.3, 27, hole=0.5), g=g, f=f, alpha=alpha, sigma=sigma,
        n_walks=25, max_steps=1000, eps=1e-4, sp_mode=SP_RATIO, sigma_bar=3.2175,
        notes="sigma' = sigma/alpha: the reference's autograd fails on these callables (SURVEY Q12)",
    )


def cfg5(n_electrodes: int = 9, n_walks: int = 100) -> Scenario:
    """tests/testGeophysicalScenario.py:11-151: DCR survey, Neumann surface, conductivity anomalies. eps=0.9."""
    norm = 1.0 / (2 * math.pi * 0.5 ** 2)                                                 # :27
    f = TermField.gaussian_sum([(norm, (-10.0, 0.0), 2.0), (norm, (10.0, 0.0), 2.0)])     # :22-33 (both positive)
    alpha = TermField.smooth_circle_sum(100.0, [(10.0 - 100.0, (-20.0, -30.0), 10.0), (1000.0 - 100.0, (25.0, -40.0), 10.0)])  # :35-55
    if n_electrodes == 9:
        xs = torch.arange(-40.0, 40.0 + 10.0, 10.0)                                       # :58-75,109-113
    else:
        xs = torch.linspace(-40.0, 40.0, n_electrodes)
    pts = torch.stack([xs, torch.zeros_like(xs)], dim=1)
    neumann = torch.tensor([[-100.0, 100.0], [100.0, 100.0]], dtype=torch.float32)        # :99-102
    return Scenario(
        name=f"cfg5_dcr_survey_{n_electrodes}e", dirichlet=square(100.0), neumann=neumann, points=pts,
        g=TermField.constant(0.0), f=f, alpha=alpha, sigma=None, n_walks=n_walks, max_steps=500, eps=0.9,
        sp_mode=SP_FULL, sigma_bar=10.0, notes="sigma_bar falls back to 10.0 (SURVEY Q13)",
    )


def ngon(radius: float, n: int) -> torch.Tensor:
    theta = torch.linspace(0, 2 * np.pi, n + 1, dtype=torch.float64)
    p = torch.stack([radius * torch.cos(theta), radius * torch.sin(theta)], dim=1).to(torch.float32)
    p[-1] = p[0]
    return p


def annulus_points(n: int, r_in: float, r_out: float, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    r = torch.sqrt(torch.rand(n, generator=g) * (r_out ** 2 - r_in ** 2) + r_in ** 2)
    t = torch.rand(n, generator=g) * 2 * math.pi
    return torch.stack([r * torch.cos(t), r * torch.sin(t)], dim=1).to(torch.float32).contiguous()


def scale_scene(n_seg: int, n_points: int = 65536, n_walks: int = 1024) -> Scenario:
    """regular N-gon Dirichlet r=1 + inner N-gon Neumann r=0.4, Laplace, g = x."""
    return Scenario(
        name=f"scale_{n_seg}", dirichlet=ngon(1.0, n_seg), neumann=ngon(0.4, n_seg),
        points=annulus_points(n_points, 0.45, 0.95), g=TermField.polynomial({(1, 0): 1.0}),
        n_walks=n_walks, max_steps=1000, eps=1e-4,
    )


def uniform_points_outside_disc(n: int, half: float, hole: float, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(int(n * 1.3) + 64, 2, generator=g) * 2 - 1) * half
    pts = pts[torch.norm(pts, dim=1) > hole][:n]
    assert pts.shape[0] == n
    return pts.to(torch.float32).contiguous()


def cfg2_throughput(n_points: int = 65536, n_walks: int = 1024) -> Scenario:
    """cfg2's scene with a throughput-sized evaluation set (bench.py workload)."""
    s = cfg2()
    s.name = "cfg2_mixed_dirichlet_neumann_throughput"
    s.points = uniform_points_outside_disc(n_points, 1.9, 0.6)
    s.n_walks = n_walks
    return s


# ---- analytic mixed-boundary problems for compat="physical" (SURVEY §4 iv; not in the reference) -------------------
def _unit_square_neumann_top():
    dirichlet = torch.tensor([[0.0, 1.0], [0.0, 0.0], [1.0, 0.0], [1.0, 1.0]])       # left, bottom, right edges
    neumann = torch.tensor([[1.0, 1.0], [0.0, 1.0]])                                 # top edge, du/dy = 0
    pts = torch.tensor([[0.5, 0.5], [0.3, 0.8], [0.7, 0.95], [0.2, 0.3], [0.5, 0.99], [0.9, 0.6], [0.1, 0.9], [0.6, 0.2]])
    return dirichlet, neumann, pts


def phys_laplace_neumann_top() -> Scenario:
    """Harmonic u = sin(pi x) cosh(pi (y-1)) / cosh(pi) on the unit square: du/dy = 0 on the top edge."""
    from .fields import GridField

    d, n, pts = _unit_square_neumann_top()
    g = GridField.from_callable(lambda p: torch.sin(math.pi * p[0]) * torch.cosh(math.pi * (p[1] - 1.0)) / math.cosh(math.pi),
                                [[0.0, 1.0], [0.0, 1.0]], n=513)
    return Scenario(name="phys_laplace_neumann_top", dirichlet=d, neumann=n, points=pts, g=g, n_walks=20000, max_steps=1000,
                    eps=1e-4, compat="physical",
                    analytic=lambda p: torch.sin(math.pi * p[:, 0]) * torch.cosh(math.pi * (p[:, 1] - 1.0)) / math.cosh(math.pi))


def phys_poisson_neumann_top() -> Scenario:
    """u = cos(pi (1-y)):  -lap u = pi^2 cos(pi (1-y)),  du/dy = 0 on the top edge."""
    d, n, pts = _unit_square_neumann_top()
    pi = float(np.float32(math.pi))
    g = TermField(0.0, [make_term(A=1.0, trig1=("cos", 0.0, -pi, pi))])
    f = TermField(0.0, [make_term(A=math.pi ** 2, trig1=("cos", 0.0, -pi, pi))])
    return Scenario(name="phys_poisson_neumann_top", dirichlet=d, neumann=n, points=pts, g=g, f=f, n_walks=20000,
                    max_steps=1000, eps=1e-4, compat="physical", analytic=lambda p: torch.cos(math.pi * (1.0 - p[:, 1])))


def phys_cylinder(n_seg: int = 256) -> Scenario:
    """Potential flow around a cylinder: u = x (1 + R^2/r^2) is harmonic with du/dr = 0 on r = R = 0.5.
    Dirichlet square +-2, Neumann n_seg-gon (>= 192 segments exercises the BVH + silhouette cones)."""
    from .fields import GridField

    R = 0.5
    g = GridField.from_callable(lambda p: p[0] * (1.0 + R * R / (p[0] ** 2 + p[1] ** 2 + 1e-12)), [[-2.0, 2.0], [-2.0, 2.0]], n=1025)
    pts = torch.tensor([[0.8, 0.1], [0.0, 0.7], [-1.2, 0.9], [0.55, 0.0], [1.5, -1.0], [-0.6, -0.3], [0.4, 0.4], [1.9, 1.9]])
    return Scenario(name=f"phys_cylinder_{n_seg}", dirichlet=square(2.0), neumann=ngon(R, n_seg).flip(0).contiguous(), points=pts, g=g,
                    n_walks=20000, max_steps=2000, eps=1e-4, compat="physical",
                    analytic=lambda p: p[:, 0] * (1.0 + R * R / (p[:, 0] ** 2 + p[:, 1] ** 2)))


def phys_varcoef_dirichlet() -> Scenario:
    """The reference's own correctness problem (tests/testWoStCorrectness.py:81-196: u = (1-x^2)(1-y^2), alpha = 2+.5x+.5y,
    sigma = 2+xy) solved by delta tracking done by the book.  The reference's estimator plateaus at RMSE 0.028 on it
    (cfg 1b); this one converges to the analytic solution."""
    s = cfg1b()
    s.name, s.compat, s.sp_mode, s.sigma_bar = "phys_varcoef_dirichlet", "physical", SP_FULL, None
    return s


def phys_varcoef_neumann_top() -> Scenario:
    """Variable coefficients with a reflecting wall: unit square, zero-Neumann top edge,
    alpha = (1 + 0.3 x)(1 + 0.5 (1-y)^2)  (d alpha/dy = 0 on the wall), sigma = 0.5 + x,
    u = (1 + x) cos(pi (1-y)),  f = -div(alpha grad u) + sigma u  expanded into polynomial x trig terms."""
    d, n, pts = _unit_square_neumann_top()
    pi = float(np.float32(math.pi))
    X, Y, ONE = {(1, 0): 1.0}, {(0, 1): 1.0}, {(0, 0): 1.0}
    omy = _padd(ONE, _pscale(Y, -1.0))                                    # 1 - y
    omy2 = _pmul(omy, omy)
    ax = _padd(ONE, _pscale(X, 0.3))                                      # 1 + 0.3 x
    ay = _padd(ONE, _pscale(omy2, 0.5))                                   # 1 + 0.5 (1-y)^2
    alpha = _pmul(ax, ay)
    alpha_x = _pscale(ay, 0.3)
    alpha_y = _pscale(_pmul(ax, omy), -1.0)
    opx = _padd(ONE, X)                                                   # 1 + x
    sigma = {(0, 0): 0.5, (1, 0): 1.0}
    # u_x = C, u_y = pi (1+x) S, u_yy = -pi^2 (1+x) C   with C = cos(pi (1-y)), S = sin(pi (1-y))
    cos_poly = _padd(_pscale(alpha_x, -1.0), _pscale(_pmul(alpha, opx), math.pi ** 2), _pmul(sigma, opx))
    sin_poly = _pscale(_pmul(alpha_y, opx), -math.pi)
    terms = [make_term(A=c, px=i, py=j, trig1=("cos", 0.0, -pi, pi)) for (i, j), c in cos_poly.items()]
    terms += [make_term(A=c, px=i, py=j, trig1=("sin", 0.0, -pi, pi)) for (i, j), c in sin_poly.items()]
    g = TermField(0.0, [make_term(A=1.0, trig1=("cos", 0.0, -pi, pi)), make_term(A=1.0, px=1, trig1=("cos", 0.0, -pi, pi))])
    return Scenario(name="phys_varcoef_neumann_top", dirichlet=d, neumann=n, points=pts, g=g, f=TermField(0.0, terms),
                    alpha=TermField.polynomial(alpha), sigma=TermField.polynomial(sigma), n_walks=20000, max_steps=2000,
                    eps=1e-4, compat="physical", analytic=lambda p: (1.0 + p[:, 0]) * torch.cos(math.pi * (1.0 - p[:, 1])))


def phys_dcr_halfspace(n_electrodes: int = 9) -> Scenario:
    """DC resistivity the way the physics has it: half-space x in [-100, 100], y in [-100, 0] with an insulating surface
    y = 0 (zero Neumann) and u = 0 on the far boundaries; conductivity 1 S/m with a smooth conductive body
    (+4 S/m, Gaussian, 6 m wide) 25 m below the surface -- deep enough that d(alpha)/dy = 0 at the surface; a current
    dipole (Gaussian electrodes, 2 m wide, 2 m deep) at x = -20 / +20; potential electrodes 1 m below the surface.
    No analytic solution: checked against a finite-difference solve (tests/fd_reference.py)."""
    d = torch.tensor([[-100.0, 0.0], [-100.0, -100.0], [100.0, -100.0], [100.0, 0.0]])      # left, bottom, right
    n = torch.tensor([[100.0, 0.0], [-100.0, 0.0]])                                         # the surface
    xs = torch.linspace(-40.0, 40.0, n_electrodes)
    pts = torch.stack([xs, torch.full_like(xs, -1.0)], dim=1).contiguous()
    w, amp = 2.0, 1.0 / (2.0 * math.pi * 2.0 ** 2)
    q = 1.0 / (2.0 * w * w)
    f = TermField.gaussian_sum([(amp, (-20.0, -2.0), q), (-amp, (20.0, -2.0), q)])
    alpha = TermField.gaussian_sum([(4.0, (10.0, -25.0), 1.0 / (2.0 * 6.0 ** 2))], base=1.0)
    return Scenario(name="phys_dcr_halfspace", dirichlet=d, neumann=n, points=pts, g=None, f=f, alpha=alpha, sigma=None,
                    n_walks=20000, max_steps=20000, eps=1e-2, compat="physical")


PHYSICAL_VARCOEF = {"phys_varcoef_dirichlet": phys_varcoef_dirichlet, "phys_varcoef_neumann": phys_varcoef_neumann_top}
PHYSICAL = {"phys_laplace": phys_laplace_neumann_top, "phys_poisson": phys_poisson_neumann_top, "phys_cylinder": phys_cylinder}
ALL = {"cfg1a": cfg1a, "cfg1b": cfg1b, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
