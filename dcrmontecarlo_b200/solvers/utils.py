"""Green's-function helpers and radius samplers (reference ``solvers/utils.py``).

The reference draws source radii from rejection-sampled, cycled caches of 10 000 values
(``solvers/utils.py:64-195``).  The CUDA kernel draws from the *same distributions* without a cache:

* Green's radius, pdf ``-ln(rho)`` on (1e-6, 1) (``:138-151``): the product of two uniforms has
  exactly that density;
* screened radius (``:181-195``): the rejection sampler's envelope is not an upper bound, so the
  density it realises is ``min(|G^sb(rho; R=1)|, envelope)``; :func:`screened_radius_icdf` tabulates its
  inverse CDF once per ``sigma_bar`` and the kernel interpolates it.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import i0, k0


def screenedGreens2D(x: torch.Tensor, y: torch.Tensor, R: float, sigmaBar: float) -> float:
    """Screened (Yukawa) Green's function of the disc of radius R, pole at the centre (reference :5-26)."""
    r = float((x - y).norm())
    s = np.sqrt(sigmaBar)
    return 1.0 / (2.0 * np.pi) * (k0(r * s) - (k0(R * s) / i0(R * s)) * i0(r * s))


def screenedGreensNorm2D(R: float, sigmaBar: float) -> float:
    """Integral of the screened Green's function over the disc: ``(1 - 1/I0(R sqrt(sb))) / sb`` (reference :29-44)."""
    return 1.0 / sigmaBar * (1.0 - 1.0 / i0(float(R) * np.sqrt(sigmaBar)))


def greensFunction2D(x: torch.Tensor, y: torch.Tensor, R: float) -> float:
    """``-ln|x-y| / 2pi`` (reference :46-54)."""
    r = (x - y).norm()
    if r < 1e-10:
        return 0.0
    return -1.0 / (2.0 * np.pi) * torch.log(r)


def greensFunctionNorm2D(R: float) -> float:
    """``R^2 / 4`` (reference :56-61)."""
    return R ** 2 / 4


def screened_density(rho: np.ndarray, sigma_bar: float) -> np.ndarray:
    """Unnormalised density of the reference's screened radius sampler on (1e-6, 1): |G^sb(rho; R=1)| clipped at
    the rejection envelope screenedGreensNorm2D(1, sb) (reference :184-194)."""
    s = np.sqrt(sigma_bar)
    g = np.abs(1.0 / (2.0 * np.pi) * (k0(rho * s) - (k0(s) / i0(s)) * i0(rho * s)))
    return np.minimum(g, screenedGreensNorm2D(1.0, sigma_bar))


def screened_radius_icdf(sigma_bar: float, n: int = 1024, quad: int = 16384) -> np.ndarray:
    """``table[i]`` = normalised radius at cumulative probability ``i/(n-1)``; the kernel draws
    ``rho = lerp(table, u (n-1))`` and scales by the star radius (reference :109-117)."""
    s = np.linspace(0.0, 1.0, quad + 1)
    x = 1e-6 + (1.0 - 1e-6) * s * s                      # nodes graded towards the log-singular end
    d = screened_density(x, float(sigma_bar))
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (d[1:] + d[:-1]) * np.diff(x))])
    u = np.linspace(0.0, cdf[-1], n)
    return np.interp(u, cdf, x).astype(np.float32)


class SamplingDistribution2D:
    """Radius distributions for Green's-function sampling (reference :64-117).  ``sample`` draws fresh
    values from the exact distribution instead of cycling a cache."""

    def __init__(self, cache_size: int = 10000):
        self.cache_size = cache_size

    def sample(self, center: torch.Tensor, radius: float) -> float:
        raise NotImplementedError

    def pdf(self, r: float, center: torch.Tensor, radius: float) -> float:
        raise NotImplementedError


class GreensDistribution2D(SamplingDistribution2D):
    def sample(self, center, radius):
        rho = max(np.random.uniform() * np.random.uniform(), 1e-6)
        return rho * radius

    def pdf(self, r, center, radius):
        if r <= 0 or r >= radius:
            return 0.0
        return -np.log(r / radius) / (radius ** 2 / 4)


class ScreenedGreensDistribution2D(SamplingDistribution2D):
    def __init__(self, sigma_bar: float, cache_size: int = 10000):
        super().__init__(cache_size)
        self.sigma_bar = sigma_bar
        self._icdf = None

    def sample(self, center, radius):
        if self._icdf is None:
            self._icdf = screened_radius_icdf(self.sigma_bar)
        pos = np.random.uniform() * (len(self._icdf) - 1)
        i = min(int(pos), len(self._icdf) - 2)
        return (self._icdf[i] + (pos - i) * (self._icdf[i + 1] - self._icdf[i])) * radius

    def pdf(self, r, center, radius):
        if r <= 0 or r >= radius:
            return 0.0
        return abs(screenedGreens2D(torch.zeros(2), torch.tensor([r, 0.0]), radius, self.sigma_bar)) / screenedGreensNorm2D(radius, self.sigma_bar)


# ---- multiple-importance-sampling helpers ----------------------------------------------------------------
# The reference ships these (solvers/utils.py:198-324) but neither its solver nor its tests call them; they are kept
# importable, host-side only, so code written against the reference's module keeps working.
class UniformDistribution2D(SamplingDistribution2D):
    """Radius uniform on [0, radius] (reference :198-217)."""

    def sample(self, center, radius):
        return np.random.uniform(0, radius)

    def pdf(self, r, center, radius):
        return 1.0 / radius if 0 <= r <= radius else 0.0


class MultipleImportanceSampler2D:
    """Mixture of radius distributions with balance-heuristic weights (reference :220-286)."""

    def __init__(self, distributions: list, weights: list = None):
        self.distributions = distributions
        w = np.asarray(weights if weights else [1.0] * len(distributions), dtype=float)
        self.weights = w / w.sum()

    def sample(self, center, radius):
        i = int(np.random.choice(len(self.distributions), p=self.weights))
        r = self.distributions[i].sample(center, radius)
        return r, i, self._compute_mis_weight(r, center, radius, i)

    def _compute_mis_weight(self, r, center, radius, sampled_idx):
        wp = self.weights * np.array([d.pdf(r, center, radius) for d in self.distributions])
        tot = wp.sum()
        return 0.0 if tot == 0 else float(wp[sampled_idx] / tot)


def sampleGreensFunction2D(center, radius, distribution: SamplingDistribution2D = None) -> float:
    """One Green's-function radius (reference :289-304)."""
    return (distribution or GreensDistribution2D()).sample(center, radius)


def sampleScreenedGreensFunction2D(center, radius, sigma_bar, distribution: "ScreenedGreensDistribution2D" = None) -> float:
    """One screened-Green's-function radius (reference :307-324)."""
    return (distribution or ScreenedGreensDistribution2D(sigma_bar)).sample(center, radius)
