#!/usr/bin/env python
"""Coefficients of include/wost_math.h (minimax-like fits by Lawson-reweighted least squares on Chebyshev nodes).

    python tools/gen_math_coeffs.py

Prints the float32 coefficients of: (sin r / r - 1)/r^2 and (cos r - 1)/r^2 on |r| <= pi/4 (in t = r^2),
(e^r - 1 - r)/r^2 on |r| <= ln2/2, and h(w) = 1/(e^-z I0(z) sqrt z), w = 1/z, on 2.9 <= z <= 21.5."""
import numpy as np
from scipy.special import i0e


def lawson(fn, lo, hi, deg, iters=60, N=6000, rel=False):
    k = np.arange(N)
    t = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / N)
    y = fn(t)
    w = np.ones(N)
    s = 1 / np.abs(y) if rel else np.ones(N)
    V = np.vander(t, deg + 1, increasing=True)
    for _ in range(iters):
        W = np.sqrt(w) * s
        c, *_ = np.linalg.lstsq(V * W[:, None], y * W, rcond=None)
        e = np.abs((V @ c - y) * s)
        w = w * e
        w /= w.sum()
    return c, e.max()


def S(t):
    r = np.sqrt(np.maximum(t, 1e-300))
    return np.where(t < 1e-8, -1 / 6 + t / 120, (np.sin(r) / r - 1) / t)


def Cc(t):
    r = np.sqrt(np.maximum(t, 1e-300))
    return np.where(t < 1e-6, -0.5 + t / 24, (np.cos(r) - 1) / t)


def E(r):
    return np.where(np.abs(r) < 1e-5, 0.5 + r / 6, (np.exp(r) - 1 - r) / np.where(r == 0, 1, r * r))


if __name__ == "__main__":
    T, L = (np.pi / 4) ** 2 * 1.02, np.log(2) / 2 * 1.01
    for name, (c, e) in {"sin": lawson(S, 0, T, 2, iters=30, N=4000), "cos": lawson(Cc, 0, T, 3, iters=30, N=4000),
                         "exp": lawson(E, -L, L, 4, iters=30, N=4000),
                         "1/I0": lawson(lambda w: 1.0 / (i0e(1 / w) * np.sqrt(1 / w)), 1 / 21.5, 1 / 2.9, 7, rel=True)}.items():
        print(name, "max fit error %.2e" % e, ["%.9g" % np.float32(v) for v in c])
