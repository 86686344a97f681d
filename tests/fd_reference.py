"""Finite-difference reference for -div(alpha grad u) + sigma u = f on a rectangle (test infrastructure).

Node-centred 5-point scheme with the coefficient evaluated at the cell faces; homogeneous Dirichlet data on the left,
right and bottom edges, zero Neumann on the top edge (mirror node).  Used to check the physical-mode estimator on
problems without an analytic solution (scenarios.phys_dcr_halfspace)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch


def _on_grid(field, X, Y):
    if field is None:
        return None
    P = torch.from_numpy(np.stack([X.ravel(), Y.ravel()], axis=1)).float()
    return field(P).double().numpy().reshape(X.shape)


def solve_rectangle(x0, x1, y0, y1, h, alpha, f, sigma=None):
    """Returns (xs, ys, U) with U[i, j] = u(xs[i], ys[j]); top edge y = y1 is the Neumann edge."""
    nx, ny = int(round((x1 - x0) / h)) + 1, int(round((y1 - y0) / h)) + 1
    xs, ys = np.linspace(x0, x1, nx), np.linspace(y0, y1, ny)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    ax = _on_grid(alpha, X[:-1] + 0.5 * h, Y[:-1])             # faces between (i, j) and (i+1, j)
    ay = _on_grid(alpha, X[:, :-1], Y[:, :-1] + 0.5 * h)       # faces between (i, j) and (i, j+1)
    F = _on_grid(f, X, Y)
    S = _on_grid(sigma, X, Y) if sigma is not None else np.zeros_like(F)
    idx = -np.ones((nx, ny), dtype=np.int64)
    interior = np.zeros((nx, ny), dtype=bool)
    interior[1:-1, 1:] = True                                   # unknowns: all but left / right / bottom edges
    idx[interior] = np.arange(interior.sum())
    rows, cols, vals = [], [], []
    rhs = np.zeros(interior.sum())
    for i in range(1, nx - 1):
        for j in range(1, ny):
            k = idx[i, j]
            aw, ae = ax[i - 1, j], ax[i, j]
            as_ = ay[i, j - 1]
            an = ay[i, j] if j < ny - 1 else ay[i, j - 1]       # mirror node across the Neumann edge
            diag = (aw + ae + as_ + an) / h ** 2 + S[i, j]
            rhs[k] = F[i, j]
            for (ii, jj, a) in ((i - 1, j, aw), (i + 1, j, ae), (i, j - 1, as_), (i, j + 1 if j < ny - 1 else j - 1, an)):
                if idx[ii, jj] >= 0:
                    rows.append(k); cols.append(idx[ii, jj]); vals.append(-a / h ** 2)
            rows.append(k); cols.append(k); vals.append(diag)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(len(rhs), len(rhs)))
    u = spla.spsolve(A.tocsc(), rhs)
    U = np.zeros((nx, ny))
    U[interior] = u
    return xs, ys, U


def interpolate(xs, ys, U, pts):
    from scipy.interpolate import RegularGridInterpolator
    return RegularGridInterpolator((xs, ys), U)(np.asarray(pts, dtype=np.float64))
