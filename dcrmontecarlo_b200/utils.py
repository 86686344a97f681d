"""Host-side helpers with the reference's names (reference ``utils.py:11-129``).

These run at solver construction only (sigma', sigma_bar); nothing here is on the per-step path.
The plotting helpers of the reference (``utils.py:237-638``) are visualisation only and not provided.
"""
from __future__ import annotations

import torch


def torchGradient(function, point: torch.Tensor) -> torch.Tensor:
    """Gradient of a scalar ``function`` at ``point`` by autograd, graph kept for second derivatives
    (reference utils.py:11-33)."""
    if not point.requires_grad:
        point = point.clone().requires_grad_(True)
    value = function(point)
    if value.numel() != 1:
        raise ValueError(f"Function must return a scalar, got tensor with {value.numel()} elements")
    (grad,) = torch.autograd.grad(value, point, create_graph=True)
    return grad


def torchLaplacian(function, point: torch.Tensor) -> torch.Tensor:
    """Sum of unmixed second derivatives plus the reference's ``1e-8`` regulariser; a failing second
    differentiation (e.g. a linear function, whose gradient has no graph) returns what has been
    accumulated so far, as the reference does (utils.py:35-63)."""
    if not point.requires_grad:
        point = point.clone().requires_grad_(True)
    grad = torchGradient(function, point)
    lap = torch.zeros_like(grad[0]) + 1e-8
    try:
        for i in range(len(grad)):
            lap = lap + torch.autograd.grad(grad[i], point, create_graph=True, retain_graph=True)[0][i]
    except Exception:
        return lap
    return lap


def gridSampleMinMax(function, domain_bounds: list, grid_resolution: int = 100) -> tuple:
    """Min / max of ``function`` over a regular lattice on a box of 1-3 dimensions, skipping points where it
    fails or is not finite; returns ``(min, max, argmin point, argmax point)`` (reference utils.py:65-120)."""
    axes = [torch.linspace(float(b[0]), float(b[1]), grid_resolution) for b in domain_bounds]
    if not 1 <= len(axes) <= 3:
        raise ValueError(f"Grid sampling for {len(axes)}D not implemented. Maximum supported dimension is 3.")
    mesh = torch.meshgrid(*axes, indexing="ij")
    grid_points = torch.stack([m.flatten() for m in mesh], dim=1)
    vals, kept = [], []
    for i, p in enumerate(grid_points):
        try:
            v = function(p)
            if torch.isnan(v) or torch.isinf(v):
                continue
            vals.append(v.item() if hasattr(v, "item") else float(v))
            kept.append(i)
        except Exception:
            continue
    if not vals:
        raise ValueError("Function could not be evaluated at any grid points")
    vals = torch.tensor(vals)
    # the reference indexes grid_points with positions in the *filtered* list (utils.py:112-118); kept as is
    lo, hi = int(torch.argmin(vals)), int(torch.argmax(vals))
    return vals[lo].item(), vals[hi].item(), grid_points[lo], grid_points[hi]


def torch_smooth_circle(x: torch.Tensor, center, radius):
    """Differentiable indicator of a disc: ``sigmoid(-100 (|x - c| - R))`` (reference utils.py:123-129).
    On the device this is the WOST_TERM_SIGMOID_CIRCLE term (``fields.TermField.smooth_circle_sum``)."""
    return (-100 * ((x - center).norm() - radius)).sigmoid()
