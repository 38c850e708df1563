"""Small helpers of lettuce/util/utility.py that user scripts import from the package root: exception and
warning classes, periodic finite differences, the Jacobi pressure solver, grid coarsening.  Plain torch on
whatever device the arguments live on; none of this is on the time-step path."""
from __future__ import annotations

import inspect

import torch

__all__ = ["get_subclasses", "LettuceException", "LettuceWarning", "InefficientCodeWarning", "ExperimentalWarning",
           "torch_gradient", "grid_fine_to_coarse", "torch_jacobi", "append_axes"]


def get_subclasses(cls, module):
    for _, obj in inspect.getmembers(module):
        if hasattr(obj, "__bases__") and cls in obj.__bases__:
            yield obj


class LettuceException(Exception):
    pass


class LettuceWarning(UserWarning):
    pass


class InefficientCodeWarning(LettuceWarning):
    pass


class ExperimentalWarning(LettuceWarning):
    pass


_FD_WEIGHTS = {2: ((1, -1 / 2), (-1, 1 / 2)),
               4: ((2, 1 / 12), (1, -2 / 3), (-1, 2 / 3), (-2, -1 / 12)),
               6: ((3, -1 / 60), (2, 3 / 20), (1, -3 / 4), (-1, 3 / 4), (-2, -3 / 20), (-3, 1 / 60))}


def torch_gradient(f: torch.Tensor, dx=1, order: int = 2) -> torch.Tensor:
    """First derivatives of a periodic 2-D or 3-D field along every axis, central differences of order 2, 4 or 6
    (Fornberg 1988; lettuce/util/utility.py:37-99).  Returns `[ndim, *f.shape]`."""
    if f.dim() not in (2, 3):
        raise LettuceException("Invalid dimension!")
    if order not in _FD_WEIGHTS:
        raise LettuceException(f"order must be 2, 4 or 6, got {order}")
    scale = torch.tensor(1.0 / dx, dtype=f.dtype, device=f.device)
    with torch.no_grad():
        return torch.stack([sum(w * f.roll(shifts=s, dims=ax) for s, w in _FD_WEIGHTS[order]) * scale
                            for ax in range(f.dim())])


def torch_jacobi(f, p, dx, dim, tol_abs=1e-10, max_num_steps=100000):
    """Jacobi iterations for lap p = f on a periodic grid until the mean squared residual drops below `tol_abs`
    (lettuce/util/utility.py:119-156)"""
    if dim not in (2, 3):
        raise LettuceException("Invalid dimension!")
    dims = range(dim)
    neighbours = lambda q: sum(q.roll(shifts=1, dims=a) + q.roll(shifts=-1, dims=a) for a in dims)
    error, it = 1.0, 0
    while error > tol_abs and it < max_num_steps:
        it += 1
        p = (f * dx ** 2 - neighbours(p)) * -1 / (2 * dim)
        residuum = f - (neighbours(p) - 2 * dim * p) / dx ** 2
        error = float(torch.mean(residuum ** 2))
    return p


def grid_fine_to_coarse(flow, f_fine, tau_fine, tau_coarse):
    """every second node of `f_fine` with the non-equilibrium part rescaled by 2 tau_coarse / tau_fine
    (lettuce/util/utility.py:101-116)"""
    if f_fine.dim() == 3:
        coarse = f_fine[:, ::2, ::2]
    elif f_fine.dim() == 4:
        coarse = f_fine[:, ::2, ::2, ::2]
    else:
        raise LettuceException("Invalid dimension!")
    coarse = coarse.contiguous()
    f_eq = flow.equilibrium(flow, rho=flow.rho(coarse), u=flow.u(coarse))
    return f_eq + 2 * tau_coarse / tau_fine * (coarse - f_eq)


def append_axes(array, n):
    return array[(Ellipsis,) + (None,) * n]
