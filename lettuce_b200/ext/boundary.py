"""Boundary conditions (API of lettuce/ext/_boundary/*).

Each class keeps the reference's constructor and mask methods; the operator itself is
implemented in csrc/lbm_step.cuh (`general_node`) and selected by class name.
"""
from __future__ import annotations

import numbers
from typing import List, Optional

import numpy as np
import torch

from .._flow import Boundary

__all__ = ["BounceBackBoundary", "EquilibriumBoundaryPU", "EquilibriumOutletP", "AntiBounceBackOutlet"]


class BounceBackBoundary(Boundary):
    """full-way bounce-back on the masked (solid) nodes: f <- f[opposite]
    (lettuce/ext/_boundary/bounce_back_boundary.py:10-32)"""

    def __init__(self, mask: torch.Tensor):
        self._mask = mask

    def make_no_collision_mask(self, shape, context):
        return self._mask

    def make_no_streaming_mask(self, shape, context):
        return None


class EquilibriumBoundaryPU(Boundary):
    """Equilibrium with prescribed velocity and pressure (physical units) on the masked nodes
    (lettuce/ext/_boundary/equilibrium_boundary_pu.py:15-98)."""

    @staticmethod
    def checked_tensor(t, context, flow) -> torch.Tensor:
        """Bring a scalar, a length-d vector, a spatial field or a [d|1, N|1, ...] tensor to
        rank d+1 with broadcastable extents (equilibrium_boundary_pu.py:23-69)."""
        if not torch.is_tensor(t):
            if isinstance(t, (numbers.Number, np.ndarray, list, tuple)):
                t = torch.as_tensor(t)
            else:
                raise TypeError(f"Cannot convert {type(t)} to tensor")
        d = flow.stencil.d
        spatial = [int(s) for s in flow.f.shape[1:]]
        rank = d + 1
        if t.ndim == 0:
            t = t.reshape([1] * rank)
        elif t.ndim == 1 and t.shape[0] == d:
            t = t.reshape([d] + [1] * d)
        elif t.ndim == d and list(t.shape) == spatial:
            t = t.unsqueeze(0)
        elif t.ndim != rank:
            raise ValueError(f"Tensor has wrong rank: expected {rank}, got {t.ndim}")
        if t.shape[0] not in (1, d):
            raise ValueError(f"Component dim must be 1 or {d}, got {t.shape[0]}")
        for a in range(d):
            if t.shape[a + 1] not in (1, spatial[a]):
                raise ValueError(f"Spatial dim {a + 1} must be 1 or {spatial[a]}, got {t.shape[a + 1]}")
        return context.convert_to_tensor(t)

    def __init__(self, context, flow, mask, velocity, pressure=0):
        self.velocity = self.checked_tensor(velocity, context, flow)
        self.pressure = self.checked_tensor(pressure, context, flow)
        self._mask = mask

    def make_no_collision_mask(self, shape, context):
        return self._mask

    def make_no_streaming_mask(self, shape, context):
        return None


def _check_direction(direction):
    direction = [int(c) for c in direction]
    assert len(direction) in (1, 2, 3), \
        f"Invalid direction parameter. Expected direction of length 1, 2 or 3 but got {len(direction)}."
    assert direction.count(0) == len(direction) - 1 and ((1 in direction) ^ (-1 in direction)), \
        f"Invalid direction parameter. Expected direction with all entries 0 except one 1 or -1 but got {direction}."
    return direction


class _PlaneOutlet(Boundary):
    """shared geometry of the two outlets: the boundary plane `index` (last or first plane along
    the direction axis), its inward `neighbor` plane and the outgoing velocity set
    {q : e_q . direction = 1} (anti_bounce_back_outlet.py:38-55)."""

    in_place = True

    def __init__(self, direction, flow):
        self.direction = _check_direction(direction)
        e = np.asarray(flow.stencil.e)
        self.velocities = np.flatnonzero(e @ np.asarray(self.direction) > 1 - 1e-6)
        self.opposite_velocities = np.asarray(flow.stencil.opposite)[self.velocities]
        self.index = [slice(None) if c == 0 else (-1 if c == 1 else 0) for c in self.direction]
        self.neighbor = [slice(None) if c == 0 else (-2 if c == 1 else 1) for c in self.direction]

    def make_no_collision_mask(self, shape, context):
        mask = context.zero_tensor(shape, dtype=torch.bool)
        mask[tuple(self.index)] = True
        return mask

    def _frozen_populations(self, q):
        raise NotImplementedError

    def make_no_streaming_mask(self, shape, context):
        mask = context.zero_tensor(shape, dtype=torch.bool)
        mask[tuple([self._frozen_populations(shape[0])] + self.index)] = True
        return mask


class EquilibriumOutletP(_PlaneOutlet):
    """constant-pressure outlet: the whole plane is set to feq(rho_outlet, u of the neighbour
    plane); every population that does not leave through the plane is frozen during streaming
    (lettuce/ext/_boundary/equilibrium_outlet_p.py:12-91)."""

    def __init__(self, direction: List[int], flow, rho_outlet: float = 1.0):
        super().__init__(direction, flow)
        self.context = flow.context
        self.rho_outlet = float(rho_outlet)

    def _frozen_populations(self, q):
        return np.setdiff1d(np.arange(q), self.velocities)


class AntiBounceBackOutlet(_PlaneOutlet):
    """anti-bounce-back outlet (Krueger et al. 2017, p. 195): the populations opposite to the
    outgoing ones are rebuilt from the wall velocity u_w = u + (u - u_neighbour)/2 and frozen
    during streaming (lettuce/ext/_boundary/anti_bounce_back_outlet.py:13-109)."""

    def __init__(self, direction: List[int], flow, collision=None):
        super().__init__(direction, flow)
        self.collision = collision

    def _frozen_populations(self, q):
        return self.opposite_velocities
