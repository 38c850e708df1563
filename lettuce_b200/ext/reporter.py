"""Observables and the observable reporter (API of lettuce/ext/_reporter/observable_reporter.py).

Every observable is ONE pass over the populations on the CUDA engine (`lbm_reduce`: warp-shuffle
block reductions, deterministic two-stage fold) instead of the reference's chain of full-size
torch temporaries; results stay on the device as 0-d float64 tensors until a reporter reads them.
"""
from __future__ import annotations

import sys
from abc import ABC, abstractmethod
from typing import Optional

import torch

from .. import native
from .._simulation import Reporter

__all__ = ["Observable", "ObservableReporter", "MaximumVelocity", "IncompressibleKineticEnergy",
           "Enstrophy", "Mass"]


class Observable(ABC):
    def __init__(self, flow):
        self.context = flow.context
        self.flow = flow

    @abstractmethod
    def __call__(self, f: Optional[torch.Tensor] = None):
        ...


class MaximumVelocity(Observable):
    """max |u| in physical units (observable_reporter.py:27-31)"""

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        return self.flow.units.convert_velocity_to_pu(native.reduce(self.flow.stencil, native.MAX_U, f))


class IncompressibleKineticEnergy(Observable):
    """sum 0.5 |u|^2 dx^d in physical units (observable_reporter.py:34-42)"""

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        units = self.flow.units
        e_lu = native.reduce(self.flow.stencil, native.SUM_HALF_U2, f)
        return units.convert_incompressible_energy_to_pu(e_lu) * units.convert_length_to_pu(1.0) ** self.flow.stencil.d


class Enstrophy(Observable):
    """sum |curl u|^2 dx^d in physical units with 6th-order periodic differences
    (observable_reporter.py:45-68).  Only meaningful on periodic domains."""

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        units, st = self.flow.units, self.flow.stencil
        _, u = native.moments(st, f, want_rho=False)
        w2_lu = native.reduce(st, native.ENSTROPHY, u)          # lattice units, dx = 1
        dx = units.convert_length_to_pu(1.0)
        scale = units.convert_velocity_to_pu(1.0) / dx           # d(u_pu)/d(x_pu) per d(u_lu)/d(x_lu)
        return w2_lu * scale ** 2 * dx ** st.d


class Mass(Observable):
    """total mass in lattice units; like the reference it skips the first and last index of the
    last two axes and subtracts the populations on `no_mass_mask` (observable_reporter.py:140-158)"""

    def __init__(self, flow, no_mass_mask=None):
        super().__init__(flow)
        self.mask = no_mass_mask

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        mass = native.reduce(self.flow.stencil, native.SUM_F_INNER, f)
        if self.mask is not None:
            mass = mass - native.reduce(self.flow.stencil, native.SUM_F_MASKED, f, self.mask)
        return mass


class ObservableReporter(Reporter):
    """Evaluates `observable` every `interval` steps and prints or stores
    `[step, time_pu, value...]` (observable_reporter.py:161-200)."""
    batchable = True

    def __init__(self, observable, interval=1, out=sys.stdout):
        super().__init__(interval)
        self.observable = observable
        self.out = [] if out is None else out
        self._parameter_name = observable.__class__.__name__
        if out is not None:
            print("steps    ", "time    ", self._parameter_name)

    def __call__(self, simulation):
        if simulation.flow.i % self.interval != 0:
            return
        observed = self.observable.context.convert_to_ndarray(self.observable(simulation.flow.f))
        assert len(observed.shape) < 2
        observed = [observed.item()] if len(observed.shape) == 0 else observed.tolist()
        entry = [simulation.flow.i, simulation.units.convert_time_to_pu(simulation.flow.i)] + observed
        if isinstance(self.out, list):
            self.out.append(entry)
        else:
            print(*entry, file=self.out)
