"""Flow state and its moments (API of lettuce/_flow.py:17-367).

`Flow.f` is the dense population tensor `[q, *resolution]` in the reference's layout.  The
moment queries (`rho`, `j`, `u`, `incompressible_energy`) run on the CUDA engine
(`lbm_moments`); they raise for CPU tensors.  Construction-time work -- evaluating the initial
condition once -- is ordinary torch code on the context device and is not part of the hot path
(SURVEY.md section 2, rows 5 and 13).
"""
from __future__ import annotations

import pickle
import warnings
from abc import ABC, abstractmethod
from typing import List, Optional, Union

import numpy as np
import torch

from . import native
from ._stencil import TorchStencil

__all__ = ["Equilibrium", "Boundary", "Flow", "QuadraticEquilibrium", "QuadraticEquilibriumLessMemory",
           "initialize_f_neq", "pressure_poisson"]


class Equilibrium(ABC):
    @abstractmethod
    def __call__(self, flow: "Flow", rho=None, u=None) -> torch.Tensor:
        ...

    def native_available(self) -> bool:
        return False


class Boundary(ABC):
    """A boundary contributes a boolean `no_collision_mask` (where it replaces the collision)
    and optionally a `no_streaming_mask` (slots streaming must not overwrite), see
    lettuce/_flow.py:31-52.  On the B200 engine a boundary is a *parameter holder*: the kernel
    family in csrc/lbm_step.cuh implements the operator, selected by class name."""

    # boundaries whose reference implementation writes `flow.f` in place and returns it
    # (equilibrium_outlet_p.py:63-73, anti_bounce_back_outlet.py:71-91)
    in_place = False

    def __call__(self, flow: "Flow") -> torch.Tensor:
        """the boundary applied to every node it would act on if the whole lattice carried its label (the
        reference evaluates `boundary(flow)` on the full grid and blends by label afterwards,
        lettuce/_simulation.py:258-305); outlets rewrite their plane of `flow.f` in place like the reference"""
        out = native.apply_operator(self, flow, is_collision=False)
        if self.in_place:
            flow.f.copy_(out)
            return flow.f
        return out

    @abstractmethod
    def make_no_collision_mask(self, shape: List[int], context) -> Optional[torch.Tensor]:
        ...

    @abstractmethod
    def make_no_streaming_mask(self, shape: List[int], context) -> Optional[torch.Tensor]:
        ...

    def native_available(self) -> bool:
        try:
            native.op_kind(self)
            return True
        except NotImplementedError:
            return False


class QuadraticEquilibrium(Equilibrium):
    """feq_q = w_q rho (1 + e.u/cs^2 + (e.u)^2/(2 cs^4) - u.u/(2 cs^2))
    (lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24).

    This tensor-level evaluation serves initial conditions and user code that asks for an
    equilibrium field; the time step evaluates the same polynomial per node inside the fused
    kernel (csrc/lbm_core.cuh `Equilibrium`)."""

    def __call__(self, flow: "Flow", rho=None, u=None) -> torch.Tensor:
        rho = flow.rho() if rho is None else rho
        u = flow.u() if u is None else u
        st = flow.torch_stencil
        cs2 = st.cs ** 2
        eu = torch.tensordot(st.e, u, dims=1)
        uu = (u * u).sum(dim=0)
        poly = (2 * eu - uu) / (2 * cs2) + 0.5 * (eu / cs2) ** 2 + 1
        w = st.w.reshape([-1] + [1] * (eu.dim() - 1))
        return w * (rho * poly)

    def native_available(self) -> bool:
        return True


class QuadraticEquilibriumLessMemory(QuadraticEquilibrium):
    """The reference's lower-memory evaluation order of the same polynomial
    (lettuce/ext/_equilibrium/quadratic_equilibrium_less_memory.py:8-32); here both names share one
    implementation (the step kernels never materialise an equilibrium tensor)."""


class Flow(ABC):
    """Physical configuration and state of a simulation (lettuce/_flow.py:55-268)."""

    initialize_pressure: bool = False
    initialize_fneq: bool = False

    def __init__(self, context, resolution: List[int], units, stencil, equilibrium=None):
        self.context = context
        self.resolution = list(resolution)
        self.units = units
        self.stencil = stencil() if callable(stencil) else stencil
        self.torch_stencil = TorchStencil(self.stencil, context)
        self.equilibrium = equilibrium or QuadraticEquilibrium()
        self.i = 0
        self.f = context.empty_tensor([self.stencil.q, *self.resolution])
        self._f_next = None
        self.initialize()

    # boundaries -------------------------------------------------------------
    @property
    def pre_boundaries(self) -> List[Boundary]:
        return []

    @property
    def post_boundaries(self) -> List[Boundary]:
        warnings.warn("post_boundaries is not defined by this flow; falling back to the deprecated "
                      "`boundaries` property (lettuce/_flow.py:103-114).", stacklevel=2)
        return self.boundaries

    @property
    def boundaries(self) -> List[Boundary]:
        return []

    @abstractmethod
    def initial_pu(self):
        """(pressure, velocity) of the initial state in physical units"""

    def initialize(self):
        """Equilibrium initialisation from `initial_pu`, optionally with first-order
        non-equilibrium (lettuce/_flow.py:127-143)."""
        p, u = self.initial_pu()
        rho = self.context.convert_to_tensor(self.units.convert_pressure_pu_to_density_lu(p))
        u = self.context.convert_to_tensor(self.units.convert_velocity_to_lu(u))
        if self.initialize_pressure:
            rho = pressure_poisson(self.units, u, rho)
        if u.is_cuda and type(self.equilibrium) is QuadraticEquilibrium and u.dim() == self.stencil.d + 1:
            # one kernel, no full-size temporaries (the torch expression needs ~6x the size of f)
            self.f = native.equilibrium_field(self.stencil, rho, u, self.resolution)
        else:
            self.f = self.equilibrium(self, rho=rho, u=u).contiguous()
        if self.initialize_fneq:
            self.f = initialize_f_neq(self).contiguous()
        self._f_next = None

    # double buffer ------------------------------------------------------------
    @property
    def f_next(self) -> torch.Tensor:
        if self._f_next is None or self._f_next.shape != self.f.shape or self._f_next.device != self.f.device:
            self._f_next = torch.empty_like(self.f)
        return self._f_next

    @f_next.setter
    def f_next(self, value: torch.Tensor):
        self._f_next = value

    # moments (CUDA engine) ------------------------------------------------------
    def rho(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """density [1, *res] (lettuce/_flow.py:157-159)"""
        return native.moments(self.stencil, self.f if f is None else f, want_u=False)[0]

    def u(self, f: Optional[torch.Tensor] = None, rho=None, acceleration=None) -> torch.Tensor:
        """velocity [d, *res] (lettuce/_flow.py:178-193)"""
        if acceleration is None:
            return native.moments(self.stencil, self.f if f is None else f, want_rho=False)[1]
        # with a forcing scheme the physical velocity averages pre- and post-collision momentum
        rho_, v = native.moments(self.stencil, self.f if f is None else f)
        rho_ = rho_ if rho is None else rho
        acceleration = self.context.convert_to_tensor(acceleration)
        if acceleration.dim() == 1:
            acceleration = acceleration.reshape([-1] + [1] * self.stencil.d)
        return v + acceleration / (2 * rho_)

    def j(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """momentum [d, *res] (lettuce/_flow.py:173-176)"""
        rho, u = native.moments(self.stencil, self.f if f is None else f)
        return u * rho

    @property
    def velocity(self):
        return self.u()

    @property
    def rho_pu(self):
        return self.units.convert_density_to_pu(self.rho())

    @property
    def p_pu(self):
        return self.units.convert_density_lu_to_pressure_pu(self.rho())

    @property
    def u_pu(self):
        return self.units.convert_velocity_to_pu(self.u())

    def incompressible_energy(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """0.5 |u|^2 per node (lettuce/_flow.py:200-204)"""
        u = self.u(f)
        return 0.5 * (u * u).sum(dim=0)

    # diagnostics (lettuce/_flow.py:206-256): plain torch reductions over q, not on the step path -------------
    def entropy(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """entropy according to the H-theorem, as the reference evaluates it (_flow.py:206-211)"""
        f = self.f if f is None else f
        w = self.torch_stencil.w.reshape([-1] + [1] * self.stencil.d)
        return (f * -torch.log((f / w).sum(dim=0))).sum(dim=0)

    def pseudo_entropy_global(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Taylor expansion of the entropy around the weights (_flow.py:213-218; like the reference the density
        term is that of `flow.f`)"""
        f = self.f if f is None else f
        w = self.torch_stencil.w.reshape([-1] + [1] * self.stencil.d)
        return self.rho() - (f * f / w).sum(dim=0)

    def pseudo_entropy_local(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Taylor expansion of the entropy around the local equilibrium (_flow.py:220-226)"""
        f = self.f if f is None else f
        return self.rho(f) - (f * (f / self.equilibrium(self))).sum(dim=0)

    def shear_tensor(self, f: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pi_ab = sum_q f_q e_qa e_qb, shape [d, d, *resolution] (_flow.py:228-234)"""
        e = self.torch_stencil.e
        return torch.einsum("q...,qab->ab...", self.f if f is None else f, torch.einsum("qa,qb->qab", e, e))

    def einsum(self, equation, fields, *args) -> torch.Tensor:
        """deprecated: Einstein summation that appends the spatial axes (_flow.py:236-256)"""
        warnings.warn("The `einsum` method is deprecated and will be removed in a future version. "
                      "Please use `torch.einsum` directly instead.", DeprecationWarning, stacklevel=2)
        inputs, output = equation.split("->")
        inputs = inputs.split(",")
        for i, inp in enumerate(inputs):
            if len(inp) == len(fields[i].shape) - self.stencil.d:
                inputs[i] += "..."
                if not output.endswith("..."):
                    output += "..."
            else:
                assert len(inp) == len(fields[i].shape), "Bad dimension."
        return torch.einsum(",".join(inputs) + "->" + output, fields, *args)

    # checkpoint (lettuce/_flow.py:258-268) -----------------------------------
    def dump(self, filename):
        with open(filename, "wb") as fh:
            pickle.dump(self.context.convert_to_ndarray(self.f), fh)

    def load(self, filename):
        with open(filename, "rb") as fh:
            self.f = self.context.convert_to_tensor(pickle.load(fh), dtype=self.context.dtype).contiguous()
        self._f_next = None


def _gradient6(a: torch.Tensor) -> torch.Tensor:
    """6th-order periodic central differences with unit spacing along every axis
    (lettuce/util/utility.py:37-99, order=6).  Initial-condition helper."""
    weights = (-1 / 60, 3 / 20, -3 / 4, 3 / 4, -3 / 20, 1 / 60)
    shifts = (3, 2, 1, -1, -2, -3)
    return torch.stack([sum(w * a.roll(s, dims=ax) for w, s in zip(weights, shifts)) for ax in range(a.dim())])


def _gradient2(a: torch.Tensor, dx) -> torch.Tensor:
    """2nd-order periodic central differences along every axis (lettuce/util/utility.py:37-99, order=2)"""
    return torch.stack([(-0.5 * a.roll(1, dims=ax) + 0.5 * a.roll(-1, dims=ax)) for ax in range(a.dim())]) \
        * torch.tensor(1.0 / dx, dtype=a.dtype, device=a.device)


def pressure_poisson(units, u: torch.Tensor, rho0: torch.Tensor, tol_abs=1e-10, max_num_steps=100000) -> torch.Tensor:
    """Density whose pressure solves lap p = -d_i d_j (u_i u_j), by Jacobi iteration on the periodic lattice
    (lettuce/_flow.py:271-320 with util/utility.py:119-156).  Like the reference it treats the first two axes
    (`dim=2`) -- "still not working in 3D" there.  Initial-condition helper: plain torch on the context device."""
    dx = units.convert_length_to_pu(1.0)
    u = units.convert_velocity_to_pu(u)
    p = units.convert_density_lu_to_pressure_pu(rho0)[0]
    rhs = torch.zeros_like(u[0])
    for i in range(u.shape[0]):
        for j in range(u.shape[0]):
            rhs -= _gradient2(_gradient2(u[i] * u[j], dx)[i], dx)[j]
    neighbours = lambda q: q.roll(1, dims=0) + q.roll(1, dims=1) + q.roll(-1, dims=0) + q.roll(-1, dims=1)
    error, it = 1.0, 0
    while error > tol_abs and it < max_num_steps:
        it += 1
        p = (rhs * dx ** 2 - neighbours(p)) * -1 / 4
        residuum = rhs - (neighbours(p) - 4 * p) / dx ** 2
        error = float(torch.mean(residuum ** 2))
    return units.convert_pressure_pu_to_density_lu(p[None, ...])


def initialize_f_neq(flow: Flow) -> torch.Tensor:
    """feq - w_i Q_i : Pi^(1) with Pi^(1) = tau rho grad(u) / cs^2 (lettuce/_flow.py:341-367,
    Krueger et al. 2017).  Runs once at construction, in torch on the context device.

    The reference forms cs^2 * identity in torch's default dtype (float32) even for a float64
    context (_flow.py:358-360); that rounding is reproduced so that initial states are
    identical to the reference's."""
    f, st, d = flow.f, flow.torch_stencil, flow.stencil.d
    if f.is_cuda and type(flow.equilibrium) is QuadraticEquilibrium and f.dtype in (torch.float32, torch.float64):
        # one kernel, no full-size temporaries (row f3 of SURVEY.md section 8)
        return native.initialize_fneq(flow.stencil, f.contiguous(), flow.units.relaxation_parameter_lu)
    rho = f.sum(dim=0, keepdim=True)
    u = torch.tensordot(st.e.t().contiguous(), f, dims=1) / rho
    grad_u = torch.stack([_gradient6(u[a]) for a in range(d)])
    pi1 = flow.units.relaxation_parameter_lu * rho * grad_u / st.cs ** 2
    del grad_u
    eye = torch.eye(d, device=f.device) * flow.stencil.cs ** 2
    Q = torch.einsum("ia,ib->iab", st.e, st.e) - eye
    fneq = torch.einsum("ab...,iab->i...", pi1, Q)
    del pi1
    fneq.mul_(st.w.reshape([-1] + [1] * d))
    if f.is_cuda and type(flow.equilibrium) is QuadraticEquilibrium:
        feq = native.equilibrium_field(flow.stencil, rho, u, f.shape[1:])     # no full-size temporaries
    else:
        feq = flow.equilibrium(flow, rho, u)
    return feq.sub_(fneq)
