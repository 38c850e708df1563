"""Flows of the benchmark configurations (API of lettuce/ext/_flows/{_ext_flow,taylorgreen,obstacle}.py).
Initial conditions are evaluated once, in torch, on the context device."""
from __future__ import annotations

import warnings
from abc import abstractmethod
from typing import List, Optional, Union

import numpy as np
import torch

from .._flow import Flow
from .._stencil import D2Q9, D3Q19
from ..units import UnitConversion
from .boundary import AntiBounceBackOutlet, BounceBackBoundary, EquilibriumBoundaryPU

__all__ = ["ExtFlow", "TaylorGreenVortex", "TaylorGreenVortex2D", "TaylorGreenVortex3D", "Obstacle", "PoiseuilleFlow2D", "Cavity2D", "DoublyPeriodicShear2D",
           "CouetteFlow2D", "LambOseenVortex2D", "DecayingTurbulence", "flow_by_name"]


class ExtFlow(Flow):
    """Common constructor: resolution + Reynolds + Mach + stencil (lettuce/ext/_flows/_ext_flow.py:8-42)."""

    def __init__(self, context, resolution: Union[int, List[int]], reynolds_number, mach_number,
                 stencil=None, equilibrium=None):
        resolution = self.make_resolution(resolution, stencil)
        assert len(resolution) in (2, 3), f"the B200 engine supports 2 and 3 dimensions, got {len(resolution)}"
        stencil = stencil or (D2Q9() if len(resolution) == 2 else D3Q19())
        stencil = stencil() if callable(stencil) else stencil
        Flow.__init__(self, context, resolution, self.make_units(reynolds_number, mach_number, resolution),
                      stencil, equilibrium)

    @abstractmethod
    def make_resolution(self, resolution, stencil=None) -> List[int]:
        ...

    @abstractmethod
    def make_units(self, reynolds_number, mach_number, resolution: List[int]) -> UnitConversion:
        ...


class TaylorGreenVortex(ExtFlow):
    """Taylor-Green vortex in 2-D and 3-D on the periodic box [0, 2 pi)^d
    (lettuce/ext/_flows/taylorgreen.py:16-98); f_neq initialisation is on by default."""

    def __init__(self, context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None,
                 initialize_fneq: bool = True):
        self.initialize_fneq = initialize_fneq
        if stencil is None and not isinstance(resolution, list):
            warnings.warn("Requiring information about dimensionality! Either via stencil or resolution. "
                          "Setting dimension to 2.", UserWarning)
            self.stencil = D2Q9()
        else:
            self.stencil = stencil() if callable(stencil) else stencil
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, self.stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        if isinstance(resolution, int):
            return [resolution] * self.stencil.d
        assert len(resolution) in (2, 3), "the resolution of a taylor-green-vortex must be 2- or 3-dimensional!"
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0] / (2 * torch.pi),
                              characteristic_length_pu=1, characteristic_velocity_pu=1)

    @property
    def grid(self):
        axes = [torch.linspace(0, 2 * torch.pi * (1 - 1 / n), steps=n, device=self.context.device,
                               dtype=self.context.dtype) for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    def initial_pu(self):
        return self.analytic_solution(t=0)

    def analytic_solution(self, t: float):
        if t > 0 and self.stencil.d > 2:
            warnings.warn("The analytic solution is only true for the 2D TGV!")
        g = self.grid
        nu = self.context.convert_to_tensor(self.units.viscosity_pu)
        if len(self.resolution) == 2:
            decay = torch.exp(-2 * nu * t)
            u = torch.stack([torch.cos(g[0]) * torch.sin(g[1]) * decay,
                             -torch.sin(g[0]) * torch.cos(g[1]) * decay])
            p = -torch.stack([0.25 * (torch.cos(2 * g[0]) + torch.cos(2 * g[1])) * torch.exp(-4 * nu * t)])
        else:
            u = torch.stack([torch.sin(g[0]) * torch.cos(g[1]) * torch.cos(g[2]),
                             -torch.cos(g[0]) * torch.sin(g[1]) * torch.cos(g[2]),
                             torch.zeros_like(g[0])])
            p = torch.stack([1 / 16. * (torch.cos(2 * g[0]) + torch.cos(2 * g[1])) * (torch.cos(2 * g[2]) + 2)])
        return p, u

    @property
    def post_boundaries(self):
        return []


def TaylorGreenVortex3D(context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None):
    """deprecated alias (lettuce/ext/_flows/taylorgreen.py:101-110)"""
    warnings.warn("TaylorGreenVortex3D is deprecated. Use TaylorGreenVortex instead", DeprecationWarning)
    return TaylorGreenVortex(context=context, resolution=resolution, reynolds_number=reynolds_number,
                             mach_number=mach_number, stencil=stencil, equilibrium=equilibrium)


def TaylorGreenVortex2D(context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None):
    """deprecated alias (lettuce/ext/_flows/taylorgreen.py:113-122)"""
    warnings.warn("TaylorGreenVortex2D is deprecated. Use TaylorGreenVortex instead", DeprecationWarning)
    return TaylorGreenVortex(context=context, resolution=resolution, reynolds_number=reynolds_number,
                             mach_number=mach_number, stencil=stencil, equilibrium=equilibrium)


class Obstacle(ExtFlow):
    """Flow in +x around a solid `mask`: equilibrium inlet at x = 0, outlet at x = nx-1,
    bounce-back on the mask (lettuce/ext/_flows/obstacle.py:16-125).  Set `flow.mask` after
    construction and call `flow.initialize()` to start from rest inside the solid."""

    def __init__(self, context, resolution, reynolds_number, mach_number, domain_length_x, char_length=1,
                 char_velocity=1, stencil=None, equilibrium=None):
        self.char_length_lu = resolution[0] / domain_length_x * char_length
        self.char_length = char_length
        self.char_velocity = char_velocity
        self.resolution = self.make_resolution(resolution, stencil)
        self._mask = torch.zeros(self.resolution, dtype=torch.bool, device=context.device)
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, stencil, equilibrium)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=self.char_length_lu,
                              characteristic_length_pu=self.char_length,
                              characteristic_velocity_pu=self.char_velocity)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        if isinstance(resolution, int):
            st = stencil() if callable(stencil) else stencil
            return [resolution] * st.d
        return list(resolution)

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, m):
        assert isinstance(m, (np.ndarray, torch.Tensor)) and list(m.shape) == list(self.resolution)
        self._mask = self.context.convert_to_tensor(m, dtype=torch.bool)

    def initial_pu(self):
        """p = 0, u = U e_x outside the solid (obstacle.py:94-99).  As in the reference the
        velocity is assembled in torch's default dtype (float32) and converted to lattice units
        before the context dtype is applied."""
        d = self.stencil.d
        p = np.zeros([1, *self.resolution], dtype=float)
        u_char = (self.units.characteristic_velocity_pu * self._unit_vector()).reshape([d] + [1] * d)
        u = (~self._mask).cpu() * u_char
        return p, u

    @property
    def grid(self):
        """node coordinates in physical units; float32 like the reference's (int64 arange divided by a
        Python float, obstacle.py:101-105)"""
        axes = [self.units.convert_length_to_pu(torch.arange(n)) for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    @property
    def post_boundaries(self):
        x = self.grid[0]
        return [EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                      velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                AntiBounceBackOutlet(self._unit_vector().tolist(), self),
                BounceBackBoundary(self.mask)]

    def _unit_vector(self, i=0):
        return torch.eye(self.stencil.d)[i]


class PoiseuilleFlow2D(ExtFlow):
    """Force-driven channel flow between two bounce-back walls (rows y = 0 and y = ny-1), periodic in x
    (lettuce/ext/_flows/poiseuille.py:18-97).  `acceleration` (physical units) drives it through a forcing
    scheme: BGKCollision(tau, force=Guo(flow, tau, flow.units.convert_acceleration_to_lu(flow.acceleration)))."""

    def __init__(self, context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None,
                 initialize_with_zeros=True):
        self.stencil = D2Q9() if stencil is None else (stencil() if callable(stencil) else stencil)
        self.initialize_with_zeros = initialize_with_zeros
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, self.stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None):
        if isinstance(resolution, int):
            return [resolution] * self.stencil.d
        assert len(resolution) == self.stencil.d
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution):
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0] - 1, characteristic_length_pu=1,
                              characteristic_velocity_pu=1)

    @property
    def grid(self):
        axes = [torch.linspace(0, 1, steps=n, device=self.context.device, dtype=self.context.dtype)
                for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    @property
    def acceleration(self):
        return self.context.convert_to_tensor([0.001, 0])

    def analytic_solution(self, t=0):
        """parabolic profile with the walls half a lattice spacing inside the boundary rows"""
        h = 0.5 / self.resolution[0]
        x, y = self.grid
        ux = self.acceleration[0] / (2 * 1 * self.units.viscosity_pu) * ((y - h) * (1 - h - y))
        u = torch.stack([ux, torch.zeros_like(ux)], dim=0)
        p = y * 0 + self.units.convert_density_lu_to_pressure_pu(1)
        return p, u

    def initial_pu(self):
        if not self.initialize_with_zeros:
            return self.analytic_solution()
        zeros = self.context.zero_tensor(self.resolution)
        return zeros[None, ...], torch.stack([zeros, zeros], dim=0)

    @property
    def post_boundaries(self):
        mask = self.context.zero_tensor(self.resolution, dtype=torch.bool)
        mask[:, [0, -1]] = True
        return [BounceBackBoundary(mask=mask)]


def _unit_box_grid(flow):
    """node coordinates on [0, 1) per axis (endpoint excluded), context dtype"""
    axes = [torch.linspace(0, 1 - 1 / n, steps=n, device=flow.context.device, dtype=flow.context.dtype)
            for n in flow.resolution]
    return torch.meshgrid(*axes, indexing="ij")


class Cavity2D(ExtFlow):
    """Lid-driven cavity: bounce-back on the left, right and bottom walls, equilibrium lid moving in +x at the
    characteristic velocity on the top row (lettuce/ext/_flows/liddrivencavity.py:14-71).  Starts at rest."""

    def __init__(self, context, resolution, reynolds_number, mach_number):
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number)

    def make_resolution(self, resolution, stencil=None):
        if isinstance(resolution, int):
            return [resolution] * 2
        assert len(resolution) == 2, "expected 2-dimensional resolution"
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution):
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0], characteristic_length_pu=1,
                              characteristic_velocity_pu=1)

    @property
    def grid(self):
        return _unit_box_grid(self)

    def initial_pu(self):
        zeros = self.context.zero_tensor(self.resolution)
        return zeros[None, ...], torch.stack([zeros, zeros])

    @property
    def post_boundaries(self):
        walls = self.context.zero_tensor(self.resolution, dtype=torch.bool)
        lid = self.context.zero_tensor(self.resolution, dtype=torch.bool)
        walls[[0, -1], 1:] = True     # left and right
        walls[:, 0] = True            # bottom
        lid[:, -1] = True             # top (wins over the side walls in the corners: later boundary)
        return [BounceBackBoundary(walls),
                EquilibriumBoundaryPU(self.context, self, lid, [float(self.units.characteristic_velocity_pu), 0.0])]


class DoublyPeriodicShear2D(ExtFlow):
    """Two tanh shear layers with a sinusoidal cross-flow perturbation on the periodic unit square
    (lettuce/ext/_flows/doublyshear.py:20-91)."""

    def __init__(self, context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None,
                 shear_layer_width=80, initial_perturbation_magnitude=0.05, initialize_fneq: bool = True):
        self.initialize_fneq = initialize_fneq
        self.initial_perturbation_magnitude = initial_perturbation_magnitude
        self.shear_layer_width = shear_layer_width
        self.stencil = D2Q9() if stencil is None else (stencil() if callable(stencil) else stencil)
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, self.stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None):
        if isinstance(resolution, int):
            return [resolution] * self.stencil.d
        assert len(resolution) == 2, "expected 2-dimensional resolution"
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution):
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0], characteristic_length_pu=1,
                              characteristic_velocity_pu=1)

    @property
    def grid(self):
        return _unit_box_grid(self)

    def initial_pu(self):
        x, y = self.grid
        w = self.shear_layer_width
        u1 = torch.where(y > 0.5, torch.tanh(w * (y - 0.25)), torch.tanh(w * (0.75 - y)))
        u2 = self.initial_perturbation_magnitude * torch.sin(2 * torch.pi * (x + 0.25))
        return torch.zeros_like(u1)[None, ...], torch.stack([u1, u2])

    @property
    def post_boundaries(self):
        return []



class CouetteFlow2D(ExtFlow):
    """Plane Couette flow, initially at rest: equilibrium "moving wall" on column y = 1, bounce-back on y = ny-1
    (lettuce/ext/_flows/couette.py:16-75)."""

    def __init__(self, context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None):
        self.u0 = 0
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        return [resolution] * 2 if isinstance(resolution, int) else list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0], characteristic_length_pu=1,
                              characteristic_velocity_pu=self.u0)

    def analytic_solution(self):
        _, y = self.grid
        return self.context.convert_to_tensor(torch.stack([y / self.resolution[0] + self.u0]))

    def initial_pu(self):
        zeros = self.context.zero_tensor(self.resolution)
        return zeros[None, ...], torch.stack([zeros, zeros], dim=0)

    @property
    def grid(self):
        axes = [torch.linspace(0, 1, steps=n, device=self.context.device, dtype=self.context.dtype)
                for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    @property
    def post_boundaries(self):
        top = torch.zeros(self.resolution, dtype=torch.bool)
        top[:, 1] = True
        bottom = torch.zeros(self.resolution, dtype=torch.bool)
        bottom[:, -1] = True
        return [EquilibriumBoundaryPU(flow=self, context=self.context, mask=top, velocity=np.array([1.0, 0.0])),
                BounceBackBoundary(bottom)]


class LambOseenVortex2D(ExtFlow):
    """Lamb-Oseen vortex convected with `velocity_init` on a periodic box (Wissocq et al. 2017;
    lettuce/ext/_flows/lamboseenvortex.py:19-134)."""

    def __init__(self, context, resolution, reynolds_number, mach_number, stencil=None, equilibrium=None,
                 initialize_fneq: bool = True, velocity_init=1, K=None, xc: int = None):
        self.initialize_fneq = initialize_fneq
        self.velocity_init = velocity_init
        if stencil is None and not isinstance(resolution, list):
            self.stencil = D2Q9()
        else:
            self.stencil = stencil() if callable(stencil) else stencil
        first = resolution if isinstance(resolution, int) else resolution[0]
        self.xc = first // 2 if xc is None else xc
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, self.stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        if isinstance(resolution, int):
            return [resolution] * self.stencil.d
        assert len(resolution) == 2, "expected 2-dimensional resolution"
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0], characteristic_length_pu=1,
                              characteristic_velocity_pu=1)

    @property
    def grid(self):
        axes = [self.units.convert_length_to_pu(torch.arange(0, n, device=self.context.device,
                                                             dtype=self.context.dtype)) for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    def initial_pu(self):
        return self.initial_lamboseenvortex()

    def initial_lamboseenvortex(self):
        units = self.units
        yc = self.resolution[1] * 0.5
        x, y = (units.convert_length_to_lu(g) for g in self.grid)
        ux0 = units.convert_velocity_to_lu(self.velocity_init)
        beta, rc, gamma, cv = 0.5, 20.0, 0.5, 1.0 / 3.0
        r2 = (x - self.xc) ** 2 + (y - yc) ** 2
        density = torch.pow(1.0 - (beta * ux0) ** 2 / (2.0 * cv) * torch.exp(1.0 - r2 / (2.0 * rc)),
                            1.0 / (gamma - 1.0))
        decay = torch.exp(-r2 / (2.0 * rc))
        ux = ux0 - beta * ux0 * (y - yc) / rc * decay
        uy = beta * ux0 * (x - self.xc) / rc * decay
        return (units.convert_density_lu_to_pressure_pu(density),
                torch.stack([units.convert_velocity_to_pu(ux), units.convert_velocity_to_pu(uy)], dim=0))

    @property
    def post_boundaries(self):
        return []


class DecayingTurbulence(ExtFlow):
    """Homogeneous isotropic turbulence with a prescribed initial spectrum E(k) ~ k^4 exp(-2 (k/k0)^2), random
    phases (`randseed`), divergence removed in spectral space; 2-D runs start from the pressure-Poisson solution
    (lettuce/ext/_flows/decayingturbulence.py:23-189).  The characteristic velocity is set from the field."""

    def __init__(self, context, resolution, reynolds_number, mach_number, k0=20, ic_energy=0.5, stencil=None,
                 equilibrium=None, initialize_pressure: bool = True, initialize_fneq: bool = True, randseed=None):
        self.initialize_pressure = initialize_pressure
        self.initialize_fneq = initialize_fneq
        self.randseed, self.k0, self.ic_energy = randseed, k0, ic_energy
        self.wavenumbers, self.spectrum = [], []
        if stencil is None:
            stencil = D2Q9() if len(resolution) == 2 else D3Q19()
        stencil = stencil() if callable(stencil) else stencil
        if stencil.d != 2:
            self.initialize_pressure = False
        ExtFlow.__init__(self, context, resolution, reynolds_number, mach_number, stencil, equilibrium)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        if isinstance(resolution, int):
            st = stencil() if callable(stencil) else stencil
            return [resolution] * st.d
        return list(resolution)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=resolution[0], characteristic_length_pu=2 * np.pi,
                              characteristic_velocity_pu=None)

    def analytic_solution(self, x, t=0):
        return

    def _generate_wavenumbers(self):
        self.dimensions = tuple(self.resolution)
        frequencies = [np.fft.fftfreq(n, d=1 / n) for n in self.dimensions]
        wavenumber = np.meshgrid(*frequencies)          # default 'xy' indexing, as in the reference (:66)
        wavenorms = np.linalg.norm(wavenumber, axis=0)
        self.wavenumbers = np.arange(int(np.max(wavenorms)))
        return wavenorms, wavenumber

    def _generate_spectrum(self):
        wavenorms, wavenumber = self._generate_wavenumbers()
        ek = wavenorms ** 4 * np.exp(-2 * (wavenorms / self.k0) ** 2)
        ek /= np.sum(ek)
        ek *= self.ic_energy
        shell = np.clip(np.ceil(wavenorms - 0.5), 0, len(self.wavenumbers)).astype(np.int64)
        self.spectrum = np.bincount(shell.ravel(), weights=ek.ravel(),
                                    minlength=len(self.wavenumbers) + 1)[:len(self.wavenumbers)]
        return ek, wavenumber

    def _generate_initial_velocity(self, ek, wavenumber):
        d = self.stencil.d
        axes = tuple(range(d))
        dx = self.units.convert_length_to_pu(1.0)
        np.random.seed(self.randseed)
        phases = np.random.random(np.array(wavenumber).shape) * 2 * np.pi + 0j
        uh = [np.fft.fftn(phases[a], axes=axes) for a in range(d)]
        for a in range(d):
            uh[a].ravel()[0] = 0
        uh = [np.sqrt(2 / d * ek / (uh[a].imag ** 2 + uh[a].real ** 2 + 1.e-15)) * uh[a] for a in range(d)]
        for a in range(d):
            uh[a].ravel()[0] = 0
        # remove the divergence with the modified wavenumber sin(k dx)/dx of 2nd-order central differences
        kmod = [np.sin(wavenumber[a] * dx) / dx for a in range(d)]
        knorm = np.linalg.norm(kmod, axis=0) + 1e-16
        divergence = sum(kmod[a] * uh[a] for a in range(d))
        uh = [uh[a] - divergence * kmod[a] / knorm ** 2 for a in range(d)]
        for a in range(d):
            uh[a].ravel()[0] = 0
        e_kin = 0.5 * sum(np.sum(uh[a].real ** 2 + uh[a].imag ** 2) for a in range(d))
        factor = np.sqrt(self.ic_energy / e_kin)
        norm = (self.resolution[0] * dx ** (1 - d) * np.sqrt(self.units.characteristic_length_pu)) if d == 3 \
            else (self.resolution[0] / dx)
        return np.asarray([(np.fft.ifftn(uh[a] * factor, axes=axes) * norm).real for a in range(d)])

    def initial_pu(self):
        ek, wavenumber = self._generate_spectrum()
        u = self._generate_initial_velocity(ek, wavenumber)
        self.units.characteristic_velocity_pu = np.linalg.norm(u, axis=0).max()
        return np.zeros(self.dimensions)[None, ...], u

    @property
    def energy_spectrum(self):
        return self.spectrum, self.wavenumbers

    @property
    def grid(self):
        axes = [torch.linspace(0, 2 * torch.pi * (1 - 1 / n), steps=n, device=self.context.device,
                               dtype=self.context.dtype) for n in self.resolution]
        return torch.meshgrid(*axes, indexing="ij")

    @property
    def post_boundaries(self):
        return []


def _flow_table():
    from .._stencil import D3Q27
    return {"taylor2d": (TaylorGreenVortex, D2Q9), "taylor3d_d3q19": (TaylorGreenVortex, D3Q19),
            "taylor3d_d3q27": (TaylorGreenVortex, D3Q27), "poiseuille2d": (PoiseuilleFlow2D, D2Q9),
            "shear2d": (DoublyPeriodicShear2D, D2Q9), "couette2d": (CouetteFlow2D, D2Q9),
            "decay2d": (DecayingTurbulence, D2Q9), "lamboseen": (LambOseenVortex2D, D2Q9)}


# name -> (flow class, stencil class), the registry behind `lettuce benchmark -f` (lettuce/ext/_flows/_flow_by_name.py)
flow_by_name = _flow_table()
