"""Link-wise bounce-back boundaries applied after streaming, momentum-exchange forces and the cylinder flow of the
reference's example project `examples/advanced_projects/efficient_bounce_back_obstacle` (`ebb/` below):

    SolidBoundaryData                       ebb/boundary/solid_boundary_data.py
    FullwayBounceBackBoundary               ebb/boundary/fullway_bounce_back_boundary.py
    HalfwayBounceBackBoundary               ebb/boundary/halfway_bounce_back_boundary.py
    LinearInterpolatedBounceBackBoundary    ebb/boundary/linear_interpolated_bounce_back_boundary.py
    EbbSimulation                           ebb/simulation/ebb_simulation.py
    ObstacleCylinder                        ebb/flow/obstacle_cylinder.py
    DragCoefficient, LiftCoefficient        ebb/reporter/observables_force_coefficients.py

The reference finds the solid/fluid links with Python loops over every solid node and direction and applies the
boundaries with advanced indexing on the full tensors; here the search is a handful of vectorised NumPy operations
(once, at construction) and each boundary is two sparse CUDA kernels over its link list (`lbm_apply_links`:
gather all right-hand sides incl. the force partials, then scatter), after the ordinary fused collide+stream step.
"""
from __future__ import annotations

import os
from timeit import default_timer as timer
from typing import List, Optional

import numpy as np
import torch

from .. import native
from .._flow import Boundary
from .._simulation import Simulation
from ..units import UnitConversion
from .boundary import EquilibriumBoundaryPU, EquilibriumOutletP
from .flows import ExtFlow
from .reporter import Observable

__all__ = ["SolidBoundaryData", "FullwayBounceBackBoundary", "HalfwayBounceBackBoundary",
           "LinearInterpolatedBounceBackBoundary", "EbbSimulation", "ObstacleCylinder", "DragCoefficient",
           "LiftCoefficient", "solid_fluid_links", "cylinder_solid_boundary_data"]

LINK_FULLWAY, LINK_HALFWAY, LINK_INTERPOLATED = 0, 1, 2


class SolidBoundaryData(dict):
    """index lists of a solid boundary: populations `[q, x, y(, z)]` on the fluid side pointing into the solid, split
    by wall distance d <= 0.5 (`lt`) and d > 0.5 (`gt`), the distances, and the solid mask
    (ebb/boundary/solid_boundary_data.py:8-14)"""
    f_index_lt: np.ndarray
    f_index_gt: np.ndarray
    d_lt: np.ndarray
    d_gt: np.ndarray
    points_inside: np.ndarray
    solid_mask: np.ndarray
    not_intersected: np.ndarray = np.ndarray([])


def _as_numpy_mask(mask) -> np.ndarray:
    if torch.is_tensor(mask):
        mask = mask.detach().cpu().numpy()
    return np.asarray(mask, dtype=bool)


def solid_fluid_links(stencil, mask, periodicity=None, other_solid=None):
    """(q_in [n], solid nodes [n, d], fluid nodes [n, d]): every pair of a solid node of `mask` and a neighbouring
    node outside `mask` (and outside `other_solid`); q_in points from that neighbour into the solid node.  Order and
    border behaviour of the reference's loops (ebb/boundary/fullway_bounce_back_boundary.py:50-123,
    halfway_bounce_back_boundary.py:65-152): solid nodes in C order, stencil direction ascending; on a non-periodic
    axis a neighbour index of -1 wraps (negative indexing) while an index of n is skipped (IndexError)."""
    mask = _as_numpy_mask(mask)
    res, d = mask.shape, mask.ndim
    e = np.asarray(stencil.e)[:, :d]
    opposite = np.asarray(stencil.opposite)
    periodicity = tuple(bool(p) for p in periodicity[:d]) if periodicity is not None else (False,) * d
    blocked = mask if other_solid is None else (mask | _as_numpy_mask(other_solid))
    solid = np.argwhere(mask)
    flat = np.ravel_multi_index(tuple(solid.T), res) if len(solid) else np.zeros(0, dtype=np.int64)
    q_in, s_nodes, f_nodes, keys = [], [], [], []
    for i in range(len(e)):
        nb = solid + e[i][None, :]
        ok = np.ones(len(solid), dtype=bool)
        for a in range(d):
            if periodicity[a]:
                nb[:, a] %= res[a]
            else:
                ok &= nb[:, a] < res[a]
                nb[:, a] = np.where(nb[:, a] < 0, nb[:, a] + res[a], nb[:, a])
        ok &= ~blocked[tuple(np.where(ok[:, None], nb, 0).T)]
        q_in.append(np.full(int(ok.sum()), opposite[i], dtype=np.int64))
        s_nodes.append(solid[ok])
        f_nodes.append(nb[ok])
        keys.append(flat[ok] * len(e) + i)
    order = np.argsort(np.concatenate(keys), kind="stable")
    return np.concatenate(q_in)[order], np.concatenate(s_nodes)[order], np.concatenate(f_nodes)[order]


def cylinder_solid_boundary_data(stencil, obstacle_mask, x_center: float, y_center: float, radius: float
                                 ) -> SolidBoundaryData:
    """Links and wall distances of a circular cylinder (axis along z): `ObstacleCylinder.make_ibb_index_lists`
    (ebb/flow/obstacle_cylinder.py:283-466).  The search treats every axis as periodic; d solves
    |p + d c - centre| = radius along the link (the first root that is <= 1)."""
    mask = _as_numpy_mask(obstacle_mask)
    q_in, _, fluid = solid_fluid_links(stencil, mask, (True,) * mask.ndim)
    c = np.asarray(stencil.e)[q_in][:, :2].astype(float)
    px, py = fluid[:, 0].astype(float), fluid[:, 1].astype(float)
    cc = c[:, 0] ** 2 + c[:, 1] ** 2
    h1 = (px * c[:, 0] + py * c[:, 1] - c[:, 0] * x_center - c[:, 1] * y_center) / cc
    h2 = (px * px + py * py + x_center * x_center + y_center * y_center - 2 * px * x_center - 2 * py * y_center
          - radius * radius) / cc
    with np.errstate(invalid="ignore"):
        root = np.sqrt(h1 * h1 - h2)
    d1, d2 = -h1 + root, -h1 - root
    dist = np.where(d1 <= 1, d1, d2)
    valid = dist <= 1                                  # the reference prints a warning and drops the link otherwise
    index = np.concatenate([q_in[:, None], fluid], axis=1)
    lt, gt = valid & (dist <= 0.5), valid & (dist > 0.5)
    sbd = SolidBoundaryData()
    sbd.solid_mask = mask
    sbd.f_index_lt, sbd.f_index_gt = index[lt], index[gt]
    sbd.d_lt, sbd.d_gt = dist[lt], dist[gt]
    return sbd


class _LinkBoundary(Boundary):
    """shared part: the link list (population index + node coordinates), its device copy for `lbm_apply_links`,
    the force of the last application"""
    link_kind = LINK_FULLWAY
    freezes_solid_nodes = True

    def _setup(self, context, flow, mask, index: np.ndarray, distance: Optional[np.ndarray], calc_force: bool):
        self.context, self.flow = context, flow
        self.mask = mask
        d = flow.stencil.d
        index = np.asarray(index, dtype=np.int64).reshape(-1, d + 1)
        self._index = index
        self._distance = None if distance is None else np.asarray(distance, dtype=np.float64).reshape(-1)
        self.calc_force = bool(calc_force)
        self._device = None
        self._force = None
        if calc_force:
            self._force = torch.zeros(3, dtype=torch.float64, device=context.device)

    @property
    def n_links(self) -> int:
        return int(self._index.shape[0])

    @property
    def force_sum(self) -> torch.Tensor:
        """momentum-exchange force [d] on the boundary in lattice units, as of the last time step"""
        if self._force is None:
            raise AttributeError("construct the boundary with calc_force=True")
        return self._force[:self.flow.stencil.d]

    def make_no_collision_mask(self, shape: List[int], context) -> Optional[torch.Tensor]:
        return context.convert_to_tensor(_as_numpy_mask(self.mask), dtype=torch.bool)

    def make_no_streaming_mask(self, shape: List[int], context) -> Optional[torch.Tensor]:
        if not self.freezes_solid_nodes:
            return None          # "FWBB needs streaming to invert the populations on the solid itself"
        return context.convert_to_tensor(_as_numpy_mask(self.mask), dtype=torch.bool)

    def native_available(self) -> bool:
        return True

    def link_descriptor(self, f: torch.Tensor) -> "native.LbmLinks":
        """ctypes descriptor over device copies of the link list (built on first use)"""
        if self._device is None or self._device["device"] != f.device or self._device["dtype"] != f.dtype:
            res = [int(s) for s in f.shape[1:]]
            n = self.n_links
            flat = np.ravel_multi_index(tuple(self._index[:, 1:].T), res) if n else np.zeros(0, dtype=np.int64)
            dev = dict(device=f.device, dtype=f.dtype,
                       node=torch.as_tensor(flat, dtype=torch.int32).to(f.device),
                       q=torch.as_tensor(self._index[:, 0], dtype=torch.uint8).to(f.device),
                       d=None if self._distance is None else torch.as_tensor(self._distance).to(f.device, f.dtype),
                       bounced=torch.empty(max(n, 1), dtype=f.dtype, device=f.device),
                       scratch=torch.empty(max(int(native.lib().lbm_links_scratch_doubles(n)), 1),
                                           dtype=torch.float64, device=f.device))
            if self._force is not None:
                self._force = self._force.to(f.device)
            links = native.LbmLinks()
            links.kind = self.link_kind
            links.n = n
            links.node, links.q = dev["node"].data_ptr(), dev["q"].data_ptr()
            links.d = dev["d"].data_ptr() if dev["d"] is not None else None
            links.bounced = dev["bounced"].data_ptr()
            links.force_scratch = dev["scratch"].data_ptr() if self._force is not None else None
            links.force = self._force.data_ptr() if self._force is not None else None
            dev["links"] = links
            self._device = dev
        return self._device["links"]


class FullwayBounceBackBoundary(_LinkBoundary):
    """Full-way bounce-back applied after streaming: on the solid nodes next to the fluid, the populations that just
    streamed in are reversed (ebb/boundary/fullway_bounce_back_boundary.py:9-185); force = 2 sum e_q f_q over them."""
    link_kind = LINK_FULLWAY
    freezes_solid_nodes = False

    def __init__(self, context, flow, mask, global_solid_mask=None, periodicity=None, calc_force: bool = False):
        m = _as_numpy_mask(mask)
        other = None if global_solid_mask is None else np.where(~m, _as_numpy_mask(global_solid_mask), False)
        q_in, solid, _ = solid_fluid_links(flow.stencil, m, periodicity, other)
        self._setup(context, flow, mask, np.concatenate([q_in[:, None], solid], axis=1), None, calc_force)
        self.f_index_fwbb = torch.as_tensor(self._index, device=context.device)


class HalfwayBounceBackBoundary(_LinkBoundary):
    """Half-way bounce-back: on the fluid nodes next to the solid, the population that left towards the solid
    comes back reversed within the same step (ebb/boundary/halfway_bounce_back_boundary.py:10-254);
    force = 2 sum e_q fc_q."""
    link_kind = LINK_HALFWAY

    def __init__(self, context, flow, solid_boundary_data: SolidBoundaryData, global_solid_mask=None,
                 periodicity=None, calc_force: bool = False):
        mask = solid_boundary_data.solid_mask
        lt = getattr(solid_boundary_data, "f_index_lt", None)
        gt = getattr(solid_boundary_data, "f_index_gt", None)
        if lt is not None or gt is not None:        # unified solid boundary data: lt first, then gt (:48-56)
            parts = [np.asarray(p).reshape(-1, flow.stencil.d + 1) for p in (lt, gt) if p is not None and len(p)]
            index = np.concatenate(parts, axis=0) if parts else np.zeros((0, flow.stencil.d + 1), dtype=np.int64)
        else:                                       # legacy neighbour search on the mask (:60-152)
            m = _as_numpy_mask(mask)
            other = m if global_solid_mask is None else _as_numpy_mask(global_solid_mask)
            q_in, _, fluid = solid_fluid_links(flow.stencil, m, periodicity, other)
            index = np.concatenate([q_in[:, None], fluid], axis=1)
        self._setup(context, flow, mask, index, None, calc_force)
        self.f_index = torch.as_tensor(self._index, device=context.device)


class LinearInterpolatedBounceBackBoundary(_LinkBoundary):
    """Bouzidi's linearly interpolated bounce-back (IBB1) with the wall distances of a `SolidBoundaryData`
    (ebb/boundary/linear_interpolated_bounce_back_boundary.py:9-212); force = sum e_q (fc_q + bounced)."""
    link_kind = LINK_INTERPOLATED

    def __init__(self, context, flow, solid_boundary_data: SolidBoundaryData, calc_force: bool = False):
        w = flow.stencil.d + 1
        lt = np.asarray(solid_boundary_data.f_index_lt).reshape(-1, w)
        gt = np.asarray(solid_boundary_data.f_index_gt).reshape(-1, w)
        d_lt = np.asarray(solid_boundary_data.d_lt, dtype=np.float64).reshape(-1)
        d_gt = np.asarray(solid_boundary_data.d_gt, dtype=np.float64).reshape(-1)
        if len(d_lt) != len(lt) or len(d_gt) != len(gt):
            raise ValueError("SolidBoundaryData: one wall distance per link is required")
        if (len(d_lt) and d_lt.max() > 0.5) or (len(d_gt) and (d_gt.min() <= 0.5 or d_gt.max() > 1)):
            raise ValueError("SolidBoundaryData: d_lt must lie in (0, 0.5], d_gt in (0.5, 1]")
        self._setup(context, flow, solid_boundary_data.solid_mask, np.concatenate([lt, gt], axis=0),
                    np.concatenate([d_lt, d_gt]), calc_force)
        self.f_index_lt = torch.as_tensor(lt, dtype=torch.int64, device=context.device)
        self.f_index_gt = torch.as_tensor(gt, dtype=torch.int64, device=context.device)
        self.d_lt, self.d_gt = context.convert_to_tensor(d_lt), context.convert_to_tensor(d_gt)


class EbbSimulation(Simulation):
    """Simulation with boundaries applied AFTER streaming (substeps collide, stream, boundary, report):
    `flow.post_streaming_boundaries` continue the label numbering of the no-collision mask and may freeze their
    solid nodes in the no-streaming mask (ebb/simulation/ebb_simulation.py:11-106).  Each time step is the fused
    collide+stream kernel followed by two sparse kernels per post-streaming boundary."""

    def __init__(self, flow, collision, reporter):
        super().__init__(flow, collision, reporter)
        psb = getattr(flow, "post_streaming_boundaries", None)
        self.post_streaming_boundaries = list(psb) if psb is not None else []
        if self.post_streaming_boundaries:
            shape = [int(s) for s in flow.f.shape]
            ctx = self.context
            if self.no_collision_mask is None:
                self.no_collision_mask = ctx.full_tensor(shape[1:], self.collision_index, dtype=torch.uint8)
            if self.no_streaming_mask is None:
                self.no_streaming_mask = ctx.full_tensor(shape, self.collision_index, dtype=torch.uint8)
            first = self.collision_index + 1 + len(self.post_boundaries)
            for label, boundary in enumerate(self.post_streaming_boundaries, start=first):
                if not isinstance(boundary, _LinkBoundary):
                    raise NotImplementedError(f"{type(boundary).__name__} has no B200 post-streaming kernel")
                if label > 127:
                    raise ValueError("more than 127 transformer labels")
                ncm = boundary.make_no_collision_mask(shape[1:], context=ctx)
                if ncm is not None:
                    self.no_collision_mask[ncm.to(device=ctx.device, dtype=torch.bool)] = label
                nsm = boundary.make_no_streaming_mask(shape, context=ctx)
                if nsm is not None:
                    self.no_streaming_mask |= nsm.to(device=ctx.device, dtype=torch.uint8)

    def _post_streaming_boundaries(self):
        engine = native.engine_of(self)
        for boundary in self.post_streaming_boundaries:
            engine.apply_links(boundary)

    def __call__(self, num_steps: int) -> float:
        self.context.synchronize()
        beg = timer()
        if self.flow.i == 0:
            self._report()
        batched = os.environ.get("LBM_B200_EBB_BATCH", "0") == "1" and self._collide_and_stream is native.invoke
        remaining = int(num_steps)
        while remaining > 0:
            # opt-in (not yet measured on hardware): the steps up to the next due reporter in one library call
            k = self._batch_length(remaining) if batched else 1
            if batched:
                native.engine_of(self).step_with_links(self.post_streaming_boundaries, k)
            else:
                self._collide_and_stream(self)
                self._post_streaming_boundaries()
            self.flow.i += k
            self._report()
            remaining -= k
        self._flush_reporters()
        self.context.synchronize()
        end = timer()
        return num_steps * int(np.prod(self.flow.resolution)) / 1e6 / (end - beg)


class ObstacleCylinder(ExtFlow):
    """Flow around a circular cylinder in 2-D or 3-D (axis along z): equilibrium inlet at x = 0,
    EquilibriumOutletP at x = nx-1, lateral y-boundaries periodic or full-way bounce-back walls, cylinder boundary
    `bc_type` in {fwbb, hwbb, ibb1} applied after streaming (ebb/flow/obstacle_cylinder.py:15-596)."""

    def __init__(self, context, resolution, reynolds_number, mach_number, char_length_pu, char_length_lu,
                 char_velocity_pu=1, lateral_walls="periodic", bc_type="fwbb", perturb_init=True, u_init=0,
                 x_offset=0, y_offset=0, calc_force_coefficients=False, stencil=None, equilibrium=None):
        self.char_length_pu, self.char_length_lu, self.char_velocity_pu = char_length_pu, char_length_lu, char_velocity_pu
        self.resolution = self.make_resolution(resolution, stencil)
        self.perturb_init, self.u_init = perturb_init, u_init
        self.lateral_walls, self.bc_type = lateral_walls, bc_type
        self.calc_force_coefficients = calc_force_coefficients
        d = len(self.resolution)
        self.periodicity = (False, False, True if d == 3 else None)
        self.x_offset, self.y_offset = x_offset, y_offset
        self.radius_lu = char_length_lu / 2
        self.y_pos_lu = self.resolution[1] / 2 + 0.5 + y_offset        # one-based node coordinates
        self.x_pos_lu = self.y_pos_lu + x_offset
        # As in the reference the masks are still empty while the base class evaluates the initial condition
        # (obstacle_cylinder.py:74-87): the initial velocity is not zeroed inside the cylinder and the walls.  A
        # later flow.initialize() sees the filled masks.
        self.solid_mask = np.zeros(self.resolution, dtype=bool)
        self.wall_mask = np.zeros_like(self.solid_mask)
        self._obstacle_mask = np.zeros_like(self.solid_mask)
        ExtFlow.__init__(self, context, self.resolution, reynolds_number, mach_number, stencil, equilibrium)
        axes = [np.linspace(1, n, n) for n in self.resolution]
        grid = np.meshgrid(*axes, indexing="ij")
        inside = np.sqrt((grid[0] - self.x_pos_lu) ** 2 + (grid[1] - self.y_pos_lu) ** 2) < self.radius_lu
        self._obstacle_mask[inside] = True
        self.solid_mask[inside] = True
        self.in_mask = np.zeros(self.resolution, dtype=bool)
        if lateral_walls in ("bounceback", "slip"):
            self.wall_mask[:, [0, -1]] = True
            self.solid_mask[self.wall_mask] = True
            self.in_mask[0, 1:-1] = True
        else:
            self.in_mask[0, :] = True
        self.u_inlet = self.units.characteristic_velocity_pu * self._unit_vector()
        if lateral_walls == "bounceback":
            ny = self.resolution[1]
            y = np.linspace(0, ny, ny)
            parabola = np.zeros((1, ny))
            parabola[:, 1:-1] = -1.5 * np.array(self.u_inlet).max() * y[1:-1] * (y[1:-1] - ny) / (ny / 2) ** 2
            if d == 2:
                self.u_inlet = np.stack([parabola, np.zeros_like(parabola)], axis=0)
            else:
                plane = parabola[:, :, None] * np.ones(self.resolution[2])
                self.u_inlet = np.stack([plane, np.zeros_like(plane), np.zeros_like(plane)], axis=0)

    def make_units(self, reynolds_number, mach_number, resolution) -> UnitConversion:
        return UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                              characteristic_length_lu=self.char_length_lu,
                              characteristic_length_pu=self.char_length_pu,
                              characteristic_velocity_pu=self.char_velocity_pu)

    def make_resolution(self, resolution, stencil=None) -> List[int]:
        if isinstance(resolution, int):
            st = stencil() if callable(stencil) else stencil
            return [resolution] * st.d
        return list(resolution)

    @property
    def obstacle_mask(self):
        return self._obstacle_mask

    @obstacle_mask.setter
    def obstacle_mask(self, m):
        assert isinstance(m, np.ndarray) and list(m.shape) == list(self.resolution)
        self._obstacle_mask = m.astype(bool)

    def _unit_vector(self, i=0):
        return np.eye(len(self.resolution))[i]

    @property
    def grid(self):
        axes = [self.units.convert_length_to_pu(np.linspace(0, n, n)) for n in self.resolution]
        return np.meshgrid(*axes, indexing="ij")

    def initial_pu(self):
        """p = 0; u = 0 (`u_init` 0) or U e_x outside the solid, parabolic between bounce-back walls; a sine
        perturbation of the second column breaks the symmetry (obstacle_cylinder.py:193-276)"""
        d = len(self.resolution)
        ny = self.resolution[1]
        p = np.zeros([1, *self.resolution], dtype=float)
        u_char = self.units.characteristic_velocity_pu
        u = (1 - self.solid_mask) * (u_char * self._unit_vector()).reshape([d] + [1] * d)
        if self.u_init == 0:
            u = u * 0
        elif self.lateral_walls == "bounceback":
            y = np.linspace(0, ny, ny)
            factor = np.zeros(ny)
            factor[1:-1] = -y[1:-1] * (y[1:-1] - ny) / (ny / 2) ** 2
            u = np.einsum("k,ijk->ijk", factor, u) if d == 2 else np.einsum("k,ijkl->ijkl", factor, u)
        if self.perturb_init:
            wave_y = np.sin(np.linspace(0, ny, ny) / ny * 2 * np.pi)
            if u.max() < 0.5 * u_char:
                if d == 2:
                    u[0][1] += wave_y * u_char * 0.1
                else:
                    nz = self.resolution[2]
                    plane = np.ones_like(u[0, 1])
                    u[0][1] = np.einsum("y,yz->yz", wave_y * u_char * 0.1, plane)
                    u[0][1] += np.einsum("z,yz->yz", np.sin(np.linspace(0, nz, nz) / nz * 2 * np.pi) * u_char * 0.1,
                                         plane)
            else:
                if d == 2:
                    u[0][1] *= 1 + wave_y * 0.1
                else:
                    # like the reference this needs ny == nz: its z factor is built with ny points (:279-284)
                    u[0][1] = np.einsum("y,yz->yz", 1 + wave_y * 0.1, u[0][1])
                    u[0][1] = np.einsum("z,yz->yz", 1 + wave_y * 0.1, u[0][1])
        return p, u

    def make_ibb_index_lists(self, x_center, y_center, radius):
        sbd = cylinder_solid_boundary_data(self.stencil, self.obstacle_mask, x_center, y_center, radius)
        return [sbd.f_index_lt, sbd.f_index_gt, sbd.d_lt, sbd.d_gt]

    @property
    def post_boundaries(self):
        direction = [1] + [0] * (len(self.resolution) - 1)
        return [EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.as_tensor(self.in_mask),
                                      velocity=self.u_inlet),
                EquilibriumOutletP(direction=direction, flow=self)]

    @property
    def post_streaming_boundaries(self):
        boundaries = []
        if self.lateral_walls in ("bounceback", "slip"):       # the reference falls back to full-way walls for both
            boundaries.append(FullwayBounceBackBoundary(self.context, self, self.wall_mask,
                                                        periodicity=self.periodicity))
        sbd = cylinder_solid_boundary_data(self.stencil, self.obstacle_mask, self.x_pos_lu - 1, self.y_pos_lu - 1,
                                           self.radius_lu)
        kind = str(self.bc_type).casefold()
        if kind == "hwbb":
            obstacle = HalfwayBounceBackBoundary(self.context, self, sbd, periodicity=self.periodicity,
                                                 calc_force=self.calc_force_coefficients)
        elif kind == "ibb1":
            obstacle = LinearInterpolatedBounceBackBoundary(self.context, self, sbd,
                                                            calc_force=self.calc_force_coefficients)
        else:
            obstacle = FullwayBounceBackBoundary(self.context, self, self.obstacle_mask, periodicity=self.periodicity,
                                                 calc_force=self.calc_force_coefficients)
        boundaries.append(obstacle)       # the obstacle comes last so that its force is that of the final state
        return boundaries


class _ForceCoefficient(Observable):
    """force component / (0.5 rho_mean U_lu^2 A_lu), rho_mean over the nodes outside `solid_mask`
    (ebb/reporter/observables_force_coefficients.py:16-98)"""
    component = 0

    def __init__(self, flow, obstacle_boundary, solid_mask, area_pu: float):
        super().__init__(flow)
        self.obstacle_boundary = obstacle_boundary
        units = flow.units
        self.area_lu = area_pu * (units.characteristic_length_lu / units.characteristic_length_pu) ** (flow.stencil.d - 1)
        self.solid_mask = self.context.convert_to_tensor(_as_numpy_mask(solid_mask), dtype=torch.bool)

    def __call__(self, f=None):
        rho = self.flow.rho(f)[0]
        rho_mean = rho[~self.solid_mask].mean()
        force = self.obstacle_boundary.force_sum[self.component]
        return force / (0.5 * rho_mean * self.flow.units.characteristic_velocity_lu ** 2 * self.area_lu)


class DragCoefficient(_ForceCoefficient):
    component = 0


class LiftCoefficient(_ForceCoefficient):
    component = 1
