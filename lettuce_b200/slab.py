"""Multi-GPU x-slab decomposition, one process per GPU.

The reference has no multi-device code at all; this module is the B200-native
addition.  The global lattice `[nx, ny(, nz)]` is cut along x -- the slowest axis, so every plane
`f[q, x, :, :]` stays contiguous -- into one slab per rank.  Streaming across a cut is not a separate
exchange: every rank maps its two neighbours' population buffers with CUDA IPC and the step kernel
loads (pull) or stores (push) the neighbour's boundary plane directly over NVLink
(csrc/lbm_step.cuh `in_plane`/`out_plane`, csrc/lbm_slab.cu).  Ranks are kept in lock step by one
8-byte progress counter per neighbour, written through the same peer mapping after every step.
`torch.distributed` is only the control plane (handle exchange, reporter all-reduce).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import native
from ._simulation import Simulation, StreamingStrategy
from .ext.flows import Obstacle, TaylorGreenVortex
from .ext.reporter import Observable

__all__ = ["SlabDecomposition", "SlabTaylorGreenVortex", "SlabObstacle", "SlabSimulation", "SlabEngine",
           "GlobalSum", "GlobalMax", "SlabEnstrophy", "make_tgv_slab_simulation", "gather_populations",
           "dump_slabs", "load_slabs"]


class SlabDecomposition:
    """Contiguous x-ranges per rank; the first `nx % world` ranks get one extra plane."""

    def __init__(self, nx_global: int, world: int, rank: int):
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"bad rank {rank} of {world}")
        if nx_global < world:
            raise ValueError(f"{nx_global} planes cannot be split over {world} ranks")
        base, rem = divmod(int(nx_global), int(world))
        self.sizes = [base + (1 if r < rem else 0) for r in range(world)]
        self.offsets = [sum(self.sizes[:r]) for r in range(world)]
        self.nx_global, self.world, self.rank = int(nx_global), int(world), int(rank)
        self.x0 = self.offsets[rank]
        self.nx_local = self.sizes[rank]
        self.x1 = self.x0 + self.nx_local
        self.lo = (rank - 1) % world          # owner of global plane x0 - 1 (periodic ring: the reference
        self.hi = (rank + 1) % world          # streams with torch.roll even for inlet/outlet flows)

    def owner_of(self, x: int) -> int:
        x %= self.nx_global
        for r in range(self.world):
            if self.offsets[r] <= x < self.offsets[r] + self.sizes[r]:
                return r
        raise AssertionError

    def local_slice(self) -> slice:
        return slice(self.x0, self.x1)

    def halo_indices(self, width: int) -> List[int]:
        """global x indices of the slab extended by `width` planes on both sides (periodic)"""
        return [(x % self.nx_global) for x in range(self.x0 - width, self.x1 + width)]


class SlabTaylorGreenVortex(TaylorGreenVortex):
    """The rank-local slab of a global Taylor-Green vortex.  `resolution` is the LOCAL slab,
    `global_resolution` the whole lattice; units come from the global lattice.  The initial state is
    evaluated on the slab extended by three planes (the radius of the 6th-order stencil in
    `initialize_f_neq`) and cropped, so it equals the corresponding slice of the global initial state."""
    _HALO = 3

    def __init__(self, context, global_resolution, reynolds_number, mach_number, stencil,
                 decomposition: SlabDecomposition, equilibrium=None, initialize_fneq: bool = True):
        self.decomposition = decomposition
        self.global_resolution = [int(r) for r in global_resolution]
        assert decomposition.nx_global == self.global_resolution[0]
        self._extended = False
        local = [decomposition.nx_local] + self.global_resolution[1:]
        TaylorGreenVortex.__init__(self, context, local, reynolds_number, mach_number, stencil, equilibrium,
                                   initialize_fneq)

    def make_units(self, reynolds_number, mach_number, resolution):
        return TaylorGreenVortex.make_units(self, reynolds_number, mach_number, self.global_resolution)

    @property
    def grid(self):
        g = self.global_resolution
        axes = [torch.linspace(0, 2 * torch.pi * (1 - 1 / n), steps=n, device=self.context.device,
                               dtype=self.context.dtype) for n in g]
        dec = self.decomposition
        idx = dec.halo_indices(self._HALO if self._extended else 0)
        axes[0] = axes[0][torch.as_tensor(idx, device=self.context.device)]
        return torch.meshgrid(*axes, indexing="ij")

    def initialize(self):
        h = self._HALO
        local = list(self.resolution)
        self._extended = True
        self.resolution = [local[0] + 2 * h] + local[1:]
        try:
            TaylorGreenVortex.initialize(self)
            self.f = self.f[:, h:-h].contiguous()
        finally:
            self._extended = False
            self.resolution = local
        self._f_next = None


class SlabObstacle(Obstacle):
    """The rank-local slab of a global `Obstacle` flow.  `grid` holds the GLOBAL physical coordinates of
    the local nodes, so masks written in terms of coordinates (inlet `x == 0`, solids) come out right on
    every rank; `mask` has the local slab's shape.  Subclass and override `post_boundaries` exactly as
    with `Obstacle`."""

    def __init__(self, context, global_resolution, reynolds_number, mach_number, domain_length_x,
                 decomposition: SlabDecomposition, char_length=1, char_velocity=1, stencil=None, equilibrium=None):
        self.decomposition = decomposition
        self.global_resolution = [int(r) for r in global_resolution]
        assert decomposition.nx_global == self.global_resolution[0]
        local = [decomposition.nx_local] + self.global_resolution[1:]
        self._global_char_length_lu = self.global_resolution[0] / domain_length_x * char_length
        Obstacle.__init__(self, context, local, reynolds_number, mach_number,
                          domain_length_x * decomposition.nx_local / self.global_resolution[0],
                          char_length, char_velocity, stencil, equilibrium)

    @property
    def grid(self):
        dec = self.decomposition
        starts = [dec.x0] + [0] * (len(self.resolution) - 1)
        axes = [self.units.convert_length_to_pu(torch.arange(s, s + n)) for s, n in zip(starts, self.resolution)]
        return torch.meshgrid(*axes, indexing="ij")

    @property
    def global_extent_pu(self):
        """largest physical coordinate per axis of the GLOBAL lattice (what `grid[i].max()` is on one GPU)"""
        return [self.units.convert_length_to_pu(torch.arange(n - 1, n))[0] for n in self.global_resolution]


class _RawCuda:
    """exposes a raw device allocation to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class _IpcBuffer:
    def __init__(self, nbytes: int):
        L = native.lib()
        self.ptr = C.c_void_p()
        self.handle = (C.c_ubyte * 64)()
        native.check(L.lbm_ipc_alloc(C.c_size_t(nbytes), C.byref(self.ptr), self.handle), "lbm_ipc_alloc")
        self.nbytes = nbytes
        self.freed = False

    def tensor(self, shape, dtype, device) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int64: "<i8"}[dtype]
        t = torch.as_tensor(_RawCuda(self.ptr.value, shape, typestr), device=device)
        assert t.data_ptr() == self.ptr.value
        return t

    def free(self):
        if not self.freed:
            native.lib().lbm_ipc_free(self.ptr)
            self.freed = True


class SlabEngine(native.Engine):
    """Engine of one rank's slab: IPC population buffers, neighbour mappings, lock-stepped stepping."""

    def __init__(self, simulation, decomposition: SlabDecomposition, group=None):
        super().__init__(simulation)
        self.dec = decomposition
        self.group = group
        flow = self.flow
        L = self.lib
        with torch.cuda.device(self.device):
            nbytes = flow.f.numel() * flow.f.element_size()
            self.buf = [_IpcBuffer(nbytes), _IpcBuffer(nbytes)]
            self.flags = _IpcBuffer(256)
            shape = tuple(flow.f.shape)
            self.t = [b.tensor(shape, flow.f.dtype, self.device) for b in self.buf]
            self.t[0].copy_(flow.f)
            torch.cuda.synchronize(self.device)
        flow.f, flow.f_next = self.t[0], self.t[1]
        self.cur = 0
        self.epoch = 0
        # control plane: ship (handles, slab thickness) to everybody, map the two neighbours
        mine = (bytes(self.buf[0].handle), bytes(self.buf[1].handle), bytes(self.flags.handle), self.dec.nx_local)
        if self.dec.world > 1:
            everyone = [None] * self.dec.world
            dist.all_gather_object(everyone, mine, group=group)
        else:
            everyone = [mine]
        self._peers = {}
        self.lo = self._map_peer(self.dec.lo, everyone)
        self.hi = self._map_peer(self.dec.hi, everyone)
        if self.labels is not None and self.dec.world > 1:
            self._stitch_masks(simulation)
        if self.dec.world > 1:
            dist.barrier(group=group)       # every rank has filled its buffer and mapped its neighbours

    def _stitch_masks(self, simulation):
        """Make the packed masks consistent across the cuts.  `lbm_pack_masks` treated the slab as periodic
        in x; here the two cut planes get (a) the neighbours' frozen-slot words, which the general scatter
        consults before storing into the neighbour's plane, and (b) the "general" bit wherever a population
        leaving through the cut would land in a frozen slot of the neighbour."""
        dec, dev = self.dec, self.device
        st = self.flow.stencil
        res = [int(r) for r in self.flow.f.shape[1:]]
        plane_shape = res[1:]
        fro = self.frozen.view(res)
        mine = torch.stack([fro[0], fro[-1]]).contiguous()                    # [2, *plane]
        planes = [torch.empty_like(mine) for _ in range(dec.world)]
        dist.all_gather(planes, mine, group=self.group)
        self.frozen_lo = planes[dec.lo][1].contiguous()                       # lo neighbour's LAST plane
        self.frozen_hi = planes[dec.hi][0].contiguous()                       # hi neighbour's FIRST plane
        self.desc.halo.frozen_lo = self.frozen_lo.data_ptr()
        self.desc.halo.frozen_hi = self.frozen_hi.data_ptr()
        lab = self.labels.view(res)
        dims = tuple(range(len(plane_shape)))
        for side, nb_frozen, ex in ((0, self.frozen_lo, -1), (-1, self.frozen_hi, 1)):
            hit = torch.zeros(plane_shape, dtype=torch.bool, device=dev)
            for q, e in enumerate(st.e):
                if e[0] != ex:
                    continue
                # node (y,z) streams population q into (y+e_y, z+e_z) of the neighbour plane
                dest_frozen = ((nb_frozen >> q) & 1).bool()
                hit |= torch.roll(dest_frozen, shifts=tuple(-int(c) for c in e[1:]), dims=dims)
            lab[side] |= (hit.to(torch.uint8) * 128)
        self._list_general_nodes()

    def _map_peer(self, rank, everyone):
        if rank in self._peers:
            return self._peers[rank]
        if rank == self.dec.rank:
            ptrs = (self.buf[0].ptr.value, self.buf[1].ptr.value, self.flags.ptr.value)
        else:
            ptrs = []
            with torch.cuda.device(self.device):
                for h in everyone[rank][:3]:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    native.check(self.lib.lbm_ipc_open(hb, C.byref(p)), f"lbm_ipc_open(rank {rank})")
                    ptrs.append(p.value)
        self._peers[rank] = dict(a=ptrs[0], b=ptrs[1], flags=ptrs[2], nx=int(everyone[rank][3]))
        return self._peers[rank]

    def load(self, f: torch.Tensor):
        """replace the rank's populations (same shape) without leaving the IPC buffers"""
        self.t[self.cur].copy_(f)
        self.flow.f, self.flow.f_next = self.t[self.cur], self.t[1 - self.cur]

    def _slab_struct(self) -> "native.LbmSlab":
        if self.flow.f.data_ptr() != self.t[self.cur].data_ptr():
            raise RuntimeError("flow.f was replaced: slab populations must stay in the engine's IPC buffers "
                               "(use engine.load(tensor))")
        a, b = ("a", "b") if self.cur == 0 else ("b", "a")
        s = native.LbmSlab()
        s.lo_a, s.lo_b, s.hi_a, s.hi_b = self.lo[a], self.lo[b], self.hi[a], self.hi[b]
        s.lo_nx, s.hi_nx = self.lo["nx"], self.hi["nx"]
        s.signal_lo = self.lo["flags"] + 8        # the lo neighbour's slot 1 = "written by my hi neighbour"
        s.signal_hi = self.hi["flags"]            # the hi neighbour's slot 0 = "written by my lo neighbour"
        s.wait_slots = self.flags.ptr.value
        s.epoch = self.epoch
        return s

    def _launch_step_with_moments(self, f, g, scratch, out):
        """one lock-stepped step whose kernels reduce (sum 0.5|u|^2, max |u|^2) over THIS rank's nodes; GlobalSum /
        GlobalMax combine the ranks"""
        s = self._slab_struct()
        native.check(self.lib.lbm_slab_step_moments(C.byref(self.desc), C.byref(s), f.data_ptr(), g.data_ptr(),
                                                    scratch.data_ptr(), scratch.numel(), out.data_ptr(),
                                                    native._stream_ptr(self.device)), "lbm_slab_step_moments")

    def _after_moments_step(self, f, g):
        self.epoch += 1
        self.cur = 1 - self.cur
        self.flow.f, self.flow.f_next = self.t[self.cur], self.t[1 - self.cur]

    def step(self, n: int = 1):
        if n <= 0:
            return
        flow = self.flow
        self.refresh_parameters()
        flow._b200_moments = None
        s = self._slab_struct()
        with torch.cuda.device(self.device):
            native.check(self.lib.lbm_slab_step_n(C.byref(self.desc), C.byref(s), self.t[self.cur].data_ptr(),
                                                  self.t[1 - self.cur].data_ptr(), n, native._stream_ptr(self.device)),
                         "lbm_slab_step_n")
        self.epoch += n
        if n % 2 == 1:
            self.cur = 1 - self.cur
        flow.f, flow.f_next = self.t[self.cur], self.t[1 - self.cur]

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.dec.world > 1:
            dist.barrier(group=self.group)
        for rank, p in self._peers.items():
            if rank != self.dec.rank:
                for k in ("a", "b", "flags"):
                    self.lib.lbm_ipc_close(C.c_void_p(p[k]))
        self._peers = {}
        if self.dec.world > 1:
            dist.barrier(group=self.group)
        self.flow.f = self.flow.f.clone()
        self.flow._f_next = None
        self.t = None                      # the tensors below alias memory that is about to be freed
        for b in self.buf + [self.flags]:
            b.free()


def _owns_outlet_plane(boundary, dec: SlabDecomposition) -> bool:
    """x-normal outlet planes exist on one rank only: the last rank for +x, the first for -x"""
    direction = getattr(boundary, "direction", None)
    if direction is None or int(direction[0]) == 0:
        return True
    return dec.rank == (dec.world - 1 if int(direction[0]) > 0 else 0)


class SlabSimulation(Simulation):
    """`Simulation` whose flow is one x-slab of a larger lattice.

    Boundaries are built from the flow's `pre_/post_boundaries` like on one GPU, with local masks; an outlet
    whose plane is normal to x is active only on the rank that owns that global plane (its transformer entry
    stays in the list on every rank so that labels mean the same everywhere)."""

    def __init__(self, flow, collision, reporter, streaming_strategy=StreamingStrategy.POST_STREAMING,
                 decomposition: Optional[SlabDecomposition] = None, group=None):
        self.decomposition = decomposition or flow.decomposition
        super().__init__(flow, collision, reporter, streaming_strategy)
        # the engine maps the neighbours' GPU buffers; on a CPU context only the host side (masks) exists
        self._b200_engine = SlabEngine(self, self.decomposition, group) if flow.f.is_cuda else None

    def _build_masks(self):
        dec = self.decomposition
        for b in self.pre_boundaries + self.post_boundaries:
            b._slab_disabled = not _owns_outlet_plane(b, dec)
            if not b._slab_disabled and getattr(b, "direction", None) is not None and int(b.direction[0]) != 0:
                if dec.nx_local < 2:
                    raise ValueError("the rank owning an x-normal outlet needs at least two planes")
        super()._build_masks()

    def _boundary_masks(self, boundary, shape):
        if getattr(boundary, "_slab_disabled", False):
            return None, None
        return super()._boundary_masks(boundary, shape)

    def close(self):
        if self._b200_engine is not None:
            self._b200_engine.close()

    def dump(self, filename, group=None):
        """Checkpoint of the GLOBAL lattice in `Flow.dump`'s format (lettuce/_flow.py:258-262), written by rank 0:
        a single-GPU `Flow.load` or a run with another number of slabs can restart from it."""
        dump_slabs(self.flow, self.decomposition, filename, group)

    def load(self, filename, group=None):
        """Restart from a global checkpoint (`Flow.dump` / `SlabSimulation.dump`): every rank takes its x-range."""
        local = load_slabs(self.flow, self.decomposition, filename, group)
        if self._b200_engine is not None:
            self._b200_engine.load(local)          # populations stay in the IPC-exported buffers
        else:
            self.flow.f = local
            self.flow._f_next = None


def gather_populations(flow, dec: SlabDecomposition, group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """The global populations `[q, nx_global, ...]` in host memory on rank `dst` (None elsewhere).  Slabs travel
    one at a time (point-to-point, on the device the populations live on), so rank `dst` needs one slab of
    staging memory, not the whole lattice, on its GPU; the device-to-host copies go to pinned memory and overlap
    the next slab's transfer."""
    local = flow.f
    if dec.world == 1:
        return local.detach().cpu()
    if dec.rank != dst:
        dist.send(local.contiguous(), dst, group=group)
        return None
    shape = list(local.shape)
    shape[1] = dec.nx_global
    whole = torch.empty(shape, dtype=local.dtype, pin_memory=local.is_cuda)
    for r in range(dec.world):
        if r == dst:
            part = local
        else:
            pshape = list(local.shape)
            pshape[1] = dec.sizes[r]
            part = torch.empty(pshape, dtype=local.dtype, device=local.device)
            dist.recv(part, r, group=group)
        whole[:, dec.offsets[r]:dec.offsets[r] + dec.sizes[r]].copy_(part, non_blocking=True)
    if local.is_cuda:
        torch.cuda.synchronize(local.device)
    return whole


def dump_slabs(flow, dec: SlabDecomposition, filename, group=None):
    whole = gather_populations(flow, dec, group, dst=0)
    if dec.rank == 0:
        import pickle
        with open(filename, "wb") as fh:
            pickle.dump(whole.numpy(), fh)
    if dec.world > 1:
        dist.barrier(group=group)           # the file is complete when any rank returns


def load_slabs(flow, dec: SlabDecomposition, filename, group=None) -> torch.Tensor:
    """This rank's x-range of the global checkpoint `filename`, on the flow's device and in its dtype.  Rank 0
    reads the file and sends every other rank its planes."""
    like = flow.f
    expect = list(like.shape)
    if dec.rank == 0:
        import pickle
        with open(filename, "rb") as fh:
            whole = torch.as_tensor(pickle.load(fh))
        full = list(expect)
        full[1] = dec.nx_global
        ok = list(whole.shape) == full
        if dec.world > 1:
            dist.broadcast_object_list([ok], src=0, group=group)
        if not ok:
            raise ValueError(f"checkpoint holds populations of shape {list(whole.shape)}, the run needs {full}")
        for r in range(1, dec.world):
            part = whole[:, dec.offsets[r]:dec.offsets[r] + dec.sizes[r]]
            dist.send(part.to(device=like.device, dtype=like.dtype).contiguous(), r, group=group)
        return whole[:, dec.x0:dec.x1].to(device=like.device, dtype=like.dtype).contiguous()
    flag = [None]
    dist.broadcast_object_list(flag, src=0, group=group)
    if not flag[0]:
        raise ValueError("checkpoint does not match the global lattice (see rank 0)")
    local = torch.empty(expect, dtype=like.dtype, device=like.device)
    dist.recv(local, 0, group=group)
    return local


class _GlobalReduce(Observable):
    op = None

    def __init__(self, observable: Observable, group=None):
        super().__init__(observable.flow)
        self.observable = observable
        self.group = group
        # the wrapped observable may take its rank-local value from a step kernel (lbm_slab_step_moments)
        self.fused_with_step = bool(getattr(observable, "fused_with_step", False))

    def __call__(self, f=None):
        v = self.observable(f).clone()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(v, op=self.op, group=self.group)
        return v


class GlobalSum(_GlobalReduce):
    """sum of a rank-local additive observable (energy, mass) over all slabs"""
    op = dist.ReduceOp.SUM


class GlobalMax(_GlobalReduce):
    """max of a rank-local observable (maximum velocity) over all slabs"""
    op = dist.ReduceOp.MAX


class SlabEnstrophy(Observable):
    """Global enstrophy of a slab-decomposed periodic flow (observable_reporter.py:45-68).  The 6th-order
    stencil reaches three planes across each cut, so the velocity field -- not the populations -- of three
    boundary planes is exchanged with both neighbours (a real exchange step: NCCL send/recv), the curl is
    evaluated on the extended slab and summed over the owned planes only, then all-reduced."""
    _HALO = 3

    def __init__(self, flow, decomposition: SlabDecomposition = None, group=None):
        super().__init__(flow)
        self.dec = decomposition or flow.decomposition
        self.group = group

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        units, st, dec, h = self.flow.units, self.flow.stencil, self.dec, self._HALO
        _, u = native.moments(st, f, want_rho=False)
        if dec.world > 1:
            if dec.nx_local < h:
                raise ValueError(f"slabs must be at least {h} planes thick for the enstrophy stencil")
            first, last = u[:, :h].contiguous(), u[:, -h:].contiguous()
            from_lo, from_hi = torch.empty_like(last), torch.empty_like(first)
            # order matters when both neighbours are the same rank (world 2): sends and receives pair up in order
            ops = [dist.P2POp(dist.isend, last, dec.hi, self.group), dist.P2POp(dist.isend, first, dec.lo, self.group),
                   dist.P2POp(dist.irecv, from_lo, dec.lo, self.group), dist.P2POp(dist.irecv, from_hi, dec.hi, self.group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            ext = torch.cat([from_lo, u, from_hi], dim=1).contiguous()
            mask = torch.zeros(ext.shape[1:], dtype=torch.uint8, device=ext.device)
            mask[h:-h] = 1
            w2 = native.reduce(st, native.ENSTROPHY, ext, mask)
            dist.all_reduce(w2, op=dist.ReduceOp.SUM, group=self.group)
        else:
            w2 = native.reduce(st, native.ENSTROPHY, u)
        dx = units.convert_length_to_pu(1.0)
        scale = units.convert_velocity_to_pu(1.0) / dx
        return w2 * scale ** 2 * dx ** st.d


def make_tgv_slab_simulation(context, global_resolution, reynolds_number, mach_number, stencil, strategy,
                             collision_factory=None, reporter=None):
    """(flow, simulation, stepper) for this rank's slab of a global Taylor-Green vortex."""
    from .ext.collision import BGKCollision
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dec = SlabDecomposition(global_resolution[0], world, rank)
    flow = SlabTaylorGreenVortex(context, global_resolution, reynolds_number, mach_number, stencil, dec)
    collision = (collision_factory or (lambda fl: BGKCollision(fl.units.relaxation_parameter_lu)))(flow)
    sim = SlabSimulation(flow, collision, reporter or [], strategy, dec)
    return flow, sim, (lambda k: native.invoke_n(sim, k))
