"""ctypes binding of include/lbm_b200.h and the `invoke(simulation)` plug.

This module is what `Simulation._collide_and_stream` is pointed at, in the same
place where the reference installs its generated extension
(lettuce/_simulation.py:229) and with the same call signature
(`invoke(simulation)`, lettuce/cuda_native/_template.py:35-39): one call = one
time step, `flow.f` and `flow.f_next` are swapped afterwards.

There is no CPU or torch fallback: if the shared library is missing, the
populations are not CUDA tensors, or an operator has no native kernel, this
raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBM_B200_LIB", os.path.join(_PKG, "liblbm_b200.so"))   # override for A/B experiments

LBM_MAX_OPS = 8
# enums of include/lbm_b200.h
D2Q9, D3Q19, D3Q27 = 0, 1, 2
F32, F64 = 0, 1
OP_NO_COLLISION, OP_BGK, OP_TRT, OP_KBC, OP_REGULARIZED, OP_SMAGORINSKY, OP_BGK_FORCED = 0, 1, 2, 3, 4, 5, 6
OP_BOUNCE_BACK, OP_EQUILIBRIUM, OP_OUTLET_P, OP_ANTI_BOUNCE_BACK, OP_IDENTITY = 16, 17, 18, 19, 20
NO_STREAMING, POST_STREAMING, PRE_STREAMING, DOUBLE_STREAMING = 0, 1, 2, 3
# Batches of at least this many POST_STREAMING steps run as (S C)^n = S (C S)^(n-1) C: one collide-only pass, n - 1
# pull steps, one stream-only pass.  The result is bit-identical to n push steps: streaming only moves values, and
# every kernel variant evaluates the collide phase with the same explicitly rounded operations (csrc/lbm_vec.cuh;
# scripts/check_packed_contraction.py, tests/test_gpu_parity.py::test_lazy_post_batches_match_push_steps).
# LBM_B200_LAZY_POST_MIN=n forces the threshold (0 = never).  Unset: automatic -- from 16 steps on, and only where the
# pull step runs the TMA-staged kernel while the push step cannot (D3Q27 KBC: push 0.83, pull 0.95 of the roofline;
# n + 1 passes instead of n).  Elsewhere push and pull are within a few percent and the extra pass does not pay
# (sphere D3Q27 TRT: +5 % at n = 20; D3Q19 BGK: break-even at n ~ 40).
_LAZY_ENV = os.environ.get("LBM_B200_LAZY_POST_MIN")
LAZY_POST_MIN_STEPS = int(_LAZY_ENV) if _LAZY_ENV is not None else 0
LAZY_POST_AUTO = _LAZY_ENV is None
LAZY_POST_AUTO_MIN_STEPS = 16
SUM_HALF_U2, MAX_U, SUM_F, SUM_F_INNER, SUM_F_MASKED, ENSTROPHY = range(6)
# lbm_step_moments_state: which state the reductions fused into a step describe
MOMENTS_UNAVAILABLE, MOMENTS_OF_OUTPUT, MOMENTS_OF_INPUT = 0, 1, 2


class LbmOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("axis", C.c_int32), ("side", C.c_int32), ("_pad", C.c_int32),
                ("p0", C.c_double), ("p1", C.c_double),
                ("rho", C.c_void_p), ("u", C.c_void_p),
                ("rho_stride", C.c_int64 * 3), ("u_stride", C.c_int64 * 4),
                ("force", C.c_double * 3), ("ueq_scale", C.c_double), ("source_scale", C.c_double)]


class LbmLattice(C.Structure):
    _fields_ = [("stencil", C.c_int32), ("dtype", C.c_int32),
                ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("_pad", C.c_int32)]


class LbmHalo(C.Structure):
    _fields_ = [("in_lo", C.c_void_p), ("in_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_hi", C.c_void_p),
                ("in_lo_qstride", C.c_int64), ("in_hi_qstride", C.c_int64),
                ("out_lo_qstride", C.c_int64), ("out_hi_qstride", C.c_int64),
                ("label_lo", C.c_void_p), ("label_hi", C.c_void_p),
                ("frozen_lo", C.c_void_p), ("frozen_hi", C.c_void_p)]


class LbmStepDesc(C.Structure):
    _fields_ = [("lat", LbmLattice), ("streaming", C.c_int32), ("n_ops", C.c_int32),
                ("collision_index", C.c_int32), ("variant", C.c_int32),
                ("ops", LbmOp * LBM_MAX_OPS),
                ("labels", C.c_void_p), ("frozen", C.c_void_p),
                ("general_nodes", C.c_void_p), ("n_general", C.c_int64), ("halo", LbmHalo)]


class LbmSlab(C.Structure):
    _fields_ = [("lo_a", C.c_void_p), ("lo_b", C.c_void_p), ("hi_a", C.c_void_p), ("hi_b", C.c_void_p),
                ("lo_nx", C.c_int32), ("hi_nx", C.c_int32),
                ("signal_lo", C.c_void_p), ("signal_hi", C.c_void_p), ("wait_slots", C.c_void_p),
                ("epoch", C.c_uint64)]


class LbmLinks(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("n", C.c_int64),
                ("node", C.c_void_p), ("q", C.c_void_p), ("d", C.c_void_p), ("bounced", C.c_void_p),
                ("force_scratch", C.c_void_p), ("force", C.c_void_p)]


EXPORTS = ["lbm_apply_links", "lbm_links_scratch_doubles", "lbm_step_links_n", "lbm_ipc_alloc", "lbm_ipc_open",
           "lbm_ipc_close", "lbm_ipc_free", "lbm_slab_step_n", "lbm_slab_step_moments",
           "lbm_step", "lbm_step_n", "lbm_step_moments", "lbm_step_moments_n", "lbm_step_moments_state",
           "lbm_step_moments_scratch_bytes",
           "lbm_pack_masks", "lbm_list_general_nodes", "lbm_equilibrium", "lbm_initialize_fneq", "lbm_moments",
           "lbm_reduce_scratch_bytes",
           "lbm_reduce", "lbm_run_host", "lbm_run_host_release", "lbm_abi_version", "lbm_status_string",
           "lbm_last_cuda_error", "lbm_launch_count", "lbm_step_variant_name"]

_lib = None


def lib() -> C.CDLL:
    """Load liblbm_b200.so (built by `python -m lettuce_b200.build`).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m lettuce_b200.build` "
                           f"(lettuce_b200 has no CPU or torch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.lbm_step.argtypes = [C.POINTER(LbmStepDesc), vp, vp, vp]
    L.lbm_step.restype = i32
    L.lbm_step_n.argtypes = [C.POINTER(LbmStepDesc), vp, vp, i64, vp]
    L.lbm_step_n.restype = i32
    L.lbm_step_moments_state.argtypes = [C.POINTER(LbmStepDesc)]
    L.lbm_step_moments_state.restype = i32
    L.lbm_step_moments_n.argtypes = [C.POINTER(LbmStepDesc), vp, vp, i64, vp, C.c_size_t, vp, vp]
    L.lbm_step_moments_n.restype = C.c_int
    L.lbm_step_moments_scratch_bytes.argtypes = [C.POINTER(LbmStepDesc)]
    L.lbm_step_moments_scratch_bytes.restype = C.c_size_t
    L.lbm_step_moments.argtypes = [C.POINTER(LbmStepDesc), vp, vp, vp, C.c_size_t, vp, vp]
    L.lbm_step_moments.restype = i32
    L.lbm_slab_step_moments.argtypes = [C.POINTER(LbmStepDesc), C.POINTER(LbmSlab), vp, vp, vp, C.c_size_t, vp, vp]
    L.lbm_slab_step_moments.restype = i32
    L.lbm_run_host_release.restype = i32
    L.lbm_links_scratch_doubles.argtypes = [i64]
    L.lbm_links_scratch_doubles.restype = i64
    L.lbm_apply_links.argtypes = [C.POINTER(LbmStepDesc), C.POINTER(LbmLinks), vp, vp, vp]
    L.lbm_apply_links.restype = i32
    L.lbm_step_links_n.argtypes = [C.POINTER(LbmStepDesc), C.POINTER(LbmLinks), i32, vp, vp, i64, vp]
    L.lbm_step_links_n.restype = i32
    L.lbm_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(vp), vp]
    L.lbm_ipc_alloc.restype = i32
    L.lbm_ipc_open.argtypes = [vp, C.POINTER(vp)]
    L.lbm_ipc_open.restype = i32
    L.lbm_ipc_close.argtypes = [vp]
    L.lbm_ipc_close.restype = i32
    L.lbm_ipc_free.argtypes = [vp]
    L.lbm_ipc_free.restype = i32
    L.lbm_slab_step_n.argtypes = [C.POINTER(LbmStepDesc), C.POINTER(LbmSlab), vp, vp, i64, vp]
    L.lbm_slab_step_n.restype = i32
    L.lbm_pack_masks.argtypes = [C.POINTER(LbmStepDesc), vp, vp, vp, vp, vp]
    L.lbm_pack_masks.restype = i32
    L.lbm_list_general_nodes.argtypes = [C.POINTER(LbmLattice), vp, vp, i64, vp, vp]
    L.lbm_list_general_nodes.restype = i32
    L.lbm_equilibrium.argtypes = [C.POINTER(LbmLattice), vp, C.POINTER(i64), vp, C.POINTER(i64), vp, vp]
    L.lbm_equilibrium.restype = i32
    L.lbm_initialize_fneq.argtypes = [C.POINTER(LbmLattice), vp, vp, C.c_double, C.c_double, vp, vp]
    L.lbm_initialize_fneq.restype = i32
    L.lbm_moments.argtypes = [C.POINTER(LbmLattice), vp, vp, vp, vp]
    L.lbm_moments.restype = i32
    L.lbm_reduce_scratch_bytes.argtypes = [C.POINTER(LbmLattice)]
    L.lbm_reduce_scratch_bytes.restype = C.c_size_t
    L.lbm_reduce.argtypes = [C.POINTER(LbmLattice), i32, vp, vp, vp, vp, vp]
    L.lbm_reduce.restype = i32
    L.lbm_run_host.argtypes = [C.POINTER(LbmStepDesc), vp, vp, i64, vp]
    L.lbm_run_host.restype = i32
    L.lbm_abi_version.restype = i32
    L.lbm_status_string.argtypes = [i32]
    L.lbm_status_string.restype = C.c_char_p
    L.lbm_last_cuda_error.restype = C.c_char_p
    L.lbm_launch_count.restype = i64
    L.lbm_step_variant_name.argtypes = [C.POINTER(LbmStepDesc)]
    L.lbm_step_variant_name.restype = C.c_char_p
    _lib = L
    return L


def check(status: int, what: str = "lbm call"):
    if status != 0:
        L = lib()
        msg = L.lbm_status_string(status).decode()
        if status == -3:
            msg += ": " + L.lbm_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed: {msg}")


def launch_count() -> int:
    return int(lib().lbm_launch_count())


# --------------------------------------------------------------------------
# descriptor construction from lettuce-style objects (duck typed by class name,
# so the reference's own classes are accepted as well as lettuce_b200's)
# --------------------------------------------------------------------------
_STENCIL_IDS = {"D2Q9": D2Q9, "D3Q19": D3Q19, "D3Q27": D3Q27}
_KIND_BY_NAME = {"NoCollision": OP_NO_COLLISION, "BGKCollision": OP_BGK, "TRTCollision": OP_TRT,
                 "KBCCollision": OP_KBC, "RegularizedCollision": OP_REGULARIZED,
                 "SmagorinskyCollision": OP_SMAGORINSKY, "BounceBackBoundary": OP_BOUNCE_BACK,
                 "EquilibriumBoundaryPU": OP_EQUILIBRIUM, "EquilibriumOutletP": OP_OUTLET_P,
                 "AntiBounceBackOutlet": OP_ANTI_BOUNCE_BACK}


_FORCES = ("Guo", "ShanChen")


def force_kind(force) -> str:
    for cls in type(force).__mro__:
        if cls.__name__ in _FORCES:
            return cls.__name__
    raise NotImplementedError(f"force {type(force).__name__} has no B200 kernel (supported: {_FORCES})")


def op_kind(op) -> int:
    """Native kind of a collision / boundary object; most-derived class name wins."""
    for cls in type(op).__mro__:
        if cls.__name__ in _KIND_BY_NAME:
            kind = _KIND_BY_NAME[cls.__name__]
            if kind == OP_BGK and getattr(op, "force", None) is not None:
                force_kind(op.force)
                return OP_BGK_FORCED
            return kind
    raise NotImplementedError(
        f"{type(op).__name__} has no B200 kernel (supported: {sorted(_KIND_BY_NAME)}); "
        f"lettuce_b200 does not fall back to torch")


def stencil_id(stencil) -> int:
    for cls in type(stencil).__mro__:
        if cls.__name__ in _STENCIL_IDS:
            return _STENCIL_IDS[cls.__name__]
    raise NotImplementedError(f"stencil {type(stencil).__name__} has no B200 kernel (D2Q9, D3Q19, D3Q27)")


def dtype_id(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise NotImplementedError(f"dtype {dtype} has no B200 kernel (float32, float64); the reference's native "
                              f"dispatch has the same limit (cuda_native/_template.py:76)")


def lattice_of(stencil, resolution, dtype) -> LbmLattice:
    res = [int(r) for r in resolution]
    if len(res) != stencil.d:
        raise ValueError(f"resolution {res} does not match a {stencil.d}-dimensional stencil")
    nx, ny = res[0], res[1]
    nz = res[2] if len(res) == 3 else 1
    return LbmLattice(stencil_id(stencil), dtype_id(dtype), nx, ny, nz, 0)


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} lives on {t.device}: lettuce_b200 runs on CUDA devices only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous [q, *resolution]")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _direction_axis_side(direction):
    direction = [int(round(float(c))) for c in direction]
    nz = [i for i, c in enumerate(direction) if c != 0]
    if len(nz) != 1 or abs(direction[nz[0]]) != 1:
        raise ValueError(f"outlet direction must have exactly one entry of +-1, got {direction}")
    return nz[0], direction[nz[0]]


class Engine:
    """Everything `invoke` needs for one Simulation: the packed descriptor, the
    label / frozen-slot fields and the boundary parameter tensors it borrows."""

    def __init__(self, simulation, dry: bool = False):
        """`dry=True` only translates the simulation into the descriptor (no CUDA tensors required, no
        mask packing): used to check the duck-typed translation against the reference's own classes."""
        self.lib = lib()
        flow = simulation.flow
        self.flow = flow
        self.simulation = simulation
        if not dry:
            _require_cuda(flow.f, "flow.f")
        self.device = flow.f.device
        self.lat = lattice_of(flow.stencil, flow.f.shape[1:], flow.f.dtype)
        # the kernels evaluate the quadratic equilibrium (quadratic_equilibrium.py:11-24; the less-memory variant
        # is the same polynomial); the reference's collisions would call any other flow.equilibrium
        equilibrium = type(getattr(flow, "equilibrium", None)).__name__
        if equilibrium not in ("QuadraticEquilibrium", "QuadraticEquilibriumLessMemory", "NoneType"):
            raise NotImplementedError(f"equilibrium {equilibrium} has no B200 kernel (quadratic equilibrium only)")
        transformer = list(simulation.transformer)
        if len(transformer) > LBM_MAX_OPS:
            raise NotImplementedError(f"at most {LBM_MAX_OPS} transformer entries, got {len(transformer)}")
        self.desc = LbmStepDesc()
        self.desc.lat = self.lat
        self.desc.streaming = int(simulation.streaming_strategy.value)
        self.desc.n_ops = len(transformer)
        self.desc.collision_index = int(simulation.collision_index)
        self.desc.variant = 0
        self._keep: List[torch.Tensor] = []     # tensors whose pointers the descriptor borrows
        self._eq_state = {}                     # EquilibriumBoundaryPU entries: converted tensors + source versions
        self.transformer = transformer
        for i, op in enumerate(transformer):
            self._fill_op(i, op)
        self.refresh_parameters()
        self.labels: Optional[torch.Tensor] = None
        self.frozen: Optional[torch.Tensor] = None
        self._pull_is_staged: Optional[bool] = None     # would a PRE_STREAMING step run the TMA-staged kernel?
        if dry:
            self.variant_name = "dry"
            return
        if len(transformer) > 1 or getattr(simulation, "no_collision_mask", None) is not None:
            # boundaries in the transformer list, or masks of post-streaming boundaries (EbbSimulation)
            self._pack_masks()
        self.variant_name = self.lib.lbm_step_variant_name(C.byref(self.desc)).decode()

    # -- descriptor pieces ---------------------------------------------------
    def _fill_op(self, i, op):
        o = self.desc.ops[i]
        kind = op_kind(op)
        o.kind = kind
        flow, units = self.flow, self.flow.units
        if kind == OP_SMAGORINSKY and getattr(op, "force", None) is not None:
            raise NotImplementedError("SmagorinskyCollision with a force term has no B200 kernel")
        if kind == OP_KBC:
            # the reference replaces tau by the flow's relaxation parameter on first call
            # (lettuce/ext/_collision/kbc_collision.py:97-99); mirror that state change
            op.tau = units.relaxation_parameter_lu
            op.beta = 1.0 / (2 * op.tau)
        if kind == OP_REGULARIZED:
            op.tau = units.relaxation_parameter_lu      # regularized_collision.py:19, same quirk
        if kind == OP_EQUILIBRIUM:
            rho, u = self._equilibrium_boundary_values(op)
            d = flow.stencil.d
            if rho.dim() != d + 1 or u.dim() != d + 1 or rho.shape[0] != 1:
                raise ValueError("EquilibriumBoundaryPU needs pressure of shape [1,...] and velocity [d or 1,...]")
            self._keep += [rho, u]
            self._eq_state[i] = (rho, u, self._tensor_key(op.pressure), self._tensor_key(op.velocity))
            o.rho, o.u = rho.data_ptr(), u.data_ptr()
            for a in range(d):
                o.rho_stride[a] = rho.stride(a + 1) if rho.shape[a + 1] > 1 else 0
                o.u_stride[a + 1] = u.stride(a + 1) if u.shape[a + 1] > 1 else 0
            o.u_stride[0] = u.stride(0) if u.shape[0] > 1 else 0
        if kind in (OP_OUTLET_P, OP_ANTI_BOUNCE_BACK):
            direction = getattr(op, "direction", None)
            if direction is None:       # the reference keeps only index lists (equilibrium_outlet_p.py:36-49)
                direction = [0 if isinstance(ix, slice) else (1 if ix == -1 else -1) for ix in op.index]
            o.axis, o.side = _direction_axis_side(direction)
            if getattr(op, "_slab_disabled", False):
                o.side = 0          # the plane belongs to another rank's slab (lettuce_b200/slab.py)
            if kind == OP_OUTLET_P:
                o.p0 = float(op.rho_outlet)

    @staticmethod
    def _tensor_key(t):
        return (id(t), getattr(t, "_version", None))

    def _equilibrium_boundary_values(self, op):
        """density and velocity of an EquilibriumBoundaryPU in lattice units, lattice dtype, on the device"""
        units, f = self.flow.units, self.flow.f
        rho = units.convert_pressure_pu_to_density_lu(op.pressure)
        u = units.convert_velocity_to_lu(op.velocity)
        return (rho.to(device=self.device, dtype=f.dtype).contiguous(),
                u.to(device=self.device, dtype=f.dtype).contiguous())

    def refresh_parameters(self):
        """Re-read operator parameters before a launch, like the reference's generated call does on every step
        (cuda_native/ext/_collision/bgk_collision.py:29-33, ext/_boundary/equilibrium_pu.py:15-18,55-58): scalars
        are copied, tensors (inlet velocity / pressure, acceleration) only when they were replaced or modified in
        place (torch's version counter), so the common case costs no device work."""
        for i, (rho, u, pkey, ukey) in self._eq_state.items():
            op = self.transformer[i]
            if (self._tensor_key(op.pressure), self._tensor_key(op.velocity)) != (pkey, ukey):
                new_rho, new_u = self._equilibrium_boundary_values(op)
                if new_rho.shape != rho.shape or new_u.shape != u.shape:
                    raise RuntimeError("EquilibriumBoundaryPU changed the shape of its velocity / pressure; "
                                       "create a new Simulation")
                rho.copy_(new_rho); u.copy_(new_u)          # same storage: the descriptor's pointers stay valid
                self._eq_state[i] = (rho, u, self._tensor_key(op.pressure), self._tensor_key(op.velocity))
        for i, op in enumerate(self.transformer):
            o = self.desc.ops[i]
            if o.kind == OP_BGK:
                o.p0 = float(op.tau)
            elif o.kind == OP_BGK_FORCED:
                # Guo: u_eq = a/(2 rho), source (1 - 1/(2 tau_force)) ... ; ShanChen: u_eq = tau a / rho, no source
                # (lettuce/ext/_force/guo.py:16-38, shan_chen.py:13-26)
                force = op.force
                o.p0 = float(op.tau)
                a_t = force.acceleration
                key = (id(a_t), getattr(a_t, "_version", None))
                if getattr(self, "_acc_key", None) != key:       # reading a CUDA tensor synchronises: only when it changed
                    acc = [float(a) for a in torch.as_tensor(a_t).flatten().tolist()]
                    if len(acc) != self.flow.stencil.d:
                        raise ValueError(f"acceleration must have {self.flow.stencil.d} components, got {len(acc)}")
                    self._acc_key, self._acc = key, acc
                for a in range(3):
                    o.force[a] = self._acc[a] if a < len(self._acc) else 0.0
                o.ueq_scale = float(force.ueq_scaling_factor)
                o.source_scale = (1.0 - 1.0 / (2.0 * float(force.tau))) if force_kind(force) == "Guo" else 0.0
            elif o.kind == OP_TRT:
                o.p0, o.p1 = float(op.tau_plus), float(op.tau_minus)
            elif o.kind in (OP_KBC, OP_REGULARIZED):
                o.p0 = float(op.tau)
            elif o.kind == OP_SMAGORINSKY:
                o.p0, o.p1 = float(op.tau), float(op.constant)

    def _pack_masks(self):
        sim = self.simulation
        ncm, nsm = sim.no_collision_mask, sim.no_streaming_mask
        if ncm is None or nsm is None:
            raise RuntimeError("boundaries are present but the simulation has no masks "
                               "(lettuce/cuda_native/_template.py python_pre asserts the same)")
        ncm = ncm.to(device=self.device, dtype=torch.uint8).contiguous()
        nsm = nsm.to(device=self.device, dtype=torch.uint8).contiguous()
        n = ncm.numel()
        self.labels = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.frozen = torch.empty(n, dtype=torch.int32, device=self.device)
        check(self.lib.lbm_pack_masks(C.byref(self.desc), ncm.data_ptr(), nsm.data_ptr(),
                                      self.labels.data_ptr(), self.frozen.data_ptr(), _stream_ptr(self.device)),
              "lbm_pack_masks")
        self.desc.labels = self.labels.data_ptr()
        self.desc.frozen = self.frozen.data_ptr()
        self._list_general_nodes()

    def _list_general_nodes(self):
        """compact list of the nodes the sparse general-nodes kernel has to visit"""
        count = torch.zeros((), dtype=torch.int64, device=self.device)
        stream = _stream_ptr(self.device)
        check(self.lib.lbm_list_general_nodes(C.byref(self.lat), self.labels.data_ptr(), None, 0, count.data_ptr(),
                                              stream), "lbm_list_general_nodes")
        n_general = int(count.item())
        self.general_nodes = torch.empty(max(n_general, 1), dtype=torch.int32, device=self.device)
        check(self.lib.lbm_list_general_nodes(C.byref(self.lat), self.labels.data_ptr(), self.general_nodes.data_ptr(),
                                              n_general, count.data_ptr(), stream), "lbm_list_general_nodes")
        self.desc.general_nodes = self.general_nodes.data_ptr()
        self.desc.n_general = n_general

    # -- stepping ------------------------------------------------------------
    def _buffers(self):
        flow = self.flow
        f, g = flow.f, flow.f_next
        _require_cuda(f, "flow.f")
        _require_cuda(g, "flow.f_next")
        if g.shape != f.shape or g.dtype != f.dtype:
            raise RuntimeError("flow.f_next does not match flow.f")
        lat = self.lat
        expect = [flow.stencil.q, lat.nx, lat.ny] + ([lat.nz] if flow.stencil.d == 3 else [])
        if list(f.shape) != expect or dtype_id(f.dtype) != lat.dtype or f.device != self.device:
            raise RuntimeError(f"flow.f changed shape, dtype or device since the engine was built "
                               f"({list(f.shape)} {f.dtype} {f.device}); create a new Simulation")
        return f, g

    def moments_state(self) -> int:
        """Which state the reductions fused into a step describe for this simulation (`lbm_step_moments_state`):
        MOMENTS_OF_OUTPUT (NO / PRE streaming), MOMENTS_OF_INPUT (POST streaming: the report after step k rides on
        step k + 1) or MOMENTS_UNAVAILABLE (DOUBLE streaming)."""
        cached = getattr(self, "_moments_state", None)       # depends on the streaming mode only
        if cached is None:
            cached = int(self.lib.lbm_step_moments_state(C.byref(self.desc)))
            self._moments_state = cached
        return cached

    def _moments_scratch(self) -> torch.Tensor:
        need = int(self.lib.lbm_step_moments_scratch_bytes(C.byref(self.desc)))
        if need == 0:
            raise RuntimeError("lbm_step_moments is not available for this simulation (DOUBLE_STREAMING)")
        if getattr(self, "_moments_buf", None) is None or self._moments_buf.numel() < need:
            self._moments_buf = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._moments_buf

    def _launch_step_with_moments(self, f, g, scratch, out):
        check(self.lib.lbm_step_moments(C.byref(self.desc), f.data_ptr(), g.data_ptr(), scratch.data_ptr(),
                                        scratch.numel(), out.data_ptr(), _stream_ptr(self.device)),
              "lbm_step_moments")

    def step_with_moments(self) -> torch.Tensor:
        """One time step whose kernels also reduce (sum 0.5|u|^2, max |u|^2) in lattice units over this engine's
        nodes -- a float64 CUDA tensor of two entries, no second pass over the populations (`lbm_step_moments`).
        With MOMENTS_OF_OUTPUT the values describe the NEW state and are left on the flow for the observables to pick
        up (`fused_moments`); with MOMENTS_OF_INPUT they describe the state the step started from."""
        self.refresh_parameters()
        f, g = self._buffers()
        scratch = self._moments_scratch()
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        self.flow._b200_moments = None
        with torch.cuda.device(self.device):
            self._launch_step_with_moments(f, g, scratch, out)
        self._after_moments_step(f, g)
        if self.moments_state() == MOMENTS_OF_OUTPUT:
            new = self.flow.f
            self.flow._b200_moments = (new.data_ptr(), new._version, out)
        return out

    def _after_moments_step(self, f, g):
        self.flow.f, self.flow.f_next = g, f

    def steps_with_moments(self, n: int) -> torch.Tensor:
        """`n` time steps in ONE library call, every one with the fused reductions (`lbm_step_moments_n`): returns a
        float64 CUDA tensor `[m, 2]` whose row k holds (sum 0.5|u|^2, max |u|^2) of the state after step k + 1, with
        m = n when the steps describe the state they write (NO / PRE streaming) and m = n - 1 when they describe the
        state they read (POST streaming)."""
        state = self.moments_state()
        if state == MOMENTS_UNAVAILABLE or n <= 0:
            raise RuntimeError("steps_with_moments needs NO / PRE / POST streaming and n > 0")
        self.refresh_parameters()
        f, g = self._buffers()
        scratch = self._moments_scratch()
        m = n if state == MOMENTS_OF_OUTPUT else n - 1
        out = torch.empty((max(m, 1), 2), dtype=torch.float64, device=self.device)
        self.flow._b200_moments = None
        with torch.cuda.device(self.device):
            check(self.lib.lbm_step_moments_n(C.byref(self.desc), f.data_ptr(), g.data_ptr(), n, scratch.data_ptr(),
                                              scratch.numel(), out.data_ptr(), _stream_ptr(self.device)),
                  "lbm_step_moments_n")
        if n & 1:
            self.flow.f, self.flow.f_next = g, f
        if state == MOMENTS_OF_OUTPUT:
            new = self.flow.f
            self.flow._b200_moments = (new.data_ptr(), new._version, out[n - 1])
        return out[:m]

    def apply_links(self, boundary):
        """One post-streaming link boundary (ext/bounce_back.py) on the populations the last `step(1)` produced:
        `lbm_apply_links` with the step's input (now `flow.f_next`, left intact by the two-buffer scheme) and its
        output (`flow.f`).  Updates the boundary's force if it was built with calc_force."""
        if self.desc.streaming != POST_STREAMING:
            raise RuntimeError("post-streaming boundaries need StreamingStrategy.POST_STREAMING "
                               "(collide, stream, boundary: ebb_simulation.py:71-104)")
        f_post, f_pre = self._buffers()
        links = boundary.link_descriptor(f_post)
        with torch.cuda.device(self.device):
            check(self.lib.lbm_apply_links(C.byref(self.desc), C.byref(links), f_pre.data_ptr(), f_post.data_ptr(),
                                           _stream_ptr(self.device)), "lbm_apply_links")

    def step_with_links(self, boundaries, n: int):
        """`n` time steps, each followed by the post-streaming `boundaries` in order, in one library call
        (`lbm_step_links_n`)."""
        if n <= 0:
            return
        if self.desc.streaming != POST_STREAMING:
            raise RuntimeError("post-streaming boundaries need StreamingStrategy.POST_STREAMING")
        self.refresh_parameters()
        f, g = self._buffers()
        self.flow._b200_moments = None
        array = (LbmLinks * max(len(boundaries), 1))()
        for i, boundary in enumerate(boundaries):
            array[i] = boundary.link_descriptor(f)
        with torch.cuda.device(self.device):
            check(self.lib.lbm_step_links_n(C.byref(self.desc), array, len(boundaries), f.data_ptr(), g.data_ptr(), n,
                                            _stream_ptr(self.device)), "lbm_step_links_n")
        if n & 1:
            self.flow.f, self.flow.f_next = g, f

    def _variant_of(self, streaming: int, stream_only: bool = False) -> LbmStepDesc:
        d = LbmStepDesc.from_buffer_copy(self.desc)
        d.streaming = streaming
        if stream_only:
            for i in range(d.n_ops):
                d.ops[i].kind = OP_NO_COLLISION if i == d.collision_index else OP_IDENTITY
        return d

    def _lazy_post(self, n: int) -> bool:
        """run a batch of n POST_STREAMING steps as collide-only + (n - 1) pull steps + stream-only?"""
        if LAZY_POST_MIN_STEPS > 0:
            return n >= LAZY_POST_MIN_STEPS
        if not LAZY_POST_AUTO or n < LAZY_POST_AUTO_MIN_STEPS:
            return False
        if self._pull_is_staged is None:
            name = self.lib.lbm_step_variant_name(C.byref(self._variant_of(PRE_STREAMING))).decode()
            self._pull_is_staged = "TMA" in name
        return self._pull_is_staged

    def step(self, n: int = 1):
        """Advance `n` time steps; afterwards flow.f holds the new populations.

        Long POST_STREAMING batches may use the identity (S C)^n = S (C S)^(n-1) C (see LAZY_POST_MIN_STEPS): one
        collide-only pass, n-1 steps of the pull kernel and one stream-only pass; bit-identical to n push steps."""
        if n <= 0:
            return
        self.refresh_parameters()
        f, g = self._buffers()
        flow = self.flow
        flow._b200_moments = None           # fused moments (step_with_moments) describe the state before this step
        bufs, cur = [f, g], 0
        with torch.cuda.device(self.device):
            stream = _stream_ptr(self.device)
            if self.desc.streaming == POST_STREAMING and self._lazy_post(n):
                first = self._variant_of(NO_STREAMING)
                middle = self._variant_of(PRE_STREAMING)
                last = self._variant_of(PRE_STREAMING, stream_only=True)
                check(self.lib.lbm_step(C.byref(first), bufs[0].data_ptr(), bufs[1].data_ptr(), stream), "lbm_step")
                cur = 1
                check(self.lib.lbm_step_n(C.byref(middle), bufs[cur].data_ptr(), bufs[1 - cur].data_ptr(), n - 1,
                                          stream), "lbm_step_n")
                cur ^= (n - 1) & 1
                check(self.lib.lbm_step(C.byref(last), bufs[cur].data_ptr(), bufs[1 - cur].data_ptr(), stream),
                      "lbm_step")
                cur ^= 1
            elif n == 1:
                check(self.lib.lbm_step(C.byref(self.desc), f.data_ptr(), g.data_ptr(), stream), "lbm_step")
                cur = 1
            else:
                check(self.lib.lbm_step_n(C.byref(self.desc), f.data_ptr(), g.data_ptr(), n, stream), "lbm_step_n")
                cur = n & 1
        flow.f, flow.f_next = bufs[cur], bufs[1 - cur]


class _OneOperator:
    """The minimal simulation-shaped object `Engine` needs to run ONE operator over the whole lattice."""

    def __init__(self, flow, op, is_collision: bool):
        from ._simulation import StreamingStrategy        # late: _simulation imports this module
        from .ext.collision import NoCollision
        self.flow = flow
        self.streaming_strategy = StreamingStrategy.NO_STREAMING
        if is_collision:
            self.transformer, self.collision_index = [op], 0
            self.no_collision_mask = self.no_streaming_mask = None
        else:
            # label 1 everywhere: the boundary acts on every node, as `boundary(flow)` does in the reference
            # before Simulation._collide blends it in by label (lettuce/_simulation.py:258-305)
            self.transformer, self.collision_index = [NoCollision(), op], 0
            shape = [int(n) for n in flow.f.shape]
            self.no_collision_mask = torch.ones(shape[1:], dtype=torch.uint8, device=flow.f.device)
            self.no_streaming_mask = torch.zeros(shape, dtype=torch.uint8, device=flow.f.device)
        self.collision = self.transformer[self.collision_index]


def apply_operator(op, flow, is_collision: bool) -> torch.Tensor:
    """`op(flow)` of the reference's operator contract (Collision.__call__, lettuce/_simulation.py:17-28;
    Boundary.__call__, lettuce/_flow.py:31-52): the operator applied to EVERY node of `flow.f`, returned as a new
    tensor -- one NO_STREAMING launch of the step kernel into a scratch buffer (no torch arithmetic).  Boundaries run
    through the general-nodes kernel with the whole lattice labelled as theirs."""
    key = (id(flow.f.untyped_storage()), tuple(flow.f.shape), flow.f.dtype, bool(is_collision))
    cached = getattr(op, "_b200_call_engine", None)
    if cached is None or cached[0] != key or cached[1].flow is not flow:
        cached = (key, Engine(_OneOperator(flow, op, is_collision)))
        try:
            op._b200_call_engine = cached
        except AttributeError:
            pass
    eng = cached[1]
    eng.refresh_parameters()
    f = flow.f
    _require_cuda(f, "flow.f")
    out = torch.empty_like(f)
    with torch.cuda.device(f.device):
        check(eng.lib.lbm_step(C.byref(eng.desc), f.data_ptr(), out.data_ptr(), _stream_ptr(f.device)), "lbm_step")
    return out


def describe(simulation):
    """The transformer list of `simulation` as the engine would see it: a list of dicts with the native op
    kind and its parameters.  Works on CPU simulations and on the reference's own `lettuce.Simulation`."""
    eng = Engine(simulation, dry=True)
    out = []
    for i in range(eng.desc.n_ops):
        o = eng.desc.ops[i]
        out.append(dict(kind=int(o.kind), axis=int(o.axis), side=int(o.side), p0=float(o.p0), p1=float(o.p1),
                        rho_stride=list(o.rho_stride), u_stride=list(o.u_stride)))
    return dict(stencil=int(eng.desc.lat.stencil), dtype=int(eng.desc.lat.dtype),
                resolution=[int(eng.desc.lat.nx), int(eng.desc.lat.ny), int(eng.desc.lat.nz)],
                streaming=int(eng.desc.streaming), collision_index=int(eng.desc.collision_index), ops=out)


def fused_moments(flow, f: torch.Tensor) -> Optional[torch.Tensor]:
    """(sum 0.5|u|^2, max |u|^2) of `f` in lattice units over the engine's nodes if a step kernel already reduced them
    (`Engine.step_with_moments`), else None.  Two cases: the step that WROTE `f` reduced its output and `f` has not
    been written through torch since; or `Simulation.__call__` is reporting a state whose moments were reduced by the
    step that followed it (`_b200_reported_moments`, set only around the reporter calls)."""
    reported = getattr(flow, "_b200_reported_moments", None)
    if reported is not None and f is flow.f:
        return reported
    cached = getattr(flow, "_b200_moments", None)
    if cached is None or f is not flow.f or cached[0] != f.data_ptr() or cached[1] != f._version:
        return None
    return cached[2]


def engine_of(simulation) -> Engine:
    eng = getattr(simulation, "_b200_engine", None)
    if eng is None or eng.flow is not simulation.flow:
        eng = Engine(simulation)
        simulation._b200_engine = eng
    return eng


def invoke(simulation):
    """One time step on the B200 engine; drop-in for the reference's generated
    `invoke(simulation)` (lettuce/cuda_native/_template.py:35-39)."""
    engine_of(simulation).step(1)


def invoke_n(simulation, num_steps: int):
    """`num_steps` time steps without returning to Python in between."""
    engine_of(simulation).step(int(num_steps))


# --------------------------------------------------------------------------
# moments and reductions on populations (Flow.rho/j/u and the observables)
# --------------------------------------------------------------------------
def moments(stencil, f: torch.Tensor, want_rho=True, want_u=True):
    """(rho [1,*res] or None, u [d,*res] or None) of populations `f` (lettuce/_flow.py:157-193)."""
    _require_cuda(f, "f")
    lat = lattice_of(stencil, f.shape[1:], f.dtype)
    res = list(f.shape[1:])
    rho = torch.empty([1, *res], dtype=f.dtype, device=f.device) if want_rho else None
    u = torch.empty([stencil.d, *res], dtype=f.dtype, device=f.device) if want_u else None
    with torch.cuda.device(f.device):
        check(lib().lbm_moments(C.byref(lat), f.data_ptr(), rho.data_ptr() if want_rho else None,
                                u.data_ptr() if want_u else None, _stream_ptr(f.device)), "lbm_moments")
    return rho, u


def equilibrium_field(stencil, rho: torch.Tensor, u: torch.Tensor, resolution) -> torch.Tensor:
    """[q, *resolution] equilibrium populations of broadcastable fields rho [1|.., *res|1] and
    u [d, *res|1] (lattice units), written by one kernel without full-size temporaries."""
    _require_cuda(u, "u")
    d = stencil.d
    res = [int(r) for r in resolution]
    if rho.dim() == d:
        rho = rho.unsqueeze(0)
    if rho.dim() != d + 1 or u.dim() != d + 1 or rho.shape[0] != 1 or u.shape[0] != d:
        raise ValueError(f"need rho [1, ...] and u [{d}, ...] of rank {d + 1}, got {list(rho.shape)} {list(u.shape)}")
    rho = rho.to(device=u.device, dtype=u.dtype).contiguous()
    u = u.contiguous()
    for a in range(d):
        if rho.shape[a + 1] not in (1, res[a]) or u.shape[a + 1] not in (1, res[a]):
            raise ValueError("rho / u do not broadcast to the resolution")
    lat = lattice_of(stencil, res, u.dtype)
    rs = (C.c_int64 * 3)(*[rho.stride(a + 1) if rho.shape[a + 1] > 1 else 0 for a in range(d)], *([0] * (3 - d)))
    us = (C.c_int64 * 4)(u.stride(0), *[u.stride(a + 1) if u.shape[a + 1] > 1 else 0 for a in range(d)],
                         *([0] * (3 - d)))
    f = torch.empty([stencil.q, *res], dtype=u.dtype, device=u.device)
    with torch.cuda.device(u.device):
        check(lib().lbm_equilibrium(C.byref(lat), rho.data_ptr(), rs, u.data_ptr(), us, f.data_ptr(),
                                    _stream_ptr(u.device)), "lbm_equilibrium")
    return f


def initialize_fneq(stencil, f: torch.Tensor, tau: float) -> torch.Tensor:
    """`initialize_f_neq` (lettuce/_flow.py:341-367) on the device: moments of `f` (4 N values), then ONE kernel
    that writes feq(rho, u) - w_q Q_q : Pi1 -- instead of the reference's chain of full-size torch temporaries
    (gradient stacks, two einsums, feq and fneq)."""
    _require_cuda(f, "f")
    rho, u = moments(stencil, f)
    lat = lattice_of(stencil, f.shape[1:], f.dtype)
    # the reference subtracts torch.eye(d) * cs**2, a float32 tensor, on the diagonal of Q (_flow.py:358-360)
    eye_cs2 = float(torch.tensor(stencil.cs ** 2, dtype=torch.float32))
    out = torch.empty_like(f)
    with torch.cuda.device(f.device):
        check(lib().lbm_initialize_fneq(C.byref(lat), rho.data_ptr(), u.data_ptr(), float(tau), eye_cs2,
                                        out.data_ptr(), _stream_ptr(f.device)), "lbm_initialize_fneq")
    return out


_scratch = {}


def reduce(stencil, what: int, t: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Deterministic reduction `what` (one of SUM_HALF_U2, MAX_U, SUM_F, SUM_F_INNER, SUM_F_MASKED,
    ENSTROPHY); returns a 0-d float64 CUDA tensor (no host sync)."""
    _require_cuda(t, "input")
    lat = lattice_of(stencil, t.shape[1:], t.dtype)
    L = lib()
    key = (t.device, torch.cuda.current_stream(t.device).cuda_stream)
    if key not in _scratch:
        _scratch[key] = torch.empty(int(L.lbm_reduce_scratch_bytes(C.byref(lat))), dtype=torch.uint8, device=t.device)
    out = torch.empty((), dtype=torch.float64, device=t.device)
    mptr = None
    if mask is not None:
        mask = mask.to(device=t.device, dtype=torch.uint8).contiguous()
        mptr = mask.data_ptr()
    with torch.cuda.device(t.device):
        check(L.lbm_reduce(C.byref(lat), what, t.data_ptr(), mptr, _scratch[key].data_ptr(), out.data_ptr(),
                           _stream_ptr(t.device)), "lbm_reduce")
    return out
