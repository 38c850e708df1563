"""Collision operators (API of lettuce/ext/_collision/{bgk,trt,kbc,no}_collision.py).

On the B200 engine these are parameter holders: the arithmetic lives in
csrc/lbm_core.cuh (`Collide<S, R, COLL>`), fused into the stream+collide kernel.
"""
from __future__ import annotations

from .._simulation import Collision

__all__ = ["NoCollision", "BGKCollision", "TRTCollision", "KBCCollision", "KBCCollision2D", "KBCCollision3D",
           "RegularizedCollision",
           "SmagorinskyCollision", "Force", "Guo", "ShanChen"]


class NoCollision(Collision):
    """identity (lettuce/ext/_collision/no_collision.py:9-11)"""


class BGKCollision(Collision):
    """f - (f - feq)/tau (lettuce/ext/_collision/bgk_collision.py:12-22).  `tau` is re-read
    before every launch, like the reference's generated call passes 1/tau per step."""

    def __init__(self, tau, force=None):
        self.tau = tau
        self.force = force          # None, Guo or ShanChen: the forced variant is its own kernel


class TRTCollision(Collision):
    """two relaxation times; tau_minus defaults to 1 (lettuce/ext/_collision/trt_collision.py:12-27)"""

    def __init__(self, tau, tau_minus=1.0):
        self.tau_plus = tau
        self.tau_minus = tau_minus


class KBCCollision(Collision):
    """entropic multi-relaxation (Karlin-Boesch-Chikatamarla), D2Q9 and D3Q27 only
    (lettuce/ext/_collision/kbc_collision.py:11-160).

    As in the reference, the constructor argument is ignored: on first use `tau` is replaced by
    `flow.units.relaxation_parameter_lu` and `beta = 1/(2 tau)` (kbc_collision.py:97-99)."""

    def __init__(self, tau: float = None):
        self.tau = tau
        self.beta = None


class RegularizedCollision(Collision):
    """regularized LBM of Latt & Chopard (lettuce/ext/_collision/regularized_collision.py:9-43).  Like KBC,
    the reference ignores the constructor argument and takes tau from the flow's units on first use."""

    def __init__(self, tau: float = None):
        self.tau = tau


class SmagorinskyCollision(Collision):
    """Smagorinsky LES model on BGK (lettuce/ext/_collision/smagorinsky_collision.py:9-40)"""

    def __init__(self, tau, smagorinsky_constant=0.17, force=None):
        if force is not None:
            raise NotImplementedError("SmagorinskyCollision with a force term has no B200 kernel")
        self.force = None
        self.tau = tau
        self.iterations = 2
        self.constant = smagorinsky_constant


class Force:
    """Body force acting through BGKCollision(tau, force=...) (lettuce/ext/_force/_force.py).  Parameter
    holder: `acceleration` in lattice units, one component per dimension."""

    def __init__(self, flow, tau, acceleration):
        self.flow = flow
        self.tau = tau
        self.acceleration = flow.context.convert_to_tensor(acceleration)

    @property
    def ueq_scaling_factor(self):
        raise NotImplementedError


class Guo(Force):
    """Guo forcing: equilibrium velocity shifted by a/(2 rho), source term
    (1 - 1/(2 tau)) w_q [(e_q - u)/cs^2 + (e_q.u) e_q/cs^4] . a  (lettuce/ext/_force/guo.py:9-38)"""

    @property
    def ueq_scaling_factor(self):
        return 0.5


class ShanChen(Force):
    """Shan-Chen forcing: equilibrium velocity shifted by tau a / rho, no source term
    (lettuce/ext/_force/shan_chen.py:9-26)"""

    @property
    def ueq_scaling_factor(self):
        return self.tau * 1


class KBCCollision2D(KBCCollision):
    """deprecated alias (lettuce/ext/_collision/kbc_collision.py:169-173)"""

    def __init__(self, tau: float = None):
        import warnings
        warnings.warn("KBCCollision2D is is deprecated! Use KBCCollision instead!")
        super().__init__()


class KBCCollision3D(KBCCollision):
    """deprecated alias (lettuce/ext/_collision/kbc_collision.py:176-180)"""

    def __init__(self, tau: float = None):
        import warnings
        warnings.warn("KBCCollision2D is is deprecated! Use KBCCollision instead!")
        super().__init__()
