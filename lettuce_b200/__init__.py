"""lettuce_b200 -- B200-native stream+collide engine behind lettuce's Python API.

    import lettuce_b200 as lt
    ctx = lt.Context("cuda", dtype=torch.float32)
    flow = lt.TaylorGreenVortex(ctx, [256] * 3, 1600, 0.05, stencil=lt.D3Q19())
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    mlups = sim(100)

Only the hot path of lettuce is provided: Context, stencils D2Q9/D3Q19/D3Q27,
UnitConversion, Flow/TaylorGreenVortex/Obstacle, BGK/TRT/KBC/NoCollision, BounceBack/
EquilibriumBoundaryPU/EquilibriumOutletP/AntiBounceBackOutlet, Simulation, the moment observables.
Every time step and every moment reduction runs in hand-written sm_100a CUDA kernels
(lettuce_b200/csrc) reached through the C ABI of include/lbm_b200.h; there is no CPU fallback.
"""
from ._context import *
from ._stencil import *
from .units import *
from ._flow import *
from ._simulation import *
from .ext import *
from .util import *
from . import native

__version__ = "0.1.0"
