"""INTEGRATION.md section 2 in executable form: `lettuce_b200.native` translates the REFERENCE's own
`lettuce.Simulation` objects (duck typing on class and attribute names) into the C-ABI descriptor.
Runs only where the reference tree is mounted (the build container); skipped on the GPU box."""
import hashlib
import os
import sys
import types

import pytest
import torch

from lettuce_b200 import native

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "lettuce")),
                                reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    stubs = {}
    m = types.ModuleType("mmh3"); m.hash_bytes = lambda v: hashlib.md5(v.encode() if isinstance(v, str) else v).digest()
    hl = types.ModuleType("pyevtk.hl"); hl.gridToVTK = lambda *a, **k: None
    pe = types.ModuleType("pyevtk"); pe.hl = hl
    h5 = types.ModuleType("h5py"); h5.File = None
    stubs.update({"mmh3": m, "pyevtk": pe, "pyevtk.hl": hl, "h5py": h5})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    sys.path.insert(0, REFERENCE)
    try:
        import lettuce
        yield lettuce
    finally:
        sys.path.remove(REFERENCE)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_tgv_simulation_translates(ref):
    ctx = ref.Context(device="cpu", dtype=torch.float32, use_native=False)
    flow = ref.TaylorGreenVortex(ctx, [8, 8, 8], 1600.0, 0.05, stencil=ref.D3Q19())
    sim = ref.Simulation(flow, ref.BGKCollision(flow.units.relaxation_parameter_lu), [],
                         ref.StreamingStrategy.PRE_STREAMING)
    d = native.describe(sim)
    assert d["stencil"] == native.D3Q19 and d["dtype"] == native.F32 and d["resolution"] == [8, 8, 8]
    assert d["streaming"] == 2 and d["collision_index"] == 0
    assert d["ops"][0]["kind"] == native.OP_BGK and d["ops"][0]["p0"] == pytest.approx(flow.units.relaxation_parameter_lu)


def test_reference_obstacle_simulation_translates(ref):
    ctx = ref.Context(device="cpu", dtype=torch.float64, use_native=False)

    class ObstacleEqOut(ref.Obstacle):
        @property
        def post_boundaries(self):
            x = self.grid[0]
            return [ref.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                              velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                    ref.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
                    ref.BounceBackBoundary(self.mask)]

    flow = ObstacleEqOut(ctx, [32, 8, 8], reynolds_number=100, mach_number=0.05, domain_length_x=32, stencil=ref.D3Q27())
    sim = ref.Simulation(flow, ref.TRTCollision(0.6, 0.9), [])
    d = native.describe(sim)
    kinds = [o["kind"] for o in d["ops"]]
    assert kinds == [native.OP_TRT, native.OP_EQUILIBRIUM, native.OP_OUTLET_P, native.OP_BOUNCE_BACK]
    assert (d["ops"][0]["p0"], d["ops"][0]["p1"]) == (0.6, 0.9)
    assert (d["ops"][2]["axis"], d["ops"][2]["side"], d["ops"][2]["p0"]) == (0, 1, 1.0)
    assert d["ops"][1]["u_stride"][0] == 1 and d["ops"][1]["rho_stride"] == [0, 0, 0]     # broadcast inlet values
    # stock Obstacle: anti-bounce-back outlet; KBC picks up tau from the units like the reference does on first call
    flow2 = ref.Obstacle(ctx, [32, 8], reynolds_number=100, mach_number=0.05, domain_length_x=32, stencil=ref.D2Q9())
    coll = ref.KBCCollision(123.0)
    d2 = native.describe(ref.Simulation(flow2, coll, []))
    assert [o["kind"] for o in d2["ops"]] == [native.OP_KBC, native.OP_EQUILIBRIUM, native.OP_ANTI_BOUNCE_BACK,
                                             native.OP_BOUNCE_BACK]
    assert d2["ops"][0]["p0"] == pytest.approx(flow2.units.relaxation_parameter_lu) and coll.tau == d2["ops"][0]["p0"]


def test_reference_unsupported_operator_raises(ref):
    ctx = ref.Context(device="cpu", dtype=torch.float64, use_native=False)
    flow = ref.TaylorGreenVortex(ctx, [8, 8], 10.0, 0.05, stencil=ref.D2Q9())
    sim = ref.Simulation(flow, ref.MRTCollision(ref.D2Q9Dellar(ref.D2Q9(), ctx), [0.6] * 9) if hasattr(ref, "D2Q9Dellar")
                         else ref.NoCollision(), [])
    if type(sim.collision).__name__ == "MRTCollision":
        with pytest.raises(NotImplementedError):
            native.describe(sim)
