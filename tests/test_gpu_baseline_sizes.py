"""GPU parity at BASELINE.json's own lattice sizes, the drop-in proof with the reference's own objects, and the
operator-call contract.

* 512^3 D3Q19 fp32: q*N = 2.5e9 >= 2^31 elements -- plane offsets must be 64-bit (reference
  streaming lettuce/_simulation.py:241-256).  Pure streaming is compared BIT-EXACTLY with per-population torch.roll.
* configs[1] (TGV3D D3Q19 BGK 256^3 fp32), configs[3] (cylinder D2Q9 BGK 4096x1024 with inlet + EquilibriumOutletP +
  bounce-back) and a configs[4]-shaped sphere (D3Q27 TRT 256x128x128): 10 steps against the NumPy oracle, fp32 <= 1e-5.
* `reference.Simulation(...)._collide_and_stream = native.invoke` on CUDA against the same
  reference objects stepped by the reference's own torch path on the CPU.
* `collision(flow)` / `boundary(flow)` (lettuce/_simulation.py:17-28, lettuce/_flow.py:31-52), checked the way the
  reference's tests/collision/*.py and tests/boundary/*.py do.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, max_rel

pytestmark = pytest.mark.gpu

lt = pytest.importorskip("lettuce_b200")
from lettuce_b200 import native  # noqa: E402
from oracle import lbm_oracle as lo  # noqa: E402

STENCILS = {"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}
STRATS = {s.name: s for s in lt.StreamingStrategy}


def cuda_ctx(dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return lt.Context("cuda", dtype=dtype)


def set_f(flow, f0):
    flow.f = flow.context.convert_to_tensor(np.ascontiguousarray(f0), dtype=flow.context.dtype).contiguous()


def get_f(flow):
    return flow.f.detach().cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------ 64-bit addressing
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_streaming_at_512_cubed_is_bit_exact(strategy):
    """19 * 512^3 = 2.55e9 elements per buffer: every population plane beyond q = 15 starts above 2^31."""
    ctx = cuda_ctx(torch.float32)
    free, _ = torch.cuda.mem_get_info()
    if free < 45e9:
        pytest.skip("needs 45 GB of device memory")
    n, steps = 512, 3
    flow = lt.TaylorGreenVortex(ctx, [n] * 3, 1600.0, 0.05, stencil=lt.D3Q19())
    assert flow.f.numel() >= 2 ** 31
    # a distinct value in every slot: a wrong plane offset or a wrapped 32-bit index cannot go unnoticed
    gen = torch.Generator(device=flow.f.device).manual_seed(3)
    flow.f.uniform_(0.5, 1.5, generator=gen)
    f0 = flow.f.clone()
    sim = lt.Simulation(flow, lt.NoCollision(), [], STRATS[strategy])
    sim(steps)
    e = flow.stencil.e
    for q in range(flow.stencil.q):
        want = torch.roll(f0[q], shifts=tuple(steps * int(c) for c in e[q]), dims=(0, 1, 2))
        assert torch.equal(flow.f[q], want), (strategy, q)
        del want


# ------------------------------------------------------------------ BASELINE.json sizes against the oracle
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_config2_tgv3d_d3q19_bgk_256_cubed_matches_oracle(strategy):
    ctx = cuda_ctx(torch.float32)
    st = lo.stencil("D3Q19")
    n, steps = 256, 10
    flow = lt.TaylorGreenVortex(ctx, [n] * 3, 1600.0, 0.05, stencil=lt.D3Q19())
    f0 = flow.f.detach().cpu().numpy()                      # fp32 initial state, shared bit for bit
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], STRATS[strategy])
    sim(steps)
    from concurrent.futures import ThreadPoolExecutor
    import os
    cores = max(1, min(os.cpu_count() or 1, 32))
    ref = f0.astype(np.float64)
    coll = dict(kind="bgk", tau=flow.units.relaxation_parameter_lu)
    with ThreadPoolExecutor(cores) as pool:
        for _ in range(steps):
            ref = lo.step_parallel(st, ref, coll, strategy=strategy, pool=pool, chunks=4 * cores)
    got = flow.f.detach().cpu().numpy()
    err = float(np.max(np.abs(got - ref) / np.abs(ref)))
    assert err < 1e-5, (strategy, err)


@pytest.mark.parametrize("name,stencil,res,coll", [("config4_cylinder", "D2Q9", [4096, 1024], "bgk"),
                                                   ("config5_sphere", "D3Q27", [256, 128, 128], "trt")])
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_obstacle_configs_match_oracle(name, stencil, res, coll, strategy):
    from test_gpu_parity import ObstacleEqOut, make_obstacle, make_collision
    ctx = cuda_ctx(torch.float32)
    st = lo.stencil(stencil)
    flow = make_obstacle(ObstacleEqOut, ctx, res, STENCILS[stencil]())
    f0, units, post, solid = lo.obstacle_setup(st, res)
    assert np.array_equal(flow.mask.cpu().numpy(), solid)
    f0 = f0.astype(np.float32).astype(np.float64)
    set_f(flow, f0)
    steps = 10
    sim = lt.Simulation(flow, make_collision(coll, flow), [], STRATS[strategy])
    sim(steps)
    ref = lo.run(st, f0, steps, dict(kind=coll, tau=units.tau), post=post, strategy=strategy)
    err = max_rel(get_f(flow), ref)
    assert err < 1e-5, (name, strategy, err)


# ------------------------------------------------------------------ the reference's own objects on the engine
@pytest.fixture(scope="module")
def ref():
    from baseline import reference
    if not reference.available():
        pytest.skip("baseline/_ref is not installed")
    return reference.load()


def _reference_case(ref, case, device, dtype):
    ctx = ref.Context(device=device, dtype=dtype, use_native=False)
    if case == "tgv3d_d3q19_bgk_pre":
        flow = ref.TaylorGreenVortex(ctx, [20, 16, 24], 1600.0, 0.05, stencil=ref.D3Q19())
        return flow, ref.Simulation(flow, ref.BGKCollision(flow.units.relaxation_parameter_lu), [],
                                    ref.StreamingStrategy.PRE_STREAMING)
    if case == "tgv3d_d3q27_kbc_post":
        flow = ref.TaylorGreenVortex(ctx, [12, 16, 20], 1600.0, 0.05, stencil=ref.D3Q27())
        return flow, ref.Simulation(flow, ref.KBCCollision(), [], ref.StreamingStrategy.POST_STREAMING)

    class ObstacleEqOut(ref.Obstacle):
        @property
        def post_boundaries(self):
            x = self.grid[0]
            return [ref.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                              velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                    ref.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
                    ref.BounceBackBoundary(self.mask)]

    class ObstacleCornerOutlets(ref.Obstacle):
        """outlets on DIFFERENT axes (their planes meet along an edge / in a corner): the pressure outlet's neighbour
        on the edge lies on the anti-bounce-back outlet's plane and vice versa"""
        @property
        def post_boundaries(self):
            x = self.grid[0]
            d = self.stencil.d
            unit = lambda a, s=1: [s if k == a else 0 for k in range(d)]
            out = [ref.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                             velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                   ref.AntiBounceBackOutlet(unit(1), self),
                   ref.EquilibriumOutletP(direction=unit(0), flow=self, rho_outlet=1.0)]
            if d == 3:
                out.append(ref.EquilibriumOutletP(direction=unit(2, -1), flow=self, rho_outlet=1.01))
            return out + [ref.BounceBackBoundary(self.mask)]

    if case == "sphere_d3q27_trt_post":
        cls, res, stencil = ObstacleEqOut, [32, 16, 16], ref.D3Q27()
    elif case == "corner_outlets_d2q9_bgk_post":
        cls, res, stencil = ObstacleCornerOutlets, [40, 24], ref.D2Q9()
    elif case == "corner_outlets_d3q19_trt_post":
        cls, res, stencil = ObstacleCornerOutlets, [24, 16, 12], ref.D3Q19()
    elif case == "cylinder_d2q9_bgk_post":
        cls, res, stencil = ObstacleEqOut, [64, 16], ref.D2Q9()
    else:                                               # stock lt.Obstacle: anti-bounce-back outlet
        cls, res, stencil = ref.Obstacle, [48, 16], ref.D2Q9()
    D = res[1] / 8
    flow = cls(ctx, list(res), reynolds_number=100, mach_number=0.05, domain_length_x=res[0] / D, stencil=stencil)
    g = flow.grid
    c = [0.25 * g[0].max()] + [0.5 * gi.max() for gi in g[1:]]
    flow.mask = sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2
    if device == "cpu":
        # (the reference's Obstacle.initial_pu mixes a CPU unit vector with the CUDA mask and raises on a CUDA
        # context, lettuce/ext/_flows/obstacle.py:94-99; the CUDA arm takes the CPU arm's initial populations)
        flow.initialize()
    tau = flow.units.relaxation_parameter_lu
    collision = ref.TRTCollision(tau) if "trt" in case else ref.BGKCollision(tau)
    return flow, ref.Simulation(flow, collision, [], ref.StreamingStrategy.POST_STREAMING)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", ["tgv3d_d3q19_bgk_pre", "tgv3d_d3q27_kbc_post", "sphere_d3q27_trt_post",
                                  "cylinder_d2q9_bgk_post", "stock_obstacle_d2q9_bgk_post",
                                  "corner_outlets_d2q9_bgk_post", "corner_outlets_d3q19_trt_post"])
def test_reference_simulation_steps_on_the_engine(ref, case, dtype):
    """INTEGRATION.md section 2, executed: the reference's Simulation / Flow / Collision / Boundary objects, unmodified,
    with `native.invoke` assigned where the reference installs its generated kernel (lettuce/_simulation.py:229)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    steps = 10
    flow_cpu, sim_cpu = _reference_case(ref, case, "cpu", dtype)
    flow_gpu, sim_gpu = _reference_case(ref, case, "cuda", dtype)
    flow_gpu.f = flow_cpu.f.to("cuda").contiguous()                 # identical initial state, bit for bit
    sim_gpu._collide_and_stream = native.invoke
    launches = native.launch_count()
    sim_cpu(steps)                                                  # the reference's torch path
    sim_gpu(steps)                                                  # the reference's step loop, our kernel
    assert native.launch_count() - launches >= steps
    assert flow_gpu.i == flow_cpu.i == steps
    err = max_rel(flow_gpu.f.cpu().numpy(), flow_cpu.f.numpy())
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    if "kbc" in case and dtype == torch.float32:
        tol = 5 * float(load_golden("kbc_fp32_floor")["tgv_D3Q27_POST_STREAMING"])
    assert err < tol, (case, dtype, err)


@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_engine_against_the_references_generated_kernel(ref, strategy):
    """The kernel this engine replaces (lettuce/cuda_native/_template.py:64-81), generated and compiled by the
    reference's own tools (baseline/build_native.py), and `native.invoke` installed in the same slot
    (lettuce/_simulation.py:229) of a second, identical reference Simulation: same populations after 10 steps.
    (The generated kernel is compiled with --use_fast_math: approximate division, fp32 only here.)"""
    from baseline import reference
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    strat = ref.StreamingStrategy[strategy]
    sims = []
    for use_native in (True, False):
        ctx = ref.Context(device="cuda", dtype=torch.float32, use_native=use_native)
        flow = ref.TaylorGreenVortex(ctx, [40, 24, 64], 1600.0, 0.05, stencil=ref.D3Q19())
        collision = ref.BGKCollision(flow.units.relaxation_parameter_lu)
        sim = (reference.native_simulation(flow, collision, [], strat) if use_native
               else ref.Simulation(flow, collision, [], strat))
        if sim is None:
            pytest.skip("the reference's generated module is not prebuilt (python baseline/build_native.py)")
        sims.append((flow, sim))
    (flow_ref, sim_ref), (flow_eng, sim_eng) = sims
    assert sim_ref._collide_and_stream.__module__.startswith("lettuce_")        # the generated package's invoke
    flow_eng.f = flow_ref.f.clone()
    sim_eng._collide_and_stream = native.invoke
    sim_ref(10)
    sim_eng(10)
    torch.cuda.synchronize()
    err = max_rel(flow_eng.f.cpu().numpy(), flow_ref.f.cpu().numpy())
    assert err < 2e-5, (strategy, err)


# ------------------------------------------------------------------ operators are callable
class RandomFlow(lt.ExtFlow):
    def make_resolution(self, resolution, stencil=None):
        return resolution

    def make_units(self, reynolds_number, mach_number, resolution):
        return lt.UnitConversion(reynolds_number=reynolds_number, mach_number=mach_number,
                                 characteristic_length_lu=resolution[0])

    def initial_pu(self):
        d = len(self.resolution)
        return np.zeros((1, *self.resolution)), np.zeros((d, *self.resolution))

    @property
    def post_boundaries(self):
        return []


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_collision_call_matches_reference_golden_and_conserves(dtype):
    """`collision(flow)` returns the post-collision populations and leaves flow.f alone; checked against the
    reference's outputs on random populations, plus the reference's own conservation tests
    (tests/collision/test_collision_conserves_mass.py, test_collision_conserves_momentum.py)."""
    g = load_golden("random_collisions")
    ctx = cuda_ctx(dtype)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    for stencil, res in (("D2Q9", [6, 5]), ("D3Q19", [4, 5, 6]), ("D3Q27", [4, 5, 6])):
        for coll in ("bgk", "trt", "kbc", "regularized", "smagorinsky"):
            if coll == "kbc" and stencil == "D3Q19":
                continue
            flow = RandomFlow(ctx, res, 50.0, 0.1, stencil=STENCILS[stencil]())
            set_f(flow, g[f"{stencil}_f0"])
            tau = flow.units.relaxation_parameter_lu
            collision = {"bgk": lambda: lt.BGKCollision(tau), "trt": lambda: lt.TRTCollision(tau, 0.8),
                         "kbc": lambda: lt.KBCCollision(), "regularized": lambda: lt.RegularizedCollision(),
                         "smagorinsky": lambda: lt.SmagorinskyCollision(tau, 0.17)}[coll]()
            before = flow.f.clone()
            out = collision(flow)
            assert torch.equal(flow.f, before) and out.data_ptr() != flow.f.data_ptr()
            assert max_rel(out.cpu().numpy(), g[f"{stencil}_{coll}"]) < tol, (stencil, coll)
            ctol = 1e-12 if dtype == torch.float64 else 2e-6
            assert torch.allclose(flow.rho(out), flow.rho(before), rtol=ctol, atol=ctol)
            assert torch.allclose(flow.j(out), flow.j(before), rtol=ctol, atol=ctol)
            assert max_rel(collision(flow).cpu().numpy(), out.cpu().numpy()) == 0.0       # cached engine, same answer


def test_boundary_call_contract():
    """tests/boundary/test_bounceback_bc.py, test_equilibrium_bc_pu.py, test_equilibrium_bc_outlet_p.py in spirit:
    `boundary(flow)` acts on the whole lattice; Simulation blends by label afterwards."""
    ctx = cuda_ctx(torch.float64)
    rng = np.random.default_rng(9)
    for stencil, res in (("D2Q9", [7, 6]), ("D3Q27", [5, 6, 4])):
        st = lo.stencil(stencil)
        d = st["d"]
        flow = RandomFlow(ctx, res, 50.0, 0.1, stencil=STENCILS[stencil]())
        f0 = st["w"].reshape((-1,) + (1,) * d) * (1 + 0.1 * rng.random((st["q"], *res)))
        set_f(flow, f0)
        # bounce-back: f[opposite] everywhere (bounce_back_boundary.py:17-18), bit-exact
        bb = lt.BounceBackBoundary(torch.zeros(res, dtype=torch.bool))
        out = bb(flow)
        assert torch.equal(out, flow.f[list(flow.stencil.opposite)])
        assert torch.equal(flow.f.cpu(), torch.as_tensor(f0))
        # equilibrium boundary: feq(p, u) broadcast to the lattice (equilibrium_boundary_pu.py:79-84)
        vel = 0.05 * rng.standard_normal(d)
        eq = lt.EquilibriumBoundaryPU(ctx, flow, torch.zeros(res, dtype=torch.bool), vel, np.array(0.01))
        out = eq(flow).cpu().numpy()
        units = lo.Units(50.0, 0.1, characteristic_length_lu=res[0])
        rho = units.pressure_pu_to_density_lu(np.full([1] + [1] * d, 0.01))
        u = units.velocity_to_lu(vel.reshape([d] + [1] * d))
        want = np.broadcast_to(lo.equilibrium(st, rho, u), out.shape)
        assert max_rel(out, want) < 1e-13
        # pressure outlet: rewrites its plane of flow.f IN PLACE and returns flow.f (equilibrium_outlet_p.py:63-73)
        direction = [1] + [0] * (d - 1)
        outlet = lt.EquilibriumOutletP(direction, flow, rho_outlet=1.02)
        rho_f, u_f = lo.rho(f0), lo.u(st, f0)
        res_t = outlet(flow)
        assert res_t is flow.f
        got = flow.f.cpu().numpy()
        want = f0.copy()
        want[:, -1] = lo.equilibrium(st, np.full_like(rho_f[-2], 1.02), u_f[:, -2])
        assert max_rel(got, want) < 1e-13
