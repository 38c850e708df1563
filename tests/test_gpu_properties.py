"""Property tests of the CUDA operators, modelled on the reference's own unit tests
(tests/collision/*.py, tests/test_equilibrium.py, tests/boundary/*.py) but run through the B200 engine."""
import numpy as np
import pytest
import torch

from conftest import max_rel

pytestmark = pytest.mark.gpu

lt = pytest.importorskip("lettuce_b200")
from oracle import lbm_oracle as lo  # noqa: E402

STENCILS = {"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}
COLLISIONS = ["bgk", "trt", "kbc", "regularized", "smagorinsky"]


def ctx(dtype=torch.float64):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return lt.Context("cuda", dtype=dtype)


class TestFlow(lt.ExtFlow):
    """uniform u = 1.01, p = 0.01 (pu): the reference's TestFlow (tests/conftest.py:195-232)"""
    __test__ = False
    boundary_factory = None

    def make_resolution(self, resolution, stencil=None):
        return resolution

    def make_units(self, reynolds_number, mach_number, resolution):
        return lt.UnitConversion(reynolds_number, mach_number, characteristic_length_lu=resolution[0])

    def initial_pu(self):
        return 0.01 * np.ones([1] + self.resolution), 1.01 * np.ones([self.stencil.d] + self.resolution)

    @property
    def post_boundaries(self):
        return self.boundary_factory(self) if self.boundary_factory else []


def random_flow(stencil, res, dtype=torch.float64, seed=1, amplitude=0.2, cls=TestFlow, re=100.0, ma=0.1):
    flow = cls(ctx(dtype), res, re, ma, stencil=STENCILS[stencil]())
    st = lo.stencil(stencil)
    rng = np.random.default_rng(seed)
    f0 = st["w"].reshape((-1,) + (1,) * st["d"]) * (1.0 + amplitude * (rng.random((st["q"], *res)) - 0.5))
    flow.f = flow.context.convert_to_tensor(f0).contiguous()
    return flow, st, f0


def collide_once(flow, coll, tau):
    c = {"bgk": lambda: lt.BGKCollision(tau), "trt": lambda: lt.TRTCollision(tau), "kbc": lambda: lt.KBCCollision(),
         "regularized": lambda: lt.RegularizedCollision(), "smagorinsky": lambda: lt.SmagorinskyCollision(tau),
         "none": lambda: lt.NoCollision()}[coll]()
    lt.Simulation(flow, c, [], lt.StreamingStrategy.NO_STREAMING)(1)
    return flow.f.cpu().numpy()


@pytest.mark.parametrize("stencil", list(STENCILS))
@pytest.mark.parametrize("coll", COLLISIONS)
def test_collision_conserves_mass_and_momentum(stencil, coll):
    """tests/collision/test_collision_conserves_{mass,momentum}.py"""
    if coll == "kbc" and stencil == "D3Q19":
        pytest.skip("KBC exists for D2Q9 and D3Q27 only")
    flow, st, f0 = random_flow(stencil, [7] * STENCILS[stencil]().d)
    f1 = collide_once(flow, coll, 0.6 if coll not in ("kbc", "regularized") else flow.units.relaxation_parameter_lu)
    assert np.max(np.abs(lo.rho(f1) - lo.rho(f0))) < 1e-14
    assert np.max(np.abs(lo.j(st, f1) - lo.j(st, f0))) < 1e-14
    assert np.max(np.abs(f1 - f0)) > 1e-6          # the operator did act


@pytest.mark.parametrize("stencil", list(STENCILS))
@pytest.mark.parametrize("coll", COLLISIONS)
def test_collision_relaxes_shear_moments(stencil, coll):
    """tests/collision/test_collision_relaxes_shear_moments.py, on a NON-equilibrium state: every second
    moment relaxes towards its equilibrium value at rate 1/tau (BGK, TRT's even part, KBC's shear part)"""
    if coll == "kbc" and stencil == "D3Q19":
        pytest.skip("KBC exists for D2Q9 and D3Q27 only")
    if coll == "smagorinsky":
        pytest.skip("relaxes with the local effective tau, not the prescribed one")
    flow, st, f0 = random_flow(stencil, [6] * STENCILS[stencil]().d)
    tau = 0.6 if coll not in ("kbc", "regularized") else flow.units.relaxation_parameter_lu
    f1 = collide_once(flow, coll, tau)
    e = st["e"].astype(float)
    shear = lambda f: np.einsum("q...,qa,qb->ab...", f, e, e)
    feq = lo.equilibrium(st, lo.rho(f0), lo.u(st, f0))
    expect = shear(f0) - (1.0 / tau) * (shear(f0) - shear(feq))
    assert np.max(np.abs(shear(f1) - expect)) < 1e-13


@pytest.mark.parametrize("stencil", list(STENCILS))
def test_equilibrium_conserves_mass_and_momentum(stencil):
    """tests/test_equilibrium.py:4-39 on the device-side equilibrium kernel (lbm_equilibrium)"""
    c = ctx()
    s = STENCILS[stencil]()
    res = [5, 6, 7][:s.d]
    rng = np.random.default_rng(2)
    rho = c.convert_to_tensor(1.0 + 0.1 * rng.random([1, *res]))
    u = c.convert_to_tensor(0.05 * rng.standard_normal([s.d, *res]))
    f = lt.native.equilibrium_field(s, rho, u, res)
    r, v = lt.native.moments(s, f)
    assert max_rel(r.cpu().numpy(), rho.cpu().numpy()) < 1e-14
    assert np.max(np.abs(v.cpu().numpy() - u.cpu().numpy())) < 1e-15
    st = lo.stencil(stencil)
    assert max_rel(f.cpu().numpy(), lo.equilibrium(st, rho.cpu().numpy()[0], u.cpu().numpy())) < 1e-14


def test_kbc_increases_pseudo_entropy_over_bgk():
    """tests/collision/test_collision_optimizes_pseudo_entropy.py: same seeded random populations"""
    for stencil in ("D2Q9", "D3Q27"):
        d = STENCILS[stencil]().d
        np.random.seed(1)
        f0 = np.random.random([STENCILS[stencil]().q] + [3] * d)
        st = lo.stencil(stencil)
        outs = {}
        for coll in ("kbc", "bgk"):
            flow = TestFlow(ctx(), [3] * d, 100.0, 0.1, stencil=STENCILS[stencil]())
            flow.f = flow.context.convert_to_tensor(f0).contiguous()
            tau = flow.units.relaxation_parameter_lu     # KBC takes tau from the units (kbc_collision.py:97-99)
            outs[coll] = collide_once(flow, coll, tau)
        feq = lo.equilibrium(st, lo.rho(f0), lo.u(st, f0))
        entropy = lambda f: lo.rho(f) - (f * f / feq).sum(axis=0)      # Flow.pseudo_entropy_local, _flow.py:222-228
        assert (entropy(outs["bgk"]) < entropy(outs["kbc"])).all()


@pytest.mark.parametrize("stencil", list(STENCILS))
def test_bounce_back_everywhere_and_nowhere(stencil):
    """tests/boundary/test_bounceback_bc.py:6-36"""
    d = STENCILS[stencil]().d
    for everywhere in (True, False):
        class F(TestFlow):
            boundary_factory = staticmethod(
                lambda self: [lt.BounceBackBoundary(torch.full(self.resolution, everywhere, dtype=torch.bool))])
        flow, st, f0 = random_flow(stencil, [5] * d, cls=F)
        f1 = collide_once(flow, "none", 1.0)
        assert np.array_equal(f1, f0[st["opposite"]] if everywhere else f0)


@pytest.mark.parametrize("stencil", list(STENCILS))
@pytest.mark.parametrize("sign", [1, -1])
def test_equilibrium_outlet_p_plane(stencil, sign):
    """tests/boundary/test_equilibrium_bc_outlet_p.py:6-31: the outlet plane equals feq(rho_outlet, u of the
    neighbour plane); here exactly (the reference accepts rel 1e-2), for both ends of the last axis"""
    d = STENCILS[stencil]().d
    direction = [0] * (d - 1) + [sign]

    class F(TestFlow):
        boundary_factory = staticmethod(lambda self: [lt.EquilibriumOutletP(direction, self, rho_outlet=1.2)])

    flow = F(ctx(), [8] * d, 1.0, 0.1, stencil=STENCILS[stencil]())
    f0 = flow.f.cpu().numpy()
    f1 = collide_once(flow, "none", 1.0)
    st = lo.stencil(stencil)
    u_lu = np.full([d] + [8] * (d - 1), flow.units.convert_velocity_to_lu(1.01))
    expect = lo.equilibrium(st, 1.2, u_lu)
    plane = -1 if sign > 0 else 0
    assert max_rel(f1[..., plane], expect) < 1e-13
    keep = [i for i in range(8) if i != plane % 8]
    assert np.array_equal(f1[..., keep], f0[..., keep])       # NoCollision elsewhere


def test_anti_bounce_back_outlet_formula():
    """tests/boundary/test_antibounceback_outlet_bc.py: textbook formula (Krueger et al., p. 195) on the plane"""
    class F(TestFlow):
        boundary_factory = staticmethod(lambda self: [lt.AntiBounceBackOutlet([1, 0], self)])

    flow, st, f0 = random_flow("D2Q9", [6, 7], cls=F, amplitude=0.05)
    f1 = collide_once(flow, "none", 1.0)
    u = lo.u(st, f0)
    u_w = u[:, -1] + 0.5 * (u[:, -1] - u[:, -2])
    rho_w = lo.rho(f0)[-1]
    expect = f0.copy()
    for q in np.flatnonzero(st["e"][:, 0] == 1):
        eu = st["e"][q] @ u_w
        expect[st["opposite"][q], -1] = (-f0[q, -1] + st["w"][q] * rho_w *
                                         (2 + eu ** 2 / lo.CS2 ** 2 - (u_w ** 2).sum(axis=0) / lo.CS2))
    assert max_rel(f1, expect) < 1e-13


def test_masks_nonempty_for_obstacle():
    """tests/boundary/test_bc_masks.py:4-19"""
    c = ctx(torch.float32)
    flow = lt.Obstacle(c, [32, 16], 100, 0.05, domain_length_x=16, stencil=lt.D2Q9())
    m = torch.zeros([32, 16], dtype=torch.bool); m[10:14, 6:10] = True
    flow.mask = m
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    assert sim.no_collision_mask.any() and sim.no_streaming_mask.any()
    sim(2)
    assert torch.isfinite(flow.f).all()


def test_reporters_change_slowly():
    """tests/reporter/test_generic_reporters.py:4-25: observables move < 5 % over two steps"""
    c = ctx(torch.float32)
    flow = lt.TaylorGreenVortex(c, [32] * 3, 1600.0, 0.05, stencil=lt.D3Q27())
    reps = [lt.ObservableReporter(o(flow), interval=1, out=None)
            for o in (lt.IncompressibleKineticEnergy, lt.Enstrophy, lt.MaximumVelocity, lt.Mass)]
    lt.Simulation(flow, lt.KBCCollision(), reps)(2)
    for r in reps:
        vals = np.array([row[2] for row in r.out])
        assert len(vals) == 3 and np.all(np.abs(vals / vals[0] - 1) < 0.05)


def test_checkpoint_round_trip(tmp_path):
    """tests/test_checkpoint.py: dump, keep stepping, load restores the dumped state"""
    c = ctx(torch.float64)
    flow = lt.TaylorGreenVortex(c, [16, 16], 10.0, 0.05, stencil=lt.D2Q9())
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    sim(3)
    saved = flow.f.clone()
    flow.dump(tmp_path / "f.pkl")
    sim(3)
    assert not torch.equal(flow.f, saved)
    flow.load(tmp_path / "f.pkl")
    assert torch.equal(flow.f, saved)
    sim(1)                                  # f_next is re-allocated lazily after load
    assert torch.isfinite(flow.f).all()


def test_high_ma_reporter_aborts_at_the_reference_iteration(tmp_path):
    """tests/reporter/test_high_ma_reporter.py:5-17: the stock obstacle at Ma 0.2 exceeds Ma 0.3 locally at
    iteration 13 (same iteration on the reference's torch path in fp32 and fp64)"""
    for dtype in (torch.float32, torch.float64):
        c = ctx(dtype)
        flow = lt.Obstacle(context=c, resolution=[16, 16], reynolds_number=10, mach_number=0.2, domain_length_x=16,
                           stencil=lt.D2Q9())
        g = flow.grid
        flow.mask = ((2 < g[0]) & (g[0] < 10) & (2 < g[1]) & (g[1] < 10))
        reporter = lt.HighMaReporter(1, outdir=str(tmp_path))
        sim = lt.BreakableSimulation(flow, lt.BGKCollision(tau=flow.units.relaxation_parameter_lu), [reporter])
        sim(100)
        assert flow.i > 100 and reporter.failed_iteration == 13
        assert (tmp_path / "HighMa_reporter.log").is_file() and len(reporter.results) > 0


def test_nan_reporter_aborts(tmp_path):
    """tests/reporter/test_nan_reporter.py: a NaN planted in the populations stops a BreakableSimulation"""
    c = ctx(torch.float32)
    flow = lt.TaylorGreenVortex(c, [16, 16], 10.0, 0.05, stencil=lt.D2Q9())
    reporter = lt.NaNReporter(2, outdir=str(tmp_path))
    sim = lt.BreakableSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [reporter])
    sim(4)
    assert reporter.failed_iteration is None and flow.i == 4
    flow.f[3, 5, 7] = float("nan")
    sim(50)
    assert reporter.failed_iteration == 6 and flow.i > 50      # next due step after the NaN was planted
    assert len(reporter.results) > 0 and (tmp_path / "NaN_reporter.log").is_file()


def test_convergence_orders():
    """`lettuce convergence` (lettuce/cli.py:134-186, run by the reference's CI): second order in u, first in p"""
    from lettuce_b200.cli import run_convergence
    order_u, order_p = run_convergence(ctx(torch.float64), echo=lambda *_: None)
    assert 1.9 < order_u < 2.1 and 0.9 < order_p < 1.1


@pytest.mark.parametrize("stencil,res,coll,dtype", [("D3Q19", [24, 20, 36], "bgk", torch.float32),
                                                    ("D3Q27", [12, 16, 20], "kbc", torch.float64),
                                                    ("D3Q27", [12, 16, 20], "kbc", torch.float32),
                                                    ("D2Q9", [40, 33], "trt", torch.float64),
                                                    # 4200 CTAs: the partials take the two-stage fold
                                                    ("D2Q9", [4200, 40], "bgk", torch.float64),
                                                    # large enough for the TMA-staged kernel (its consumers reduce)
                                                    ("D3Q27", [16, 16, 320], "kbc", torch.float32)])
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "NO_STREAMING", "POST_STREAMING"])
def test_fused_step_moments_equal_the_reductions(stencil, res, coll, dtype, strategy):
    """lbm_step_moments: same populations as lbm_step, and (sum 0.5|u|^2, max |u|^2) equal the stand-alone
    reductions of the state they describe -- the new state (NO / PRE streaming) or the step's input (POST)"""
    from lettuce_b200 import native as nv
    c = ctx(dtype)
    mk = lambda: lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=STENCILS[stencil]())
    make = lambda fl: {"bgk": lt.BGKCollision(fl.units.relaxation_parameter_lu), "kbc": lt.KBCCollision(),
                       "trt": lt.TRTCollision(fl.units.relaxation_parameter_lu)}[coll]
    fa, fb = mk(), mk()
    sa = lt.Simulation(fa, make(fa), [], lt.StreamingStrategy[strategy])
    sb = lt.Simulation(fb, make(fb), [], lt.StreamingStrategy[strategy])
    eng = nv.engine_of(sb)
    assert eng.moments_state() == (nv.MOMENTS_OF_INPUT if strategy == "POST_STREAMING" else nv.MOMENTS_OF_OUTPUT)
    rel = 1e-12 if dtype == torch.float64 else 1e-6
    for _ in range(3):
        before = fa.f.clone()
        nv.invoke(sa)
        got = eng.step_with_moments().cpu().tolist()
        assert torch.equal(fa.f, fb.f)
        described = before if strategy == "POST_STREAMING" else fa.f
        assert got[0] == pytest.approx(float(nv.reduce(fa.stencil, nv.SUM_HALF_U2, described).cpu()), rel=rel)
        assert got[1] ** 0.5 == pytest.approx(float(nv.reduce(fa.stencil, nv.MAX_U, described).cpu()), rel=rel)
    double = lt.Simulation(mk(), lt.NoCollision(), [], lt.StreamingStrategy.DOUBLE_STREAMING)   # not available
    assert nv.engine_of(double).moments_state() == nv.MOMENTS_UNAVAILABLE
    with pytest.raises(RuntimeError):
        nv.engine_of(double).step_with_moments()


@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_step_moments_with_boundaries(strategy, dtype):
    """masked runs: the bulk kernel leaves the general nodes out of its partial sums and the sparse kernel adds
    them (inlet, pressure outlet, bounce-back cylinder)"""
    from lettuce_b200 import native as nv
    from test_gpu_parity import ObstacleEqOut, make_obstacle
    c = ctx(dtype)
    for stencil, res in ((lt.D2Q9, [96, 32]), (lt.D3Q27, [48, 24, 24])):
        fa, fb = (make_obstacle(ObstacleEqOut, c, res, stencil()) for _ in range(2))
        mk = lambda fl: lt.Simulation(fl, lt.BGKCollision(fl.units.relaxation_parameter_lu), [],
                                      lt.StreamingStrategy[strategy])
        sa, sb = mk(fa), mk(fb)
        eng = nv.engine_of(sb)
        assert int(eng.desc.n_general) > 0
        rel = 1e-12 if dtype == torch.float64 else 1e-6
        for _ in range(4):
            before = fa.f.clone()
            nv.invoke(sa)
            got = eng.step_with_moments().cpu().tolist()
            assert torch.equal(fa.f, fb.f)
            described = before if strategy == "POST_STREAMING" else fa.f
            assert got[0] == pytest.approx(float(nv.reduce(fa.stencil, nv.SUM_HALF_U2, described).cpu()), rel=rel)
            assert got[1] ** 0.5 == pytest.approx(float(nv.reduce(fa.stencil, nv.MAX_U, described).cpu()), rel=rel)


@pytest.mark.parametrize("interval", [1, 3])
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_moment_reporters_ride_on_the_step_kernels(interval, strategy, dtype):
    """Simulation.__call__ lets the step kernels reduce due IncompressibleKineticEnergy / MaximumVelocity reports
    (PRE: the step that writes the state; POST: the step that follows it): same populations and same reporter
    values as a step-by-step run with the stand-alone reductions, and no stand-alone reduction is launched except
    for the report at step 0 and, with POST streaming, the last one"""
    from lettuce_b200 import native as nv
    c = ctx(dtype)
    res, steps = [40, 24, 36], 6
    mk = lambda: lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=lt.D3Q19())
    fa, fb = mk(), mk()
    rep = lt.ObservableReporter(lt.IncompressibleKineticEnergy(fa), interval=interval, out=None)
    rep_u = lt.ObservableReporter(lt.MaximumVelocity(fa), interval=interval, out=None)
    strat = lt.StreamingStrategy[strategy]
    sa = lt.Simulation(fa, lt.BGKCollision(fa.units.relaxation_parameter_lu), [rep, rep_u], strat)
    sb = lt.Simulation(fb, lt.BGKCollision(fb.units.relaxation_parameter_lu), [], strat)
    nv.engine_of(sa)
    before = nv.launch_count()
    sa(steps)
    launched = nv.launch_count() - before
    energy_b, umax_b = lt.IncompressibleKineticEnergy(fb), lt.MaximumVelocity(fb)
    expect = [[0, float(energy_b().cpu()), float(umax_b().cpu())]]
    for i in range(1, steps + 1):
        nv.invoke(sb)
        if i % interval == 0:
            expect.append([i, float(energy_b().cpu()), float(umax_b().cpu())])
    assert torch.equal(fa.f, fb.f)
    assert fa.i == steps
    assert [e[0] for e in rep.out] == [e[0] for e in expect] == [e[0] for e in rep_u.out]
    rel = 1e-12 if dtype == torch.float64 else 1e-6
    for got, got_u, want in zip(rep.out, rep_u.out, expect):
        assert got[2] == pytest.approx(want[1], rel=rel)
        assert got_u[2] == pytest.approx(want[2], rel=rel)
    # stand-alone reductions (2 launches each: reduce + fold) for both observables at step 0 and, with POST streaming,
    # at the last step; `steps` step kernels; one fold per due report that rode on a step kernel (tiny lattice: one
    # stage)
    reports = steps // interval
    fused = reports if strategy == "PRE_STREAMING" else reports - 1
    standalone = 1 if strategy == "PRE_STREAMING" else 2
    assert launched == 4 * standalone + steps + fused
    if strategy == "PRE_STREAMING":
        # the cached values are dropped as soon as the populations move on or are written through torch
        assert nv.fused_moments(fa, fa.f) is not None
        fa.f.mul_(1.0)
        assert nv.fused_moments(fa, fa.f) is None
        sa(interval)
        nv.invoke(sa)
    assert nv.fused_moments(fa, fa.f) is None


@pytest.mark.parametrize("stencil,res,coll", [("D3Q27", [12, 10, 64], "kbc"), ("D3Q19", [10, 12, 96], "bgk"),
                                              ("D2Q9", [40, 130], "trt"), ("D3Q27", [8, 12, 66], "smagorinsky"),
                                              ("D2Q9", [36, 64], "kbc"), ("D3Q19", [8, 8, 64], "regularized")])
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_two_nodes_per_thread_kernel_is_bit_identical(stencil, res, coll, strategy):
    """the packed float2 kernel (two neighbouring nodes per thread, FADD2 / FMUL2 / FFMA2) against the one-node
    kernel: same bits, because every operation is an explicit round-to-nearest intrinsic in both (lbm_vec.cuh)"""
    from lettuce_b200 import native as nv
    from test_gpu_parity import make_collision
    c = ctx(torch.float32)
    flows = []
    for lanes in (1, 2):
        flow = lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=STENCILS[stencil]())
        gen = torch.Generator(device=flow.f.device).manual_seed(5)
        flow.f.mul_(1.0 + 1e-2 * (torch.rand(flow.f.shape, generator=gen, device=flow.f.device) - 0.5))
        sim = lt.Simulation(flow, make_collision(coll, flow), [], lt.StreamingStrategy[strategy])
        eng = nv.engine_of(sim)
        eng.desc.variant = lanes
        assert ("2 nodes" in eng.lib.lbm_step_variant_name(eng.desc).decode()) == (lanes == 2)
        nv.invoke_n(sim, 7)
        flows.append(flow)
    assert torch.equal(flows[0].f, flows[1].f)


@pytest.mark.parametrize("stencil,res,coll", [("D3Q27", [12, 10, 640], "kbc"), ("D3Q19", [37, 9, 256], "bgk"),
                                              ("D2Q9", [300, 320], "trt"), ("D3Q27", [24, 24, 192], "trt"),
                                              ("D2Q9", [2048, 64], "kbc"), ("D3Q19", [5, 70, 512], "regularized"),
                                              ("D3Q27", [16, 16, 320], "kbc"), ("D3Q19", [12, 32, 256], "bgk")])
@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "NO_STREAMING", "POST_STREAMING"])
def test_tma_staged_kernel_is_bit_identical(stencil, res, coll, strategy):
    """the TMA-staged persistent kernel (csrc/lbm_tma.cuh: bulk tensor loads of rows shifted in x and y, halo quads
    for the shift along z, partial last tile, several z chunks per row) against the one-node LDG kernel: same bits,
    it only moves data differently"""
    from lettuce_b200 import native as nv
    from test_gpu_parity import make_collision
    c = ctx(torch.float32)
    flows = []
    for variant in (1, 3):
        flow = lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=STENCILS[stencil]())
        gen = torch.Generator(device=flow.f.device).manual_seed(11)
        flow.f.mul_(1.0 + 1e-2 * (torch.rand(flow.f.shape, generator=gen, device=flow.f.device) - 0.5))
        sim = lt.Simulation(flow, make_collision(coll, flow), [], lt.StreamingStrategy[strategy])
        eng = nv.engine_of(sim)
        eng.desc.variant = variant
        assert ("TMA" in eng.lib.lbm_step_variant_name(eng.desc).decode()) == (variant == 3)
        nv.invoke_n(sim, 5)
        flows.append(flow)
    assert torch.equal(flows[0].f, flows[1].f)


@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_tma_staged_kernel_with_boundaries(strategy):
    from lettuce_b200 import native as nv
    from test_gpu_parity import ObstacleEqOut, make_obstacle
    c = ctx(torch.float32)
    for stencil, res, coll in ((lt.D2Q9, [1200, 64], "bgk"), (lt.D3Q27, [48, 32, 64], "trt")):
        out = []
        for variant in (1, 3):
            flow = make_obstacle(ObstacleEqOut, c, res, stencil())
            tau = flow.units.relaxation_parameter_lu
            sim = lt.Simulation(flow, lt.BGKCollision(tau) if coll == "bgk" else lt.TRTCollision(tau), [],
                                lt.StreamingStrategy[strategy])
            eng = nv.engine_of(sim)
            eng.desc.variant = variant
            assert ("TMA" in eng.lib.lbm_step_variant_name(eng.desc).decode()) == (variant == 3)
            nv.invoke_n(sim, 9)
            out.append(flow.f)
        assert torch.equal(out[0], out[1])


def test_tma_staged_kernel_refuses_what_it_cannot_run():
    """variant 3 is an explicit request: a lattice it cannot stage (contiguous extent not a multiple of 64) or a step
    that both pulls and pushes is an error, never a silent change of kernel"""
    from lettuce_b200 import native as nv
    c = ctx(torch.float32)
    for res, strategy in (([64, 48, 100], "PRE_STREAMING"), ([64, 48, 128], "DOUBLE_STREAMING")):
        flow = lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=lt.D3Q19())
        sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [],
                            lt.StreamingStrategy[strategy])
        eng = nv.engine_of(sim)
        eng.desc.variant = 3
        with pytest.raises(RuntimeError):
            nv.invoke_n(sim, 1)
        eng.desc.variant = 0
        nv.invoke_n(sim, 1)


@pytest.mark.parametrize("strategy", ["PRE_STREAMING", "POST_STREAMING"])
def test_two_nodes_per_thread_kernel_with_boundaries(strategy):
    from lettuce_b200 import native as nv
    from test_gpu_parity import ObstacleEqOut, make_obstacle
    c = ctx(torch.float32)
    for stencil, res, coll in ((lt.D2Q9, [96, 32], "bgk"), (lt.D3Q27, [48, 24, 24], "trt")):
        out = []
        for lanes in (1, 2):
            flow = make_obstacle(ObstacleEqOut, c, res, stencil())
            tau = flow.units.relaxation_parameter_lu
            sim = lt.Simulation(flow, lt.BGKCollision(tau) if coll == "bgk" else lt.TRTCollision(tau), [],
                                lt.StreamingStrategy[strategy])
            nv.engine_of(sim).desc.variant = lanes
            nv.invoke_n(sim, 9)
            out.append(flow.f)
        assert torch.equal(out[0], out[1])


def test_vtk_reporter_writes_engine_fields(tmp_path):
    """VTKReporter (tests/reporter/test_vtk_reporter_no_mask.py, test_vtk_reporter_mask.py): one file per due
    step, written in the background, holding the engine's pressure and velocity in physical units"""
    c = ctx(torch.float32)
    flow = lt.TaylorGreenVortex(c, [16, 24], 10.0, 0.05, stencil=lt.D2Q9())
    rep = lt.VTKReporter(interval=2, filename_base=str(tmp_path / "data" / "output"))
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [rep])
    sim(4)
    rep.wait()
    names = sorted(p.name for p in (tmp_path / "data").iterdir())
    assert names == ["output_00000000.vtr", "output_00000002.vtr", "output_00000004.vtr"]
    back = lt.read_vtr(tmp_path / "data" / "output_00000004.vtr")
    assert back["p"].shape == (16, 24, 1) and back["ux"].dtype == np.float32
    assert np.array_equal(back["p"][..., 0], flow.p_pu[0].cpu().numpy())
    assert np.array_equal(back["uy"][..., 0], flow.u_pu[1].cpu().numpy())
    with pytest.raises(ValueError):
        rep.output_mask(sim)
    obstacle = lt.Obstacle(c, [24, 12, 12], 100.0, 0.05, 4.0, stencil=lt.D3Q19())
    obstacle.mask = (obstacle.grid[0] - 1.0) ** 2 + (obstacle.grid[1] - 1.0) ** 2 < 0.3
    sim3 = lt.Simulation(obstacle, lt.BGKCollision(obstacle.units.relaxation_parameter_lu), [])
    mask_file = lt.VTKReporter(1, str(tmp_path / "m" / "o")).output_mask(sim3)
    assert np.array_equal(lt.read_vtr(mask_file)["mask"], sim3.no_collision_mask.cpu().numpy())


@pytest.mark.parametrize("stencil,res,strategy", [("D2Q9", [64, 48], "PRE_STREAMING"), ("D3Q19", [16, 12, 20], "POST_STREAMING")])
def test_graph_replay_of_small_lattices_is_bit_identical(stencil, res, strategy, monkeypatch):
    """LBM_B200_GRAPH_MAX_NODES: lbm_step_n replays 32-step CUDA graphs on small lattices; same populations as
    step-by-step launches, same launch count, odd remainders and changed parameters handled"""
    from lettuce_b200 import native as nv
    c = ctx(torch.float64)
    mk = lambda: lt.TaylorGreenVortex(c, res, 100.0, 0.05, stencil=STENCILS[stencil]())
    fa, fb = mk(), mk()
    sa = lt.Simulation(fa, lt.BGKCollision(fa.units.relaxation_parameter_lu), [], lt.StreamingStrategy[strategy])
    sb = lt.Simulation(fb, lt.BGKCollision(fb.units.relaxation_parameter_lu), [], lt.StreamingStrategy[strategy])
    for _ in range(71):
        nv.invoke(sb)
    monkeypatch.setenv("LBM_B200_GRAPH_MAX_NODES", "100000")
    before = nv.launch_count()
    nv.invoke_n(sa, 71)                       # two graphs of 32 steps + 7 plain steps
    assert nv.launch_count() - before == 71
    assert torch.equal(fa.f, fb.f)
    nv.invoke_n(sa, 33)                       # odd batch on the swapped buffer pair: a second cached graph
    sa.collision.tau = 0.8                    # parameters are part of the cache key: a third graph
    nv.invoke_n(sa, 40)
    monkeypatch.delenv("LBM_B200_GRAPH_MAX_NODES")
    nv.invoke_n(sb, 33)
    sb.collision.tau = 0.8
    nv.invoke_n(sb, 40)
    assert torch.equal(fa.f, fb.f)
    monkeypatch.setenv("LBM_B200_GRAPH_MAX_NODES", "10")        # lattice larger than the limit: plain launches
    nv.invoke_n(sa, 32)
    monkeypatch.delenv("LBM_B200_GRAPH_MAX_NODES")
    nv.invoke_n(sb, 32)
    assert torch.equal(fa.f, fb.f)
