"""CPU-only tests of the host side: API surface, stencil tables, unit conversion, initial
conditions and mask construction against the golden vectors, descriptor packing, the C ABI's
exported symbols, and that nothing steps on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, max_rel
import lettuce_b200 as lt
from lettuce_b200 import native
from oracle import lbm_oracle as lo

STENCILS = {"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}


def cpu(dtype=torch.float64):
    return lt.Context("cpu", dtype=dtype)


def test_stencils_match_oracle_tables():
    for name, cls in STENCILS.items():
        s, o = cls(), lo.stencil(name)
        assert np.array_equal(np.asarray(s.e), o["e"])
        assert np.allclose(np.asarray(s.w), o["w"], rtol=0, atol=0)
        assert list(s.opposite) == list(o["opposite"])
        assert (s.d, s.q) == (o["d"], o["q"])


def test_cuda_stencil_tables_match_host_tables():
    """parse the constexpr tables in csrc/lbm_core.cuh and compare with the Python ones"""
    src = open(os.path.join(ROOT, "lettuce_b200", "csrc", "lbm_core.cuh")).read()
    for name, cls in STENCILS.items():
        body = src[src.index(f"struct {name} "):]
        table = body[body.index("constexpr int t["):body.index("};") + 2]
        nums = [int(x) for x in re.findall(r"-?\d+", table[table.index("=") + 1:])]
        e = np.array(nums).reshape(-1, 3)
        ref = np.asarray(cls().e)
        if name == "D2Q9":      # internal axes (x, -, y)
            assert np.array_equal(e[:, [0, 2]], ref) and (e[:, 1] == 0).all()
        else:
            assert np.array_equal(e, ref)


def test_context_rules():
    c = lt.Context("cpu")
    assert c.dtype == torch.float32 and c.use_native is False
    with pytest.raises(AssertionError):
        lt.Context("cpu", use_native=True)
    with pytest.raises(AssertionError):
        lt.Context("cpu", dtype=torch.int32)
    assert c.convert_to_tensor(np.array([True, False])).dtype == torch.uint8
    assert c.convert_to_tensor([1, 2]).dtype == torch.float32


def test_units_match_oracle():
    u = lt.UnitConversion(reynolds_number=1600, mach_number=0.05, characteristic_length_lu=32 / (2 * np.pi))
    o = lo.Units(1600, 0.05, 32 / (2 * np.pi))
    assert u.relaxation_parameter_lu == o.tau
    assert u.convert_velocity_to_lu(0.7) == o.velocity_to_lu(0.7)
    assert u.convert_pressure_pu_to_density_lu(0.3) == o.pressure_pu_to_density_lu(0.3)
    assert u.convert_incompressible_energy_to_pu(2.0) == o.incompressible_energy_to_pu(2.0)
    assert u.convert_time_to_pu(10) == o.time_to_pu(10)
    for a, b in (("velocity", 0.3), ("time", 4.0), ("length", 2.0), ("pressure", 0.1), ("density", 1.1),
                 ("energy", 0.2), ("incompressible_energy", 0.2), ("acceleration", 0.4)):
        there = getattr(u, f"convert_{a}_to_lu")(b)
        assert getattr(u, f"convert_{a}_to_pu")(there) == pytest.approx(b, rel=1e-14)


@pytest.mark.parametrize("name", ["tgv2d_d2q9_bgk", "tgv3d_d3q19_bgk", "tgv3d_d3q27_kbc", "tgv3d_d3q27_trt"])
def test_tgv_initial_condition_matches_reference(name):
    g = load_golden(name)
    stencil, coll, steps, re_, ma = g["meta"]
    flow = lt.TaylorGreenVortex(cpu(), [int(r) for r in g["res"]], float(re_), float(ma), stencil=STENCILS[stencil]())
    assert max_rel(flow.f.numpy(), g["f0"]) < 1e-14
    assert flow.units.relaxation_parameter_lu == pytest.approx(float(g["tau"]), rel=1e-15)
    assert flow.f.is_contiguous()


class ObstacleEqOut(lt.Obstacle):
    @property
    def post_boundaries(self):
        x = self.grid[0]
        return [lt.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                         velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                lt.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
                lt.BounceBackBoundary(self.mask)]


def make_obstacle(cls, ctx, res, stencil):
    D = res[1] / 8
    flow = cls(ctx, list(res), reynolds_number=100, mach_number=0.05, domain_length_x=res[0] / D, stencil=stencil)
    g = flow.grid
    c = [0.25 * g[0].max()] + [0.5 * gi.max() for gi in g[1:]]
    flow.mask = sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2
    flow.initialize()
    return flow


@pytest.mark.parametrize("name,cls", [("cylinder_d2q9_bgk", ObstacleEqOut), ("sphere_d3q27_trt", ObstacleEqOut),
                                      ("obstacle2d_abb_bgk", lt.Obstacle), ("obstacle3d_abb_bgk", lt.Obstacle)])
def test_obstacle_state_and_masks_match_reference(name, cls):
    g = load_golden(name)
    stencil = str(g["meta"][0])
    flow = make_obstacle(cls, cpu(), [int(r) for r in g["res"]], STENCILS[stencil]())
    assert np.array_equal(flow.mask.numpy(), g["solid"].astype(bool))
    assert max_rel(flow.f.numpy(), g["f0"]) < 1e-14
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    assert sim.no_collision_mask.dtype == torch.uint8 and sim.no_streaming_mask.dtype == torch.uint8
    assert np.array_equal(sim.no_collision_mask.numpy(), g["ncm"])
    assert np.array_equal(sim.no_streaming_mask.numpy(), g["nsm"])
    assert sim.collision_index == 0 and len(sim.transformer) == 4


def test_checked_tensor_rules():
    """equilibrium_boundary_pu.py:23-69"""
    flow = lt.TaylorGreenVortex(cpu(), [6, 5], 10, 0.05, stencil=lt.D2Q9())
    ct = lambda t: lt.EquilibriumBoundaryPU.checked_tensor(t, flow.context, flow)
    assert list(ct(0.5).shape) == [1, 1, 1]
    assert list(ct([0.1, 0.2]).shape) == [2, 1, 1]
    assert list(ct(np.zeros((6, 5))).shape) == [1, 6, 5]
    assert list(ct(np.zeros((2, 6, 1))).shape) == [2, 6, 1]
    with pytest.raises(ValueError):
        ct(np.zeros((3, 6, 5)))
    with pytest.raises(ValueError):
        ct(np.zeros((2, 4, 5)))
    with pytest.raises(ValueError):
        ct(np.zeros((2, 2, 2, 2)))
    with pytest.raises(TypeError):
        ct("nope")


def test_outlet_direction_validation():
    flow = lt.TaylorGreenVortex(cpu(), [6, 5], 10, 0.05, stencil=lt.D2Q9())
    with pytest.raises(AssertionError):
        lt.EquilibriumOutletP([1, 1], flow)
    with pytest.raises(AssertionError):
        lt.EquilibriumOutletP([0, 0], flow)
    o = lt.EquilibriumOutletP([0, -1], flow, rho_outlet=1.2)
    assert list(o.velocities) == [4, 7, 8] and o.index == [slice(None), 0] and o.neighbor == [slice(None), 1]


def test_no_cpu_path():
    flow = lt.TaylorGreenVortex(cpu(), [8, 8], 10, 0.05, stencil=lt.D2Q9())
    sim = lt.Simulation(flow, lt.BGKCollision(0.8), [])
    with pytest.raises(RuntimeError, match="CUDA"):
        sim(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        flow.rho()
    with pytest.raises(RuntimeError, match="CUDA"):
        lt.IncompressibleKineticEnergy(flow)()


def test_unsupported_operators_raise():
    class MRTCollision(lt.Collision):
        pass

    with pytest.raises(NotImplementedError):
        native.op_kind(MRTCollision())
    with pytest.raises(NotImplementedError):
        lt.SmagorinskyCollision(0.6, force=object())
    with pytest.raises(NotImplementedError):
        native.op_kind(lt.BGKCollision(0.6, force=object()))          # unknown forcing scheme
    flow = lt.PoiseuilleFlow2D(cpu(), 9, 1.0, 0.02)
    assert native.op_kind(lt.BGKCollision(0.6, force=lt.Guo(flow, 0.6, [1e-5, 0]))) == native.OP_BGK_FORCED
    d = native.describe(lt.Simulation(flow, lt.BGKCollision(0.6, force=lt.ShanChen(flow, 0.7, [1e-5, 0])), []))
    assert d["ops"][0]["kind"] == native.OP_BGK_FORCED and d["ops"][1]["kind"] == native.OP_BOUNCE_BACK
    assert native.op_kind(lt.KBCCollision()) == native.OP_KBC

    class MyBB(lt.BounceBackBoundary):
        pass

    assert native.op_kind(MyBB(None)) == native.OP_BOUNCE_BACK
    assert MyBB(None).native_available() and not MRTCollision().native_available()


def test_streaming_strategy_bits():
    S = lt.StreamingStrategy
    assert [s.value for s in (S.NO_STREAMING, S.POST_STREAMING, S.PRE_STREAMING, S.DOUBLE_STREAMING)] == [0, 1, 2, 3]
    assert S.DOUBLE_STREAMING.pre_streaming() and S.DOUBLE_STREAMING.post_streaming()
    assert S.PRE_STREAMING.pre_streaming() and not S.PRE_STREAMING.post_streaming()


def test_batch_length_respects_reporters():
    flow = lt.TaylorGreenVortex(cpu(), [8, 8], 10, 0.05, stencil=lt.D2Q9())
    rep = lt.ObservableReporter(lt.Mass(flow), interval=7, out=None)
    sim = lt.Simulation(flow, lt.BGKCollision(0.8), [rep])
    flow.i = 0
    assert sim._batch_length(100) == 7
    flow.i = 5
    assert sim._batch_length(100) == 2
    assert sim._batch_length(1) == 1

    class EveryStep(lt.Reporter):
        def __call__(self, simulation):
            pass

    sim.reporter.append(EveryStep(3))
    assert sim._batch_length(100) == 1
    sim.reporter.pop()
    sim._collide_and_stream = lambda s: None          # user-installed step
    assert sim._batch_length(100) == 1


def test_abi_library_loads_and_exports_every_declared_symbol():
    L = native.lib()
    header = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    declared = set(re.findall(r"\b(lbm_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/lbm_b200.h but not exported"
    assert set(native.EXPORTS) == declared
    assert L.lbm_abi_version() == 5
    assert b"unsupported" in L.lbm_status_string(-2)


def test_ctypes_struct_layout_matches_header():
    """sizes computed from the C declaration: lbm_op = 4*4 + 2*8 + 2*8 + 3*8 + 4*8 = 104 bytes"""
    assert ctypes.sizeof(native.LbmOp) == 104 + 5 * 8
    assert ctypes.sizeof(native.LbmLattice) == 24
    assert ctypes.sizeof(native.LbmHalo) == 12 * 8
    assert ctypes.sizeof(native.LbmStepDesc) == 24 + 16 + 8 * 144 + 16 + 16 + 96
    assert ctypes.sizeof(native.LbmLinks) == 8 + 8 + 6 * 8


def test_link_boundary_validation_without_gpu():
    """lbm_apply_links checks its arguments before any CUDA call"""
    L = native.lib()
    d = native.LbmStepDesc()
    d.lat = native.LbmLattice(native.D2Q9, native.F64, 8, 8, 1, 0)
    d.streaming, d.n_ops, d.collision_index = 1, 1, 0
    d.ops[0].kind, d.ops[0].p0 = native.OP_BGK, 0.6
    links = native.LbmLinks()
    links.kind, links.n = 1, 4
    links.node, links.q, links.bounced = 1 << 20, 1 << 21, 1 << 22
    pre, post = 1 << 24, 1 << 26
    assert L.lbm_links_scratch_doubles(0) == 0 and L.lbm_links_scratch_doubles(129) > 3 * 2
    assert L.lbm_apply_links(ctypes.byref(d), None, pre, post, None) == -1
    links.kind = 7
    assert L.lbm_apply_links(ctypes.byref(d), ctypes.byref(links), pre, post, None) == -1        # unknown kind
    links.kind = 2
    assert L.lbm_apply_links(ctypes.byref(d), ctypes.byref(links), pre, post, None) == -1        # interpolated without d
    links.kind = 1
    links.force = 1 << 23
    assert L.lbm_apply_links(ctypes.byref(d), ctypes.byref(links), pre, post, None) == -1        # force without scratch
    links.force = None
    d.streaming = 2
    assert L.lbm_apply_links(ctypes.byref(d), ctypes.byref(links), pre, post, None) == -2        # needs POST_STREAMING
    d.streaming = 1
    d.labels = 1 << 27
    assert L.lbm_apply_links(ctypes.byref(d), ctypes.byref(links), pre, post, None) == -1        # labels without frozen


def test_descriptor_validation_without_gpu():
    """argument validation happens before any CUDA call, so it is testable here"""
    L = native.lib()
    d = native.LbmStepDesc()
    d.lat = native.LbmLattice(native.D3Q19, native.F32, 8, 8, 8, 0)
    d.streaming, d.n_ops, d.collision_index = 1, 1, 0
    d.ops[0].kind, d.ops[0].p0 = native.OP_KBC, 0.6
    assert L.lbm_step(ctypes.byref(d), 1 << 20, 1 << 30, None) == -2          # KBC on D3Q19
    d.ops[0].kind = native.OP_BGK
    assert L.lbm_step(ctypes.byref(d), None, 1 << 30, None) == -1            # NULL buffer
    assert L.lbm_step(ctypes.byref(d), 1 << 20, (1 << 20) + 64, None) == -4  # overlap
    d.lat.nx = 0
    assert L.lbm_step(ctypes.byref(d), 1 << 20, 1 << 30, None) == -1
    d.lat = native.LbmLattice(native.D2Q9, native.F32, 8, 8, 2, 0)
    assert L.lbm_step(ctypes.byref(d), 1 << 20, 1 << 30, None) == -1          # D2Q9 needs nz = 1
    d.lat = native.LbmLattice(native.D3Q27, native.F32, 2048, 2048, 1024, 0)
    assert L.lbm_step(ctypes.byref(d), 1 << 20, 1 << 50, None) == -5          # >= 2^31 nodes
    d.lat = native.LbmLattice(native.D3Q27, native.F32, 8, 8, 8, 0)
    # outlets on up to three different axes are evaluated (nested neighbour look-ups in the sparse kernel); a fourth
    # active outlet is refused
    d.n_ops = 5
    for i, (axis, side) in enumerate([(0, 1), (1, 1), (2, -1), (0, -1)], start=1):
        d.ops[i].kind, d.ops[i].axis, d.ops[i].side = native.OP_OUTLET_P, axis, side
    d.labels, d.frozen = 1 << 20, 1 << 21
    assert L.lbm_step(ctypes.byref(d), 1 << 22, 1 << 30, None) == -2


def test_equilibrium_boundary_parameters_are_re_read_when_modified():
    """the reference converts the inlet velocity / pressure on every step (cuda_native/ext/_boundary/
    equilibrium_pu.py:15-18,55-58); the engine re-converts them when the tensors change"""
    ctx = cpu()
    boundaries = []

    class F(lt.TaylorGreenVortex):
        @property
        def post_boundaries(self):
            if not boundaries:
                m = torch.zeros(self.resolution, dtype=torch.bool); m[0] = True
                boundaries.append(lt.EquilibriumBoundaryPU(self.context, self, m, velocity=[0.5, 0.0], pressure=0.1))
            return boundaries

    flow = F(ctx, [8, 8], 10.0, 0.05, stencil=lt.D2Q9())
    sim = lt.Simulation(flow, lt.NoCollision(), [])
    eng = native.Engine(sim, dry=True)
    rho, u = eng._eq_state[1][:2]
    u_before, ptr = u.clone(), eng.desc.ops[1].u
    eng.refresh_parameters()
    assert torch.equal(u, u_before)
    boundaries[0].velocity[0] = 1.0                       # in-place edit bumps the version counter
    eng.refresh_parameters()
    assert torch.allclose(u.flatten()[0], flow.units.convert_velocity_to_lu(torch.tensor(1.0, dtype=u.dtype)))
    assert eng.desc.ops[1].u == ptr                       # same storage, descriptor still valid
    boundaries[0].pressure = boundaries[0].pressure * 2   # replaced tensor
    eng.refresh_parameters()
    assert torch.allclose(rho.flatten()[0], flow.units.convert_pressure_pu_to_density_lu(torch.tensor(0.2, dtype=rho.dtype)))


def test_write_vtk_round_trip(tmp_path):
    """the .vtr writer (replacement of pyevtk.hl.gridToVTK, lettuce/ext/_reporter/vtk_reporter.py:10-15): file
    name, header, x-fastest point order, appended blocks with 8-byte sizes -- read back bit for bit"""
    rng = np.random.default_rng(3)
    p = rng.random((5, 4, 3))
    ux = rng.random((5, 4, 3)).astype(np.float32)
    path = lt.write_vtk({"p": p, "ux": ux}, id=7, filename_base=str(tmp_path / "out"))
    assert path.endswith("out_00000007.vtr") and os.path.isfile(path)
    back = lt.read_vtr(path)
    assert np.array_equal(back["p"], p) and back["p"].dtype == np.float64
    assert np.array_equal(back["ux"], ux) and back["ux"].dtype == np.float32
    assert np.array_equal(back["x_coordinates"], np.arange(5)) and np.array_equal(back["z_coordinates"], np.arange(3))
    raw = open(path, "rb").read()
    assert raw.startswith(b'<?xml version="1.0"?>\n<VTKFile type="RectilinearGrid"')
    assert b'WholeExtent="0 4 0 3 0 2"' in raw and raw.rstrip().endswith(b"</VTKFile>")
    # the first appended block is p with x fastest: size header, then p[0,0,0], p[1,0,0]
    start = raw.index(b'<AppendedData encoding="raw">\n_') + len(b'<AppendedData encoding="raw">\n_')
    assert np.frombuffer(raw[start:start + 8], "<u8")[0] == p.nbytes
    assert np.array_equal(np.frombuffer(raw[start + 8:start + 24], "<f8"), p[:2, 0, 0])
    # 2-D fields get a trailing axis of one node, as in the reference's reporter (vtk_reporter.py:33-40)
    path2 = lt.write_vtk({"p": p[:, :, 0]}, id=1, filename_base=str(tmp_path / "flat"))
    assert lt.read_vtr(path2)["p"].shape == (5, 4, 1)
    with pytest.raises(ValueError):
        lt.write_vtk({"p": p, "ux": ux[:4]}, id=2, filename_base=str(tmp_path / "bad"))


def test_remaining_flows_match_reference():
    """Lamb-Oseen vortex, decaying turbulence (incl. the pressure-Poisson start in 2-D), Couette masks and the
    flow_by_name registry against the reference's initial states (tests/golden/more_flows.npz)"""
    g = load_golden("more_flows")
    ctx = cpu()
    assert max_rel(lt.LambOseenVortex2D(ctx, [48, 40], 100, 0.05).f.numpy(), g["lamb_f0"]) < 1e-14
    decay = lt.DecayingTurbulence(ctx, [32, 32], 1000, 0.05, k0=4, randseed=3)
    assert max_rel(decay.f.numpy(), g["decay2d_f0"]) < 1e-12
    assert np.allclose(decay.energy_spectrum[0], g["decay2d_spectrum"], rtol=1e-12, atol=1e-300)
    assert decay.units.relaxation_parameter_lu == pytest.approx(float(g["decay2d_tau"]), rel=1e-13)
    assert max_rel(lt.DecayingTurbulence(ctx, [12, 12, 12], 1000, 0.05, k0=3, randseed=5).f.numpy(),
                   g["decay3d_f0"]) < 1e-12
    couette = lt.Simulation(lt.CouetteFlow2D(ctx, [16, 12], 100, 0.05), lt.BGKCollision(0.6), [])
    assert np.array_equal(couette.no_collision_mask.numpy(), g["couette_ncm"])
    assert np.array_equal(np.isnan(couette.flow.f.numpy()), g["couette_f0_nan"])
    assert sorted(lt.flow_by_name) == ["couette2d", "decay2d", "lamboseen", "poiseuille2d", "shear2d", "taylor2d",
                                       "taylor3d_d3q19", "taylor3d_d3q27"]
    for name, (flow_class, stencil) in lt.flow_by_name.items():
        assert issubclass(flow_class, lt.ExtFlow) and stencil().q in (9, 19, 27)


def test_progress_reporter_logs_and_stays_batchable(tmp_path):
    """ProgressReporter (lettuce/ext/_reporter/progress_reporter.py): header on the first call, one line per due
    step; it reads no device data, so the step loop may still batch"""
    class FakeFlow:
        i = 0

    class FakeSim:
        flow = FakeFlow()

    rep = lt.ProgressReporter(interval=5, i_target=20, outdir=tmp_path / "log", t_max=1e9)
    sim = FakeSim()
    for i in (0, 5, 7, 10):
        sim.flow.i = i
        rep(sim)
    lines = open(tmp_path / "log" / "progress_reporter_log.txt").read().splitlines()
    assert lines[0].startswith("t_start:") and "t_per_step" in lines[1]
    assert len(lines) == 4 and lines[2].split()[1] == "5" and lines[3].split()[1] == "10"
    assert "WARNING" not in lines[2]
    assert rep.batchable


def test_observable_reporter_defers_device_values_until_read():
    """values that live on the device are fetched in one transfer when `reporter.out` is read; rows, order and
    number formats are those of the eager path (observable_reporter.py:161-200)"""
    class FakeUnits:
        @staticmethod
        def convert_time_to_pu(i):
            return 0.5 * i

    class FakeFlow:
        i = 0
        f = None

    class FakeSim:
        flow = FakeFlow()
        units = FakeUnits()

    class Scalar:
        context = cpu(torch.float32)

        def __call__(self, f=None):
            return torch.tensor(1.0 + FakeSim.flow.i, dtype=torch.float32)

    class Vector(Scalar):
        def __call__(self, f=None):
            return torch.arange(3, dtype=torch.float64) * FakeSim.flow.i

    for obs, rows_at_4 in ((Scalar(), [[0, 0.0, 1.0], [2, 1.0, 3.0], [4, 2.0, 5.0]]),
                           (Vector(), [[0, 0.0, 0.0, 0.0, 0.0], [2, 1.0, 0.0, 2.0, 4.0], [4, 2.0, 0.0, 4.0, 8.0]])):
        lazy = lt.ObservableReporter(obs, interval=2, out=None, defer=True)
        eager = lt.ObservableReporter(obs, interval=2, out=None, defer=False)
        sim = FakeSim()
        for i in range(5):
            sim.flow.i = i
            lazy(sim); eager(sim)
            if i == 2:
                assert lazy.out == eager.out and len(lazy._pending) == 0       # reading mid-run fetches what is there
        assert len(lazy._pending) == 1 and len(lazy._rows[-1]) == 2
        held = lazy._rows                                     # a list a user obtained from `.out` earlier ...
        sim_like = lt.Simulation.__new__(lt.Simulation)
        sim_like.reporter = [lazy, eager]
        sim_like._flush_reporters()                           # ... is complete once Simulation.__call__ returns
        assert held == rows_at_4 and not lazy._pending
        assert lazy.out == eager.out == rows_at_4
        assert all(isinstance(v, float) for row in lazy.out for v in row[1:])
        lazy.out = []
        assert lazy.out == []
    import io
    stream = io.StringIO()
    printed = lt.ObservableReporter(Scalar(), interval=1, out=stream, defer=True)     # streams are always eager
    FakeSim.flow.i = 3
    printed(FakeSim())
    assert stream.getvalue().split() == ["3", "1.5", "4.0"] and printed.out is stream


def test_observable_reporter_takes_whole_batches_of_fused_moments():
    """`Simulation.__call__` hands reporters with interval 1 the step kernels' own reductions for a whole run at once
    (`lbm_step_moments_n`): same rows as step-by-step reports, mixed freely with single deferred values"""
    class FakeUnits:
        @staticmethod
        def convert_time_to_pu(i):
            return 0.25 * i

    class FakeFlow:
        i = 0
        f = torch.zeros(1)                       # a CPU tensor: batches are refused, see below

    class FakeSim:
        flow = FakeFlow()
        units = FakeUnits()

    class Energy:
        context = cpu(torch.float64)
        fused_with_step = True
        flow = FakeSim.flow

        def __call__(self, f=None):
            return torch.tensor(10.0 * FakeSim.flow.i, dtype=torch.float64)

        def from_fused_moments(self, moments):
            return 10.0 * moments[:, 0]

    sim = FakeSim()
    rep = lt.ObservableReporter(Energy(), interval=1, out=None, defer=True)
    assert not rep.accepts_fused_batches(sim)                # CPU populations: no engine, no batches
    sim.flow.i = 0
    rep(sim)                                                 # step 0, single deferred value
    moments = torch.tensor([[1.0, 0.5], [2.0, 0.5], [3.0, 0.5]], dtype=torch.float64)
    rep.ingest_fused_batch(sim, 1, moments)                  # steps 1..3 in one go
    sim.flow.i = 4
    rep(sim)                                                 # step 4, single again
    assert len(rep._pending) == 3 and [len(r) for r in rep._rows] == [2] * 5
    assert rep.out == [[0, 0.0, 0.0], [1, 0.25, 10.0], [2, 0.5, 20.0], [3, 0.75, 30.0], [4, 1.0, 40.0]]
    assert not rep._pending
    # reporters that print, report less often, or evaluate anything else keep the step-by-step path
    import io
    assert not lt.ObservableReporter(Energy(), interval=1, out=io.StringIO()).accepts_fused_batches(sim)
    assert not lt.ObservableReporter(Energy(), interval=2, out=None).accepts_fused_batches(sim)


def test_engine_rejects_other_equilibria():
    """the kernels evaluate the quadratic equilibrium; a flow built with another one must not run silently"""
    class Other(lt.Equilibrium):
        def __call__(self, flow, rho=None, u=None):
            return lt.QuadraticEquilibrium()(flow, rho, u)

    ctx = cpu()
    flow = lt.TaylorGreenVortex(ctx, [8, 8], 10, 0.05, stencil=lt.D2Q9(), equilibrium=Other())
    sim = lt.Simulation(flow, lt.BGKCollision(0.6), [])
    with pytest.raises(NotImplementedError):
        native.describe(sim)
    same = lt.TaylorGreenVortex(ctx, [8, 8], 10, 0.05, stencil=lt.D2Q9(), equilibrium=lt.QuadraticEquilibriumLessMemory())
    assert native.describe(lt.Simulation(same, lt.BGKCollision(0.6), []))["ops"][0]["kind"] == native.OP_BGK
    assert torch.equal(same.f, lt.TaylorGreenVortex(ctx, [8, 8], 10, 0.05, stencil=lt.D2Q9()).f)


def test_util_helpers_and_deprecated_aliases():
    """torch_gradient (orders 2/4/6) and torch_jacobi on analytic periodic fields, append_axes, the exception /
    warning classes and the deprecated aliases user scripts still import (lettuce/util/utility.py)"""
    n = 64
    x = torch.linspace(0, 2 * np.pi * (1 - 1 / n), n, dtype=torch.float64)
    X, Y = torch.meshgrid(x, x, indexing="ij")
    field = torch.sin(X) * torch.cos(2 * Y)
    dx = 2 * np.pi / n
    errors = []
    for order in (2, 4, 6):
        g = lt.torch_gradient(field, dx=dx, order=order)
        assert g.shape == (2, n, n)
        errors.append(max(float((g[0] - torch.cos(X) * torch.cos(2 * Y)).abs().max()),
                          float((g[1] + 2 * torch.sin(X) * torch.sin(2 * Y)).abs().max())))
    assert errors[0] > 10 * errors[1] > 100 * errors[2] and errors[2] < 1e-6
    # the order-6 weights are the ones the enstrophy kernel and initialize_f_neq use (oracle.gradient6)
    assert np.allclose(lt.torch_gradient(field, dx=dx, order=6).numpy(), lo.gradient6(field.numpy(), dx), atol=1e-13)
    with pytest.raises(lt.LettuceException):
        lt.torch_gradient(torch.zeros(4), 1)
    rhs = -5 * field                                        # lap(field) = -(1 + 4) field
    p = lt.torch_jacobi(rhs, torch.zeros_like(field), dx, dim=2, tol_abs=1e-12, max_num_steps=20000)
    assert float((p - p.mean() - field).abs().max()) < 5e-3          # 2nd-order discrete Laplacian
    assert lt.append_axes(np.ones(3), 2).shape == (3, 1, 1)
    assert issubclass(lt.InefficientCodeWarning, lt.LettuceWarning) and issubclass(lt.LettuceWarning, UserWarning)
    ctx = cpu()
    with pytest.warns(DeprecationWarning):
        flow = lt.TaylorGreenVortex3D(ctx, [8, 8, 8], 100, 0.05, stencil=lt.D3Q19())
    assert type(flow) is lt.TaylorGreenVortex
    with pytest.warns(UserWarning):
        kbc = lt.KBCCollision2D()
    assert native.op_kind(kbc) == native.OP_KBC


def test_header_is_plain_c_and_matches_ctypes_sizes(tmp_path):
    """include/lbm_b200.h compiles as strict C99 (the drop-in boundary has no C++ or torch types), and a C
    program's sizeof of every struct equals the ctypes mirror in native.py"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lbm_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(lbm_op), sizeof(lbm_lattice), '
                   'sizeof(lbm_halo), sizeof(lbm_step_desc), sizeof(lbm_slab), sizeof(lbm_links));\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", offsetof(lbm_op, rho_stride), offsetof(lbm_op, force), '
                   'offsetof(lbm_step_desc, ops), offsetof(lbm_step_desc, labels), offsetof(lbm_step_desc, n_general), '
                   'offsetof(lbm_step_desc, halo), offsetof(lbm_links, force)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True, capture_output=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    sizes, offsets = ([int(v) for v in line.split()] for line in lines)
    mirrors = [native.LbmOp, native.LbmLattice, native.LbmHalo, native.LbmStepDesc, native.LbmSlab, native.LbmLinks]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]
    assert offsets == [native.LbmOp.rho_stride.offset, native.LbmOp.force.offset, native.LbmStepDesc.ops.offset,
                       native.LbmStepDesc.labels.offset, native.LbmStepDesc.n_general.offset,
                       native.LbmStepDesc.halo.offset, native.LbmLinks.force.offset]


def test_flow_diagnostics():
    """shear tensor, H-theorem entropy and the deprecated einsum helper (lettuce/_flow.py:206-256); the
    reference's own `entropy` expression raises a shape error in its current API, ours evaluates the formula"""
    flow = lt.TaylorGreenVortex(cpu(), [10, 12], 100, 0.05, stencil=lt.D2Q9())
    f = flow.f.numpy()
    st = lo.stencil("D2Q9")
    e = st["e"].astype(float)
    want = np.einsum("qxy,qa,qb->abxy", f, e, e)
    assert np.allclose(flow.shear_tensor().numpy(), want, atol=1e-15)
    w = st["w"].reshape(-1, 1, 1)
    assert np.allclose(flow.entropy().numpy(), (f * -np.log((f / w).sum(axis=0))).sum(axis=0), atol=1e-14)
    with pytest.warns(DeprecationWarning):
        assert torch.allclose(flow.einsum("q,q->", [flow.f, flow.f]), (flow.f * flow.f).sum(dim=0))


def test_hdf5_reporter_with_a_stand_in_h5py(tmp_path, monkeypatch):
    """HDF5Reporter (lettuce/util/datautils.py:17-80): file layout and append logic, exercised with an in-memory
    stand-in for h5py (the real module is not installed here)"""
    import sys
    import types
    store = {}

    class Dataset:
        def __init__(self, shape, dtype):
            self.data = np.zeros(shape, dtype=dtype)

        shape = property(lambda self: self.data.shape)

        def resize(self, n, axis=0):
            new = np.zeros((n, *self.data.shape[1:]), dtype=self.data.dtype)
            new[:self.data.shape[0]] = self.data
            self.data = new

        def __setitem__(self, key, value):
            self.data[key] = value

    class File:
        def __init__(self, name, mode):
            if mode == "w":
                store[name] = dict(attrs={}, sets={})
            self.attrs, self.sets = store[name]["attrs"], store[name]["sets"]

        def create_dataset(self, name, shape, maxshape, dtype):
            self.sets[name] = Dataset(shape, dtype)

        def __getitem__(self, name):
            return self.sets[name]

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    fake = types.ModuleType("h5py")
    fake.File = File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    flow = lt.TaylorGreenVortex(cpu(torch.float32), [6, 8], 100, 0.05, stencil=lt.D2Q9())
    collision = lt.BGKCollision(0.7)
    rep = lt.HDF5Reporter(flow, collision, interval=2, filebase=str(tmp_path / "out"), metadata={"note": "x"})

    class Sim:
        pass

    sim = Sim()
    sim.flow = flow
    first = flow.f.clone()
    for i in range(5):
        flow.i = i
        flow.f = first * (1 + i)
        rep(sim)
    rec = store[str(tmp_path / "out") + ".h5"]
    assert rec["sets"]["f"].shape == (3, 9, 6, 8) and rec["sets"]["f"].data.dtype == np.float32
    assert np.array_equal(rec["sets"]["f"].data[2], (first * 5).numpy())
    assert rec["attrs"]["data"] == "3" and rec["attrs"]["steps"] == "4" and rec["attrs"]["note"] == "x"
    import pickle
    assert pickle.loads(bytes(rec["attrs"]["_collision"]))["cls"] == "BGKCollision"


def test_pre_boundary_masks_follow_the_reference_quirk():
    """`Flow.pre_boundaries`: collision_index = 1 and BOTH masks start from collision_index
    (lettuce/_simulation.py:104-107) -- the no-streaming mask is 1 everywhere."""
    g = load_golden("pre_boundary")
    solid = g["D2Q9_solid"]
    mask = torch.tensor(solid)

    class PreFlow(lt.TaylorGreenVortex):
        @property
        def pre_boundaries(self):
            return [lt.BounceBackBoundary(mask)]

    flow = PreFlow(cpu(), list(solid.shape), 100.0, 0.05, stencil=lt.D2Q9())
    assert max_rel(flow.f.numpy(), g["D2Q9_f0"]) < 1e-14
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    assert sim.collision_index == 1 and len(sim.transformer) == 2
    assert bool((sim.no_streaming_mask == 1).all())
    assert np.array_equal(sim.no_collision_mask.numpy(), np.where(solid, 0, 1).astype(np.uint8))
