"""Generate the golden vectors under tests/golden/ from the REFERENCE's own torch CPU path.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference with three stub
modules for third-party imports that are not on the arithmetic path (mmh3,
pyevtk, h5py).  Every case stores the exact input
populations and the reference's output after N steps, so the oracle
(oracle/lbm_oracle.py) and the CUDA path can both be pinned against it without
the reference being present.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    m = types.ModuleType("mmh3")
    m.hash_bytes = lambda v: hashlib.md5(v.encode() if isinstance(v, str) else v).digest()
    sys.modules["mmh3"] = m
    hl = types.ModuleType("pyevtk.hl"); hl.gridToVTK = lambda *a, **k: None
    pe = types.ModuleType("pyevtk"); pe.hl = hl
    sys.modules["pyevtk"] = pe; sys.modules["pyevtk.hl"] = hl
    h5 = types.ModuleType("h5py"); h5.File = None
    sys.modules["h5py"] = h5
    sys.path.insert(0, "/root/reference")
    import lettuce as lt
    return lt


lt = import_reference()
torch.set_num_threads(1)
STRATS = {s.name: s for s in lt.StreamingStrategy}
STENCILS = {"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def make_collision(kind, flow, tau_minus=1.0):
    tau = flow.units.relaxation_parameter_lu
    return {"bgk": lambda: lt.BGKCollision(tau), "trt": lambda: lt.TRTCollision(tau, tau_minus),
            "kbc": lambda: lt.KBCCollision(), "none": lambda: lt.NoCollision(),
            "regularized": lambda: lt.RegularizedCollision(),
            "smagorinsky": lambda: lt.SmagorinskyCollision(tau, 0.17)}[kind]()


# ---------------------------------------------------------------- TGV cases
def tgv_case(name, stencil, res, re, ma, coll, steps, strategies, dtype=torch.float64, observables=False,
             perturb=0.0):
    """`perturb` > 0 multiplies the initial populations by (1 + perturb * uniform(-.5,.5)) with a fixed seed.
    Needed for D2Q9 KBC: the analytic f_neq initialisation has no higher-order part, so KBC's
    sum_h is pure rounding noise on the first step and gamma (kbc_collision.py:152) is ill-conditioned."""
    out = {}
    out["perturbed"] = np.array(perturb > 0)
    for sname in strategies:
        ctx = lt.Context(device="cpu", dtype=dtype, use_native=False)
        flow = lt.TaylorGreenVortex(ctx, list(res), re, ma, stencil=STENCILS[stencil]())
        if perturb > 0:
            rng = np.random.default_rng(7)
            flow.f = flow.f * ctx.convert_to_tensor(1.0 + perturb * (rng.random(tuple(flow.f.shape)) - 0.5))
        out["f0"] = npy(flow.f)
        out["tau"] = np.float64(flow.units.relaxation_parameter_lu)
        sim = lt.Simulation(flow, make_collision(coll, flow), [], STRATS[sname])
        sim(steps)
        out["f_" + sname] = npy(flow.f)
        if observables:
            out["energy_" + sname] = npy(lt.IncompressibleKineticEnergy(flow)(flow.f))
            out["enstrophy_" + sname] = npy(lt.Enstrophy(flow)(flow.f))
            out["maxvel_" + sname] = npy(lt.MaximumVelocity(flow)(flow.f))
            out["mass_" + sname] = npy(lt.Mass(flow)(flow.f))
            out["rho_" + sname] = npy(flow.rho())
            out["u_" + sname] = npy(flow.u())
    out["meta"] = np.array([stencil, coll, str(steps), str(re), str(ma)])
    out["res"] = np.array(res)
    save(name, **out)


# ---------------------------------------------------------------- obstacle cases (BASELINE.md section 5 helper)
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
    flow = cls(ctx, list(res), reynolds_number=100, mach_number=0.05,
               domain_length_x=res[0] / D, stencil=stencil)
    g = flow.grid
    c = [0.25 * g[0].max()] + [0.5 * gi.max() for gi in g[1:]]
    flow.mask = sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2
    return flow


def obstacle_case(name, cls, stencil, res, coll, steps, strategies, dtype=torch.float64):
    out = {}
    for sname in strategies:
        ctx = lt.Context(device="cpu", dtype=dtype, use_native=False)
        flow = make_obstacle(cls, ctx, res, STENCILS[stencil]())
        flow.initialize()                      # mask was set after __init__; re-initialise like a user would
        out["f0"] = npy(flow.f)
        out["solid"] = npy(flow.mask)
        out["tau"] = np.float64(flow.units.relaxation_parameter_lu)
        sim = lt.Simulation(flow, make_collision(coll, flow), [], STRATS[sname])
        out["ncm"] = npy(sim.no_collision_mask)
        out["nsm"] = npy(sim.no_streaming_mask)
        sim(steps)
        out["f_" + sname] = npy(flow.f)
        out["mass_" + sname] = npy(lt.Mass(flow, no_mass_mask=flow.mask)(flow.f))
    out["meta"] = np.array([stencil, coll, str(steps), "100", "0.05"])
    out["res"] = np.array(res)
    save(name, **out)


# ---------------------------------------------------------------- single-operator known answers on random f
class RandomFlow(lt.ExtFlow):
    """Uniform-resolution flow whose populations are overwritten by the caller."""

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


def random_collision_case():
    out = {}
    rng = np.random.default_rng(20261017)
    for stencil, res in (("D2Q9", [6, 5]), ("D3Q19", [4, 5, 6]), ("D3Q27", [4, 5, 6])):
        ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
        flow = RandomFlow(ctx, res, 50.0, 0.1, stencil=STENCILS[stencil]())
        w = np.asarray(flow.stencil.w).reshape((-1,) + (1,) * len(res))
        f0 = w * (1.0 + 0.2 * (rng.random((flow.stencil.q, *res)) - 0.5))
        out[f"{stencil}_f0"] = f0
        out[f"{stencil}_tau"] = np.float64(flow.units.relaxation_parameter_lu)
        for coll in ("bgk", "trt", "kbc", "regularized", "smagorinsky"):
            if coll == "kbc" and stencil == "D3Q19":
                continue
            flow.f = ctx.convert_to_tensor(f0)
            c = make_collision(coll, flow, tau_minus=0.8)
            out[f"{stencil}_{coll}"] = npy(c(flow))
        flow.f = ctx.convert_to_tensor(f0)
        out[f"{stencil}_rho"] = npy(flow.rho())
        out[f"{stencil}_u"] = npy(flow.u())
        out[f"{stencil}_feq"] = npy(flow.equilibrium(flow))
    save("random_collisions", **out)


# ---------------------------------------------------------------- the reference's own native known-answer cases
class Dummy16(RandomFlow):
    pass


def native_known_answers():
    """Torch-path side of tests/native/*.py (16x16 D2Q9)."""
    out = {}
    ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)

    def fresh():
        return Dummy16(ctx, [16, 16], 1.0, 0.05, stencil=lt.D2Q9())

    # tests/native/test_native_streaming.py:9-51 : nine tagged populations move by e_q
    flow = fresh(); flow.f[:] = 0.0
    for q in range(9):
        flow.f[q, 1, 1] = q + 1.0
    out["streaming_f0"] = npy(flow.f)
    lt.Simulation(flow, lt.NoCollision(), [])(1)
    out["streaming_f1"] = npy(flow.f)

    # tests/native/test_native_bgk_collision.py:26-71 and test_native_streaming_strategy.py:9-59
    for sname in STRATS:
        flow = fresh(); flow.f[:] = 1.0; flow.f[:, 2, 2] = 2.0
        out["bgk_f0"] = npy(flow.f)
        lt.Simulation(flow, lt.BGKCollision(2.0), [], STRATS[sname])(1)
        out["bgk_f1_" + sname] = npy(flow.f)

    # tests/native/test_native_bounce_back.py:11-74 : bounce-back everywhere but node (1,1), two steps
    class BB(lt.BounceBackBoundary):
        def make_no_collision_mask(self, shape, context):
            m = context.zero_tensor(shape, dtype=bool)
            m[0, :] = True; m[:, 0] = True; m[2:, :] = True; m[:, 2:] = True
            return m

    class BBFlow(Dummy16):
        @property
        def post_boundaries(self):
            return [BB(torch.ones(self.resolution))]

    flow = BBFlow(ctx, [16, 16], 1.0, 0.05, stencil=lt.D2Q9()); flow.f[:] = 0.0; flow.f[:, 1, 1] = 1.0
    out["bb_f0"] = npy(flow.f)
    sim = lt.Simulation(flow, lt.NoCollision(), [])
    sim(1); out["bb_f1"] = npy(flow.f)
    sim(1); out["bb_f2"] = npy(flow.f)

    # tests/native/test_native_equilibrium_pu.py:12-71 : equilibrium boundary + all-ones no-stream mask
    class EQ(lt.EquilibriumBoundaryPU):
        def make_no_streaming_mask(self, shape, context):
            return context.one_tensor(shape, dtype=bool)

    class EQFlow(Dummy16):
        @property
        def post_boundaries(self):
            m = torch.zeros(self.resolution, dtype=torch.bool); m[:, 3:5] = True
            return [EQ(self.context, self, m, velocity=[0.1, 0.05], pressure=0.02)]

    flow = EQFlow(ctx, [16, 16], 1.0, 0.05, stencil=lt.D2Q9())
    flow.f[:] = torch.rand(flow.f.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64) + 0.5
    out["eq_f0"] = npy(flow.f)
    lt.Simulation(flow, lt.NoCollision(), [])(1)
    out["eq_f1"] = npy(flow.f)
    save("native_known_answers", **out)


def poiseuille_case():
    """tests/collision/test_force.py set-up: PoiseuilleFlow2D 17^2, Re 1, Ma 0.02, BGK + Guo / ShanChen"""
    out = {}
    for fname, cls in (("guo", lt.Guo), ("shanchen", lt.ShanChen)):
        ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
        flow = lt.PoiseuilleFlow2D(context=ctx, resolution=17, reynolds_number=1, mach_number=0.02,
                                   initialize_with_zeros=True)
        acc = flow.units.convert_acceleration_to_lu(flow.acceleration)
        tau = flow.units.relaxation_parameter_lu
        sim = lt.Simulation(flow, lt.BGKCollision(tau, force=cls(flow=flow, tau=tau, acceleration=acc)), [])
        out["f0"] = npy(flow.f)
        out["tau"] = np.float64(tau)
        out["acceleration_lu"] = npy(acc)
        out["ncm"] = npy(sim.no_collision_mask)
        sim(40)
        out[f"f_{fname}_40"] = npy(flow.f)
        out[f"u_{fname}_40"] = npy(flow.u(acceleration=acc))
    save("poiseuille2d_forced", **out)


def other_flows_case():
    """DoublyPeriodicShear2D, BGK, a few steps.  (The reference's Cavity2D cannot be constructed: its
    post_boundaries calls EquilibriumBoundaryPU without the flow argument, liddrivencavity.py:63-70.)"""
    out = {}
    ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
    flow = lt.DoublyPeriodicShear2D(ctx, [24, 20], reynolds_number=1000, mach_number=0.05)
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    out["shear_f0"] = npy(flow.f)
    sim(15)
    out["shear_f15"] = npy(flow.f)
    save("shear2d_bgk", **out)


def stock_obstacle_case():
    """Stock lt.Obstacle (AntiBounceBackOutlet default, lettuce/ext/_flows/obstacle.py:107-122)."""
    obstacle_case("obstacle2d_abb_bgk", lt.Obstacle, "D2Q9", [48, 16], "bgk", 20, ["POST_STREAMING"])
    obstacle_case("obstacle3d_abb_bgk", lt.Obstacle, "D3Q19", [24, 12, 12], "bgk", 8, ["POST_STREAMING"])


def more_flows_case():
    """Initial states of the remaining entries of flow_by_name (lettuce/ext/_flows/_flow_by_name.py): Lamb-Oseen
    vortex, decaying turbulence (2-D with the pressure-Poisson start, 3-D without), Couette masks."""
    ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
    out = {}
    out["lamb_f0"] = npy(lt.LambOseenVortex2D(ctx, [48, 40], 100, 0.05).f)
    decay = lt.DecayingTurbulence(ctx, [32, 32], 1000, 0.05, k0=4, randseed=3)
    out["decay2d_f0"] = npy(decay.f)
    out["decay2d_spectrum"] = np.asarray(decay.energy_spectrum[0])
    out["decay2d_tau"] = np.float64(decay.units.relaxation_parameter_lu)
    out["decay3d_f0"] = npy(lt.DecayingTurbulence(ctx, [12, 12, 12], 1000, 0.05, k0=3, randseed=5).f)
    couette = lt.Simulation(lt.CouetteFlow2D(ctx, [16, 12], 100, 0.05), lt.BGKCollision(0.6), [])
    out["couette_ncm"] = npy(couette.no_collision_mask)
    out["couette_f0_nan"] = np.isnan(npy(couette.flow.f))     # characteristic velocity 0: the reference's state is NaN
    save("more_flows", **out)


def ebb_random_links_case():
    """Link lists of the example project's FullwayBounceBackBoundary for random solid masks that touch the domain
    border, with every periodicity combination: pins the wrap-at-minus-one / skip-at-n border behaviour of the
    reference's loops.  (HalfwayBounceBackBoundary's own legacy search cannot run in the reference: it reads
    `flow.context.d`, which does not exist, halfway_bounce_back_boundary.py:77,113.)"""
    import contextlib
    import io
    import itertools
    base = "/root/reference/examples/advanced_projects/efficient_bounce_back_obstacle"
    for sub in ("boundary", "simulation", "flow"):
        sys.path.insert(0, os.path.join(base, sub))
    from examples.advanced_projects.efficient_bounce_back_obstacle import FullwayBounceBackBoundary
    rng = np.random.default_rng(11)
    out = {}
    ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
    for tag, stencil, res in (("2d", "D2Q9", [9, 7]), ("3d", "D3Q19", [6, 5, 4]), ("3d27", "D3Q27", [5, 4, 4])):
        flow = lt.TaylorGreenVortex(ctx, list(res), 10, 0.05, stencil=STENCILS[stencil]())
        mask = rng.random(res) < 0.3
        other = (rng.random(res) < 0.15) & ~mask
        out[f"mask_{tag}"], out[f"other_{tag}"] = mask, other
        for k, per in enumerate(itertools.product([False, True], repeat=len(res))):
            per_arg = tuple(per) if len(res) == 3 else (per[0], per[1], None)
            with contextlib.redirect_stdout(io.StringIO()):
                fw = FullwayBounceBackBoundary(ctx, flow, mask, global_solid_mask=mask | other, periodicity=per_arg)
            out[f"fw_{tag}_{k}"] = npy(fw.f_index_fwbb)
            out[f"per_{tag}_{k}"] = np.array(per)
    save("ebb_random_links", **out)


def spectrum_case():
    """EnergySpectrum (observable_reporter.py:71-137) of evolved TGV states, incl. a non-cubic lattice."""
    out = {}
    for tag, stencil, res in (("2d", "D2Q9", [24, 24]), ("3d", "D3Q19", [12, 12, 12]), ("3d_ragged", "D3Q27", [12, 10, 8])):
        ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
        flow = lt.TaylorGreenVortex(ctx, list(res), 100.0, 0.05, stencil=STENCILS[stencil]())
        lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])(5)
        out["f_" + tag] = npy(flow.f)
        out["spectrum_" + tag] = npy(lt.EnergySpectrum(flow)(flow.f))
        out["meta_" + tag] = np.array([stencil, "100.0", "0.05"])
    save("energy_spectrum", **out)


def ebb_cases():
    """The reference's example project examples/advanced_projects/efficient_bounce_back_obstacle (EbbSimulation with
    fullway / halfway / linearly interpolated bounce-back applied after streaming, momentum-exchange force)."""
    base = "/root/reference/examples/advanced_projects/efficient_bounce_back_obstacle"
    for sub in ("boundary", "simulation", "flow"):
        sys.path.insert(0, os.path.join(base, sub))
    from examples.advanced_projects.efficient_bounce_back_obstacle.flow.obstacle_cylinder import ObstacleCylinder
    from examples.advanced_projects.efficient_bounce_back_obstacle.simulation.ebb_simulation import EbbSimulation
    import contextlib
    import io
    cases = [("ebb2d_hwbb", "D2Q9", [40, 20], "hwbb", "periodic", 6.0, 12),
             ("ebb2d_ibb1", "D2Q9", [40, 20], "ibb1", "periodic", 6.0, 12),
             ("ebb2d_fwbb", "D2Q9", [40, 20], "fwbb", "periodic", 6.0, 12),
             # (lateral_walls='bounceback' unpacks three resolution entries: 3-D only, obstacle_cylinder.py:161)
             ("ebb3d_fwbb_walls", "D3Q19", [20, 12, 5], "fwbb", "bounceback", 4.0, 6),
             ("ebb3d_ibb1", "D3Q19", [24, 14, 6], "ibb1", "periodic", 5.0, 8),
             ("ebb3d_hwbb_walls", "D3Q27", [20, 12, 5], "hwbb", "bounceback", 4.0, 6)]
    for name, stencil, res, bc, walls, diameter, steps in cases:
        ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
        with contextlib.redirect_stdout(io.StringIO()):
            flow = ObstacleCylinder(ctx, list(res), 100.0, 0.05, char_length_pu=1.0, char_length_lu=diameter,
                                    bc_type=bc, lateral_walls=walls, calc_force_coefficients=True,
                                    stencil=STENCILS[stencil](), u_init=1,
                                    perturb_init=len(res) == 2)   # the 3-D perturbation needs ny == nz (:279-284)
            sim = EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
        out = dict(f0=npy(flow.f), tau=np.float64(flow.units.relaxation_parameter_lu), res=np.array(res),
                   meta=np.array([stencil, bc, walls, str(steps), str(diameter)]),
                   obstacle_mask=flow.obstacle_mask.copy(), wall_mask=flow.wall_mask.copy(),
                   in_mask=flow.in_mask.copy(), u_inlet=np.asarray(flow.u_inlet, dtype=np.float64),
                   center=np.array([flow.x_pos_lu - 1, flow.y_pos_lu - 1, flow.radius_lu]),
                   ncm=npy(sim.no_collision_mask), nsm=npy(sim.no_streaming_mask))
        obstacle = sim.post_streaming_boundaries[-1]
        if bc == "fwbb":
            out["f_index_fwbb"] = npy(obstacle.f_index_fwbb)
        elif bc == "hwbb":
            out["f_index"] = npy(obstacle.f_index)
        else:
            out.update(f_index_lt=npy(obstacle.f_index_lt), f_index_gt=npy(obstacle.f_index_gt),
                       d_lt=npy(obstacle.d_lt), d_gt=npy(obstacle.d_gt))
        if walls == "bounceback":
            out["wall_f_index_fwbb"] = npy(sim.post_streaming_boundaries[0].f_index_fwbb)
        with contextlib.redirect_stdout(io.StringIO()):
            sim(steps)
        out["f"] = npy(flow.f)
        out["force"] = npy(obstacle.force_sum)
        save(name, **out)


def kbc_fp32_floor_case():
    """How far the REFERENCE's own torch fp32 path is from its own fp64 path for KBC, on exactly the inputs of the
    GPU parity tests (tests/test_gpu_parity.py: test_tgv_matches_live_oracle's KBC cases and the cylinder_d2q9_kbc
    golden).  KBC's stabiliser gamma (kbc_collision.py:152) is a ratio of two sums that shrink to rounding level
    in smooth flow, so any fp32 evaluation is noise-limited; these floors pin the tolerance of the fp32 KBC tests to
    the reference instead of a hand-written table.  Keys: tgv_<stencil>_<strategy>, cylinder_<strategy>."""
    out = {}
    steps = 10
    for stencil, res, re in (("D2Q9", [48, 40], 800.0), ("D3Q27", [20, 24, 28], 1600.0)):
        for sname in STRATS:
            fs = {}
            for dtype in (torch.float64, torch.float32):
                ctx = lt.Context(device="cpu", dtype=dtype, use_native=False)
                flow = lt.TaylorGreenVortex(ctx, list(res), re, 0.05, stencil=STENCILS[stencil]())
                if dtype == torch.float64:
                    # same construction as the test: fp64 initial state, seeded 1e-3 perturbation, rounded to fp32
                    rng = np.random.default_rng(11)
                    f0 = npy(flow.f) * (1.0 + 1e-3 * (rng.random(tuple(flow.f.shape)) - 0.5))
                    f0 = f0.astype(np.float32).astype(np.float64)
                flow.f = ctx.convert_to_tensor(f0)
                sim = lt.Simulation(flow, lt.KBCCollision(), [], STRATS[sname])
                sim(steps)
                fs[dtype] = npy(flow.f).astype(np.float64)
            out[f"tgv_{stencil}_{sname}"] = np.float64(np.max(np.abs(fs[torch.float32] - fs[torch.float64])
                                                              / np.abs(fs[torch.float64])))
    g = np.load(os.path.join(HERE, "cylinder_d2q9_kbc.npz"))
    for key in [k for k in g.files if k.startswith("f_")]:
        ctx = lt.Context(device="cpu", dtype=torch.float32, use_native=False)
        flow = make_obstacle(ObstacleEqOut, ctx, [int(r) for r in g["res"]], lt.D2Q9())
        flow.initialize()
        flow.f = ctx.convert_to_tensor(g["f0"])
        sim = lt.Simulation(flow, lt.KBCCollision(), [], STRATS[key[2:]])
        sim(int(g["meta"][2]))
        out["cylinder_" + key[2:]] = np.float64(np.max(np.abs(npy(flow.f).astype(np.float64) - g[key]) / np.abs(g[key])))
    for k, v in out.items():
        print(f"  {k}: {float(v):.3e}")
    save("kbc_fp32_floor", **out)


def pre_boundary_case():
    """A boundary BEFORE the collision (`Flow.pre_boundaries`): `collision_index` is 1, and because the reference
    fills the no-streaming mask with `collision_index` (lettuce/_simulation.py:104-107) every
    slot of every node is frozen -- nothing streams, every node is a "general" node of the engine's sparse kernel."""
    ctx = lt.Context(device="cpu", dtype=torch.float64, use_native=False)
    out = {}
    for stencil, res in (("D2Q9", [24, 16]), ("D3Q19", [12, 8, 16])):
        solid = np.zeros(res, dtype=bool)
        solid[3:6, 2:5] = True

        class PreFlow(lt.TaylorGreenVortex):
            @property
            def pre_boundaries(self):
                return [lt.BounceBackBoundary(torch.tensor(solid))]

        flow = PreFlow(ctx, list(res), 100.0, 0.05, stencil=STENCILS[stencil]())
        f0 = npy(flow.f).copy()
        out[f"{stencil}_f0"], out[f"{stencil}_solid"] = f0, solid
        out[f"{stencil}_tau"] = np.float64(flow.units.relaxation_parameter_lu)
        for sname in ("POST_STREAMING", "PRE_STREAMING"):
            flow.f = ctx.convert_to_tensor(f0.copy())
            sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], STRATS[sname])
            assert sim.collision_index == 1 and bool((sim.no_streaming_mask == 1).all())
            sim(6)
            out[f"{stencil}_{sname}"] = npy(flow.f)
    save("pre_boundary", **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["kbc_floor"]:
        kbc_fp32_floor_case()
        sys.exit(0)
    if sys.argv[1:] == ["pre_boundary"]:
        pre_boundary_case()
        sys.exit(0)
    if sys.argv[1:] == ["ebb"]:
        ebb_cases()
        sys.exit(0)
    if sys.argv[1:] == ["links"]:
        ebb_random_links_case()
        sys.exit(0)
    if sys.argv[1:] == ["flows"]:
        more_flows_case()
        sys.exit(0)
    if sys.argv[1:] == ["spectrum"]:
        spectrum_case()
        sys.exit(0)
    all4 = list(STRATS)
    tgv_case("tgv2d_d2q9_bgk", "D2Q9", [24, 24], 1.0, 0.05, "bgk", 10, all4, observables=True)
    tgv_case("tgv3d_d3q19_bgk", "D3Q19", [12, 12, 12], 1600.0, 0.05, "bgk", 10, all4, observables=True)
    tgv_case("tgv3d_d3q27_kbc", "D3Q27", [12, 10, 8], 1600.0, 0.05, "kbc", 10, ["POST_STREAMING", "PRE_STREAMING"],
             observables=True)
    tgv_case("tgv2d_d2q9_kbc", "D2Q9", [20, 16], 800.0, 0.1, "kbc", 10, ["POST_STREAMING"], perturb=1e-3)
    tgv_case("tgv3d_d3q27_trt", "D3Q27", [10, 12, 8], 400.0, 0.05, "trt", 6, ["POST_STREAMING", "PRE_STREAMING"])
    tgv_case("tgv3d_d3q19_trt", "D3Q19", [8, 8, 12], 400.0, 0.05, "trt", 6, ["POST_STREAMING"])
    tgv_case("tgv3d_d3q19_bgk_fp32", "D3Q19", [12, 12, 12], 1600.0, 0.05, "bgk", 10, ["POST_STREAMING"],
             dtype=torch.float32)
    tgv_case("tgv3d_d3q19_regularized", "D3Q19", [10, 8, 12], 1600.0, 0.05, "regularized", 8,
             ["POST_STREAMING", "PRE_STREAMING"])
    tgv_case("tgv3d_d3q27_smagorinsky", "D3Q27", [8, 10, 12], 1600.0, 0.05, "smagorinsky", 8,
             ["POST_STREAMING", "PRE_STREAMING"])
    tgv_case("tgv2d_d2q9_smagorinsky", "D2Q9", [20, 16], 3000.0, 0.1, "smagorinsky", 8, ["POST_STREAMING"])
    tgv_case("tgv2d_d2q9_regularized", "D2Q9", [20, 16], 3000.0, 0.1, "regularized", 8, ["POST_STREAMING"])
    obstacle_case("cylinder_d2q9_bgk", ObstacleEqOut, "D2Q9", [64, 16], "bgk", 30, all4)
    obstacle_case("sphere_d3q27_trt", ObstacleEqOut, "D3Q27", [32, 16, 16], "trt", 10,
                  ["POST_STREAMING", "PRE_STREAMING"])
    obstacle_case("sphere_d3q19_bgk", ObstacleEqOut, "D3Q19", [32, 16, 16], "bgk", 10, ["POST_STREAMING"])
    obstacle_case("cylinder_d2q9_kbc", ObstacleEqOut, "D2Q9", [64, 16], "kbc", 20, ["POST_STREAMING"])
    stock_obstacle_case()
    poiseuille_case()
    other_flows_case()
    random_collision_case()
    native_known_answers()
    spectrum_case()
    ebb_cases()
    more_flows_case()
    ebb_random_links_case()
    kbc_fp32_floor_case()
    pre_boundary_case()
