"""Measure every BASELINE.json configuration on ONE B200 (not the driver's bench contract; numbers
for DESIGN.md / profiles).  Prints one JSON line per case.

    python scripts/bench_configs.py [c1 c2 c3 c4 c5 ebb extra ...] [--small]
"""
import gc
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lettuce_b200 as lt  # noqa: E402
from lettuce_b200 import native  # noqa: E402

PEAK = 6544.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


class ObstacleEqOut(lt.Obstacle):
    @property
    def post_boundaries(self):
        x = self.grid[0]
        return [lt.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                         velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
                lt.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
                lt.BounceBackBoundary(self.mask)]


def make_obstacle(ctx, res, stencil):
    D = res[1] / 8
    flow = ObstacleEqOut(ctx, list(res), reynolds_number=100, mach_number=0.05, domain_length_x=res[0] / D,
                         stencil=stencil)
    g = flow.grid
    c = [0.25 * g[0].max()] + [0.5 * gi.max() for gi in g[1:]]
    flow.mask = sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2
    flow.initialize()
    return flow


def timed(sim, steps, warmup):
    native.invoke_n(sim, warmup)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = native.launch_count()
    a.record()
    native.invoke_n(sim, steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, native.launch_count() - l0


def report(name, flow, sim, steps=50, warmup=5, **extra):
    gc.collect()
    torch.cuda.empty_cache()
    ms, launches = timed(sim, steps, warmup)
    n = 1
    for r in flow.resolution:
        n *= r
    q, es = flow.stencil.q, flow.f.element_size()
    mlups = n / 1e6 / (ms * 1e-3)
    gbs = mlups * 1e6 * 2 * q * es / 1e9
    ok = bool(torch.isfinite(flow.f).all())
    print(json.dumps(dict(case=name, resolution=flow.resolution, stencil=type(flow.stencil).__name__,
                          dtype=str(flow.f.dtype), collision=type(sim.collision).__name__,
                          streaming=sim.streaming_strategy.name, ms_per_step=ms, mlups=mlups, bytes_per_node=2 * q * es,
                          gbs=gbs, frac_of_measured_peak=gbs / PEAK, launches_per_step=launches / steps,
                          kernel=native.engine_of(sim).variant_name, finite=ok, **extra)), flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c3", "c4", "c5"]
    small = "--small" in sys.argv
    pre_only = "--pre-only" in sys.argv
    S = lt.StreamingStrategy
    f32, f64 = torch.float32, torch.float64
    for case in args:
        if case == "c1":      # TGV2D D2Q9 BGK 256^2 fp64, 1000 steps, PRE (cli.py:98-118)
            ctx = lt.Context("cuda", dtype=f64)
            flow = lt.TaylorGreenVortex(ctx, [256, 256], 1.0, 0.05, stencil=lt.D2Q9())
            sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], S.PRE_STREAMING)
            report("C1 TGV2D D2Q9 BGK 256^2 fp64", flow, sim, steps=1000, warmup=50, note="4.7 MB: L2 resident, launch bound")
            t0 = time.perf_counter(); m = sim(1000); dt = time.perf_counter() - t0
            print(json.dumps(dict(case="C1 via Simulation.__call__(1000)", mlups=m, wall_s=dt)), flush=True)
        elif case == "c2":
            for n in ((256,) if small else (256, 512)):
                for strat in ((S.PRE_STREAMING,) if pre_only else (S.PRE_STREAMING, S.POST_STREAMING)):
                    ctx = lt.Context("cuda", dtype=f32)
                    flow = lt.TaylorGreenVortex(ctx, [n] * 3, 1600.0, 0.05, stencil=lt.D3Q19())
                    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], strat)
                    report(f"C2 TGV3D D3Q19 BGK {n}^3 fp32", flow, sim)
                    del flow, sim
                    gc.collect(); torch.cuda.empty_cache()
        elif case == "c3":
            for n in ((256,) if small else (256, 512)):
                for strat in ((S.PRE_STREAMING,) if pre_only else (S.PRE_STREAMING, S.POST_STREAMING)):
                    ctx = lt.Context("cuda", dtype=f32)
                    flow = lt.TaylorGreenVortex(ctx, [n] * 3, 1600.0, 0.05, stencil=lt.D3Q27())
                    sim = lt.Simulation(flow, lt.KBCCollision(), [], strat)
                    report(f"C3 TGV3D D3Q27 KBC {n}^3 fp32", flow, sim, steps=30)
                    del flow, sim
                    gc.collect(); torch.cuda.empty_cache()
        elif case == "c4":
            for dt_ in ((f32,) if pre_only else (f32, f64)):
                for strat in ((S.PRE_STREAMING,) if pre_only else (S.POST_STREAMING, S.PRE_STREAMING)):
                    ctx = lt.Context("cuda", dtype=dt_)
                    flow = make_obstacle(ctx, [4096, 1024], lt.D2Q9())
                    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], strat)
                    eng = native.engine_of(sim)
                    report("C4 cylinder D2Q9 BGK 4096x1024", flow, sim, steps=200, warmup=20,
                           general_nodes=int(eng.desc.n_general))
                    del flow, sim, eng
        elif case == "c5":
            res = [512, 256, 256] if small else [1024, 512, 512]
            for strat in ((S.PRE_STREAMING,) if pre_only else (S.POST_STREAMING, S.PRE_STREAMING)):
                ctx = lt.Context("cuda", dtype=f32)
                flow = make_obstacle(ctx, res, lt.D3Q27())
                torch.cuda.empty_cache()
                sim = lt.Simulation(flow, lt.TRTCollision(flow.units.relaxation_parameter_lu), [], strat)
                eng = native.engine_of(sim)
                sim.no_streaming_mask = None      # 27 N bytes; the engine keeps the packed form
                torch.cuda.empty_cache()
                report(f"C5 sphere D3Q27 TRT {'x'.join(map(str, res))} fp32", flow, sim, steps=20, warmup=3,
                       general_nodes=int(eng.desc.n_general),
                       max_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
                del flow, sim, eng
        elif case == "ebb":
            # cylinder with link-wise bounce-back after streaming (EbbSimulation): step loop in Python vs the
            # batched library call (LBM_B200_EBB_BATCH=1)
            res, diameter = ([1024, 256], 32.0) if small else ([4096, 1024], 128.0)
            for bc in ("fwbb", "hwbb", "ibb1"):
                for batch in ("0", "1"):
                    os.environ["LBM_B200_EBB_BATCH"] = batch
                    ctx = lt.Context("cuda", dtype=f32)
                    flow = lt.ObstacleCylinder(ctx, res, 100.0, 0.05, char_length_pu=1.0, char_length_lu=diameter,
                                               bc_type=bc, u_init=1, calc_force_coefficients=True, stencil=lt.D2Q9())
                    sim = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
                    sim(20)
                    torch.cuda.synchronize()
                    l0 = native.launch_count()
                    t0 = time.perf_counter()
                    sim(200)
                    dt = (time.perf_counter() - t0) / 200
                    n = res[0] * res[1]
                    print(json.dumps(dict(case=f"EBB cylinder D2Q9 BGK {res[0]}x{res[1]} {bc}", batch=batch,
                                          links=sim.post_streaming_boundaries[-1].n_links, ms_per_step=dt * 1e3,
                                          mlups=n / 1e6 / dt, launches_per_step=(native.launch_count() - l0) / 200,
                                          force=sim.post_streaming_boundaries[-1].force_sum.tolist(),
                                          finite=bool(torch.isfinite(flow.f).all()))), flush=True)
                    del flow, sim
            os.environ.pop("LBM_B200_EBB_BATCH", None)
        elif case == "kbc2d":     # D2Q9 KBC: LDG vs staged kernel (LBM_B200_TMA=0|1)
            for strat in (S.PRE_STREAMING,):
                ctx = lt.Context("cuda", dtype=f32)
                flow = lt.TaylorGreenVortex(ctx, [4096, 4096], 1600.0, 0.05, stencil=lt.D2Q9())
                sim = lt.Simulation(flow, lt.KBCCollision(), [], strat)
                report("TGV2D D2Q9 KBC 4096^2 fp32", flow, sim, steps=100)
                del flow, sim
        elif case == "extra":
            for st, coll, dt_ in ((lt.D3Q27, "bgk", f32), (lt.D3Q27, "trt", f32), (lt.D3Q19, "bgk", f64),
                                  (lt.D3Q19, "trt", f32), (lt.D3Q27, "kbc", f64)):
                if pre_only and dt_ == f64:
                    continue
                ctx = lt.Context("cuda", dtype=dt_)
                flow = lt.TaylorGreenVortex(ctx, [256] * 3, 1600.0, 0.05, stencil=st())
                tau = flow.units.relaxation_parameter_lu
                c = {"bgk": lt.BGKCollision(tau), "trt": lt.TRTCollision(tau), "kbc": lt.KBCCollision()}[coll]
                sim = lt.Simulation(flow, c, [], S.PRE_STREAMING)
                report(f"extra TGV3D {st.__name__} {coll} 256^3", flow, sim, steps=30)
                del flow, sim
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
