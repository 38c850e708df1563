"""BASELINE.json configs[2] and configs[4] on N GPUs of one box (x-slabs, one process per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_multi.py c3|c5 [--small] [--steps K] [--strategy PRE_STREAMING] [--energy]

  c3  TaylorGreenVortex3D D3Q27 KBC 512^3 fp32, the FIXED global lattice split into N x-slabs (strong scaling)
  c5  sphere Obstacle D3Q27 TRT, [1024 N, 512, 512] fp32: 1024x512x512 per GPU (weak scaling), inlet +
      EquilibriumOutletP + bounce-back, the halo exchange done by the step kernels through peer memory
  --small   c3 at 256^3, c5 at 512x256x256 per GPU
  --energy  additionally time the same steps with a global IncompressibleKineticEnergy report after EVERY step
            (reduced inside the slab step kernels + one all-reduce per step)

Rank 0 prints one JSON line per measurement; times are CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lettuce_b200 as lt  # noqa: E402
from lettuce_b200 import native, slab  # noqa: E402

PEAK = 6544.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def _eq_out_boundaries(self):
    x = self.grid[0]
    return [lt.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                     velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
            lt.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
            lt.BounceBackBoundary(self.mask)]


class SlabObstacleEqOut(slab.SlabObstacle):
    post_boundaries = property(_eq_out_boundaries)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c5"])
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--strategy", default="PRE_STREAMING")
    ap.add_argument("--energy", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = lt.Context(dev, dtype=torch.float32)
    strategy = lt.StreamingStrategy[args.strategy]

    if args.config == "c3":
        n = 256 if args.small else 512
        res = [n, n, n]
        flow, sim, _ = slab.make_tgv_slab_simulation(ctx, res, 1600.0, 0.05, lt.D3Q27(), strategy,
                                                     collision_factory=lambda fl: lt.KBCCollision())
        name, scaling = f"C3 TGV3D D3Q27 KBC {n}^3 fp32 split over {world} GPUs", "strong"
    else:
        per = [512, 256, 256] if args.small else [1024, 512, 512]
        res = [per[0] * world, per[1], per[2]]
        dec = slab.SlabDecomposition(res[0], world, rank)
        D = res[1] / 8
        flow = SlabObstacleEqOut(ctx, res, 100, 0.05, res[0] / D, dec, stencil=lt.D3Q27())
        g = flow.grid
        ext = flow.global_extent_pu
        # the sphere sits where the single-GPU config has it (a quarter of ONE slab length from the inlet), so the
        # solid-node count does not depend on the number of GPUs
        c = [0.25 * ext[0] / world] + [0.5 * e for e in ext[1:]]
        flow.mask = sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2
        flow.initialize()
        torch.cuda.empty_cache()
        sim = slab.SlabSimulation(flow, lt.TRTCollision(flow.units.relaxation_parameter_lu), [], strategy, dec)
        sim.no_streaming_mask = None          # 27 N bytes; the engine keeps the packed form
        torch.cuda.empty_cache()
        name, scaling = f"C5 sphere D3Q27 TRT {'x'.join(map(str, per))} fp32 per GPU, {world} GPUs", "weak"
    nodes = 1
    for r in res:
        nodes *= r
    eng = native.engine_of(sim)

    def timed(stepper, steps):
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stepper(steps)
        b.record()
        torch.cuda.synchronize(dev)
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    native.invoke_n(sim, args.warmup)
    ms = timed(lambda k: native.invoke_n(sim, k), args.steps)
    out = dict(case=name, scaling=scaling, n_gpus=world, global_lattice=res, streaming=strategy.name,
               steps=args.steps, ms_per_step=ms / args.steps, mlups=args.steps * nodes / 1e6 / (ms * 1e-3),
               general_nodes_rank0=int(eng.desc.n_general), kernel=eng.variant_name,
               max_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
    out["frac_of_measured_peak_per_gpu"] = out["mlups"] * 1e6 * 2 * 27 * 4 / 1e9 / PEAK / world
    if args.energy:
        rep = lt.ObservableReporter(slab.GlobalSum(lt.IncompressibleKineticEnergy(flow)), interval=1, out=None)
        sim.reporter.append(rep)
        flow.i = 1
        ms_e = timed(lambda k: sim(k), args.steps)
        out["ms_per_step_with_energy_every_step"] = ms_e / args.steps
        out["energy_overhead"] = ms_e / ms - 1.0
        out["energy_last"] = rep.out[-1][2]
        sim.reporter.pop()
    ok = bool(torch.isfinite(flow.f).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["finite"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
