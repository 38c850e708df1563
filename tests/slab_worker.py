"""Worker for the multi-rank slab tests (run under torch.distributed.run or mp.spawn).

GPU mode (backend nccl): every rank steps its x-slab with the CUDA engine (peer-mapped halos), rank 0
additionally steps the whole lattice on its own GPU; the gathered slabs must equal the single-GPU
result BIT FOR BIT (same kernel arithmetic per node).

CPU mode (backend gloo): the host-side decomposition logic is exercised with the NumPy oracle as the
local stepper and explicit plane exchange through torch.distributed.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import lettuce_b200 as lt  # noqa: E402
from lettuce_b200 import slab  # noqa: E402


def gather_slabs(local: torch.Tensor, dec, device):
    """all ranks' slabs concatenated along x on every rank"""
    parts = []
    for r in range(dec.world):
        shape = list(local.shape)
        shape[1] = dec.sizes[r]
        buf = local.contiguous() if r == dec.rank else torch.empty(shape, dtype=local.dtype, device=device)
        dist.broadcast(buf, src=r)
        parts.append(buf)
    return torch.cat(parts, dim=1)


def gpu_case(stencil_cls, res, coll, strategy, dtype, steps, rank, world, dev, every=None):
    """`every`: report interval (default: one report after the last step).  With every = 1 the energy / maximum
    velocity reports ride on the slab step kernels (lbm_slab_step_moments), enstrophy is left out."""
    every = every or steps
    ctx = lt.Context(dev, dtype=dtype)
    dec = slab.SlabDecomposition(res[0], world, rank)
    flow = slab.SlabTaylorGreenVortex(ctx, res, 1600.0, 0.05, stencil_cls(), dec)
    make = {"bgk": lambda f: lt.BGKCollision(f.units.relaxation_parameter_lu), "kbc": lambda f: lt.KBCCollision(),
            "trt": lambda f: lt.TRTCollision(f.units.relaxation_parameter_lu)}[coll]
    energy = lt.ObservableReporter(slab.GlobalSum(lt.IncompressibleKineticEnergy(flow)), interval=every, out=None)
    enst = (lt.ObservableReporter(slab.SlabEnstrophy(flow), interval=every, out=None)
            if min(dec.sizes) >= 3 and every == steps else None)
    umax = lt.ObservableReporter(slab.GlobalMax(lt.MaximumVelocity(flow)), interval=every, out=None)
    sim = slab.SlabSimulation(flow, make(flow), [r for r in (energy, enst, umax) if r is not None], strategy, dec)
    f0 = gather_slabs(flow.f, dec, dev)
    # odd and even batch lengths exercise the buffer parity logic
    sim(1); sim(2); sim(steps - 3)
    got = gather_slabs(flow.f, dec, dev)
    ok = True
    if rank == 0:
        ref_flow = lt.TaylorGreenVortex(ctx, res, 1600.0, 0.05, stencil=stencil_cls())
        init_err = float((ref_flow.f - f0).abs().max())
        ref_flow.f = f0.clone()
        ref_energy = lt.ObservableReporter(lt.IncompressibleKineticEnergy(ref_flow), interval=every, out=None)
        ref_enst = lt.ObservableReporter(lt.Enstrophy(ref_flow), interval=every, out=None)
        ref_umax = lt.ObservableReporter(lt.MaximumVelocity(ref_flow), interval=every, out=None)
        ref = lt.Simulation(ref_flow, make(ref_flow), [ref_energy, ref_enst, ref_umax], strategy)
        ref(steps)
        same = torch.equal(ref_flow.f, got)
        e_rel = 0.0
        pairs = [(energy, ref_energy), (umax, ref_umax)] + ([(enst, ref_enst)] if enst is not None else [])
        for mine, theirs in pairs:
            rows_ok = [r[0] for r in mine.out] == [r[0] for r in theirs.out]
            same = same and rows_ok
            for a, b in zip(mine.out, theirs.out):
                e_rel = max(e_rel, abs(a[2] - b[2]) / abs(b[2]))
        print(f"[slab] {stencil_cls.__name__} {res} {coll} {strategy.name} {dtype} world={world} every={every}: "
              f"bit-exact={same} init_err={init_err:.1e} observables_rel={e_rel:.1e}", flush=True)
        # (rank partial sums are combined in a different order than the single-GPU fold: rounding-level differences)
        ok = same and init_err < 1e-6 and e_rel < (1e-9 if dtype == torch.float64 else 1e-6)
    sim.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    return bool(flag.item())


def _eq_out_boundaries(self):
    x = self.grid[0]
    return [lt.EquilibriumBoundaryPU(flow=self, context=self.context, mask=torch.abs(x) < 1e-6,
                                     velocity=self.units.characteristic_velocity_pu * self._unit_vector()),
            lt.EquilibriumOutletP(direction=self._unit_vector().tolist(), flow=self, rho_outlet=1.0),
            lt.BounceBackBoundary(self.mask)]


class ObstacleEqOut(lt.Obstacle):
    post_boundaries = property(_eq_out_boundaries)


class SlabObstacleEqOut(slab.SlabObstacle):
    post_boundaries = property(_eq_out_boundaries)


def _solid(flow, extent):
    g = flow.grid
    c = [0.25 * extent[0]] + [0.5 * e for e in extent[1:]]
    return sum((gi - ci) ** 2 for gi, ci in zip(g, c)) < 0.5 ** 2


def gpu_obstacle_case(stencil_cls, res, coll, strategy, dtype, steps, rank, world, dev, stock=False):
    """inlet + outlet + bounce-back obstacle flow on slabs vs the same flow on one GPU, bit for bit"""
    ctx = lt.Context(dev, dtype=dtype)
    dec = slab.SlabDecomposition(res[0], world, rank)
    D = res[1] / 8
    cls_slab, cls_one = (slab.SlabObstacle, lt.Obstacle) if stock else (SlabObstacleEqOut, ObstacleEqOut)
    flow = cls_slab(ctx, res, 100, 0.05, res[0] / D, dec, stencil=stencil_cls())
    flow.mask = _solid(flow, flow.global_extent_pu)
    flow.initialize()
    make = {"bgk": lambda f: lt.BGKCollision(f.units.relaxation_parameter_lu),
            "trt": lambda f: lt.TRTCollision(f.units.relaxation_parameter_lu), "kbc": lambda f: lt.KBCCollision()}[coll]
    sim = slab.SlabSimulation(flow, make(flow), [], strategy, dec)
    sim(1); sim(2); sim(steps - 3)
    got = gather_slabs(flow.f, dec, dev)
    ok = True
    if rank == 0:
        one = cls_one(ctx, res, 100, 0.05, res[0] / D, stencil=stencil_cls())
        one.mask = _solid(one, [gi.max() for gi in one.grid])
        one.initialize()
        ref = lt.Simulation(one, make(one), [], strategy)
        ref(steps)
        same = torch.equal(one.f, got)
        diff = float((one.f - got).abs().max())
        print(f"[slab-obstacle] {stencil_cls.__name__} {res} {coll} {strategy.name} {dtype} world={world} "
              f"stock={stock}: bit-exact={same} maxdiff={diff:.1e}", flush=True)
        ok = same
    sim.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    return bool(flag.item())


def gpu_main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S = lt.StreamingStrategy
    cases = [(lt.D3Q19, [32, 24, 40], "bgk", S.PRE_STREAMING, torch.float32, 9),
             (lt.D3Q19, [32, 24, 40], "bgk", S.POST_STREAMING, torch.float32, 9),
             (lt.D3Q27, [19, 16, 24], "kbc", S.POST_STREAMING, torch.float64, 8),
             (lt.D3Q27, [19, 16, 24], "trt", S.DOUBLE_STREAMING, torch.float32, 8),
             (lt.D2Q9, [48, 40], "bgk", S.PRE_STREAMING, torch.float64, 10),
             (lt.D2Q9, [50, 32], "kbc", S.POST_STREAMING, torch.float32, 10),
             (lt.D3Q19, [world * 2, 16, 32], "bgk", S.POST_STREAMING, torch.float32, 7),
             # large enough for the TMA-staged kernel: interior planes staged, cut planes by the lock-step kernel
             (lt.D3Q27, [world * 8, 32, 320], "kbc", S.PRE_STREAMING, torch.float32, 7),
             (lt.D3Q27, [world * 8 + 1, 40, 256], "kbc", S.PRE_STREAMING, torch.float32, 6)]
    ok = all([gpu_case(*c, rank, world, dev) for c in cases])
    # reports after every step: energy and maximum velocity are reduced inside the slab step kernels
    ok = all([gpu_case(*c, rank, world, dev, every=1) for c in cases[:2] + cases[4:6] + cases[7:8]]) and ok
    ocases = [(lt.D2Q9, [64, 32], "bgk", S.POST_STREAMING, torch.float64, 12, False),
              (lt.D2Q9, [64, 32], "bgk", S.PRE_STREAMING, torch.float32, 12, False),
              (lt.D3Q27, [32, 16, 16], "trt", S.POST_STREAMING, torch.float32, 9, False),
              (lt.D3Q19, [24, 16, 24], "bgk", S.DOUBLE_STREAMING, torch.float64, 8, False),
              (lt.D2Q9, [48, 24], "bgk", S.POST_STREAMING, torch.float64, 10, True),
              (lt.D3Q19, [world * 2, 16, 16], "bgk", S.POST_STREAMING, torch.float32, 6, False)]
    ok = all([gpu_obstacle_case(*c[:6], rank, world, dev, stock=c[6]) for c in ocases]) and ok
    print(f"[slab-worker] rank {rank}: all cases done, ok={ok}", flush=True)
    torch.cuda.synchronize(dev)
    dist.barrier()
    dist.destroy_process_group()
    print(f"[slab-worker] rank {rank}: process group destroyed", flush=True)
    if not ok:
        sys.exit(1)


# ------------------------------------------------------------------------------- CPU / gloo
def cpu_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lbm_oracle as lo
        ctx = lt.Context("cpu", dtype=torch.float64)
        dev = torch.device("cpu")
        out = {}
        for name, cls, res in (("D3Q19", lt.D3Q19, [11, 6, 8]), ("D2Q9", lt.D2Q9, [10, 12])):
            dec = slab.SlabDecomposition(res[0], world, rank)
            flow = slab.SlabTaylorGreenVortex(ctx, res, 400.0, 0.05, cls(), dec)
            assert list(flow.f.shape[1:]) == [dec.nx_local] + res[1:]
            whole = gather_slabs(flow.f, dec, dev)
            ref = lt.TaylorGreenVortex(ctx, res, 400.0, 0.05, stencil=cls())
            out[name + "_init"] = float((whole - ref.f).abs().max())
            assert flow.units.relaxation_parameter_lu == ref.units.relaxation_parameter_lu
            # lock-stepped slab stepping with explicit plane exchange, oracle as the local stepper
            st = lo.stencil(name)
            coll = dict(kind="bgk", tau=ref.units.relaxation_parameter_lu)
            f = flow.f.numpy().copy()
            g = ref.f.numpy().copy()
            for _ in range(4):
                fc = lo.collide(st, f, coll)                      # node-local
                # ring exchange of the boundary planes (post-collision), then stream on the extended slab
                send_hi, send_lo = torch.from_numpy(fc[:, -1].copy()), torch.from_numpy(fc[:, 0].copy())
                recv_lo, recv_hi = torch.empty_like(send_hi), torch.empty_like(send_lo)
                reqs = [dist.isend(send_hi, dec.hi), dist.isend(send_lo, dec.lo),
                        dist.irecv(recv_lo, dec.lo), dist.irecv(recv_hi, dec.hi)] if world > 2 else None
                if world == 2:      # both neighbours are the same rank: order the messages by tag
                    reqs = [dist.isend(send_hi, dec.hi, tag=1), dist.isend(send_lo, dec.lo, tag=2),
                            dist.irecv(recv_lo, dec.lo, tag=1), dist.irecv(recv_hi, dec.hi, tag=2)]
                for r_ in reqs:
                    r_.wait()
                ext = np.concatenate([recv_lo.numpy()[:, None], fc, recv_hi.numpy()[:, None]], axis=1)
                f = lo.stream(st, ext)[:, 1:-1]
                g = lo.step(st, g, coll)
            whole = gather_slabs(torch.from_numpy(np.ascontiguousarray(f)), dec, dev).numpy()
            out[name + "_step"] = float(np.abs(whole - g).max())
        # boundaries on slabs: every rank's local label / no-stream masks are the slices of the global ones,
        # with the x-normal outlet active only on the rank that owns the plane
        for name, cls, res in (("D2Q9", lt.D2Q9, [24, 16]), ("D3Q19", lt.D3Q19, [12, 8, 8])):
            dec = slab.SlabDecomposition(res[0], world, rank)
            D = res[1] / 8
            flow = SlabObstacleEqOut(ctx, res, 100, 0.05, res[0] / D, dec, stencil=cls())
            flow.mask = _solid(flow, flow.global_extent_pu)
            sim = slab.SlabSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [],
                                      lt.StreamingStrategy.POST_STREAMING, dec)
            one = ObstacleEqOut(ctx, res, 100, 0.05, res[0] / D, stencil=cls())
            one.mask = _solid(one, [gi.max() for gi in one.grid])
            ref = lt.Simulation(one, lt.BGKCollision(one.units.relaxation_parameter_lu), [])
            sl = dec.local_slice()
            out[name + "_mask"] = float((flow.mask.to(torch.uint8) != one.mask[sl].to(torch.uint8)).sum())
            out[name + "_ncm"] = float((sim.no_collision_mask != ref.no_collision_mask[sl]).sum())
            out[name + "_nsm"] = float((sim.no_streaming_mask != ref.no_streaming_mask[:, sl]).sum())
            owner = rank == world - 1
            assert [getattr(b, "_slab_disabled", False) for b in sim.post_boundaries] == [False, not owner, False]
            d = lt.native.describe(sim)
            assert d["ops"][2]["side"] == (1 if owner else 0)
        # checkpoints: rank 0 writes the GLOBAL lattice in Flow.dump's format; a single-process Flow.load reads it,
        # and SlabSimulation.load hands every rank its planes back
        import pickle
        path = f"/tmp/lettuce_b200_slab_ckpt_{port}.pkl"
        res = [11, 6, 8]
        dec = slab.SlabDecomposition(res[0], world, rank)
        flow = slab.SlabTaylorGreenVortex(ctx, res, 400.0, 0.05, lt.D3Q19(), dec)
        sim = slab.SlabSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [],
                                  lt.StreamingStrategy.POST_STREAMING, dec)
        mine = flow.f.clone()
        sim.dump(path)
        one = lt.TaylorGreenVortex(ctx, res, 400.0, 0.05, stencil=lt.D3Q19())
        expect = one.f.clone()
        one.f = torch.zeros_like(one.f)
        one.load(path)
        out["ckpt_global_step"] = float((one.f - expect).abs().max())
        flow.f = torch.zeros_like(mine)
        sim.load(path)
        out["ckpt_reload_step"] = float((flow.f - mine).abs().max())
        dist.barrier()
        if rank == 0:
            with open(path, "wb") as fh:
                pickle.dump(np.zeros((19, res[0] + 1, 6, 8)), fh)       # wrong global shape: every rank raises
        dist.barrier()
        try:
            sim.load(path)
            out["ckpt_shape_step"] = 1.0
        except ValueError:
            out["ckpt_shape_step"] = 0.0
        dist.barrier()
        if rank == 0:
            os.remove(path)
        results[rank] = out
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    import traceback
    faulthandler.enable()
    try:
        gpu_main()
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        raise
