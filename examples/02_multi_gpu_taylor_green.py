"""Taylor-Green vortex on x-slabs, one process per GPU:

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/02_multi_gpu_taylor_green.py [--resolution 512]

The global lattice is [resolution * world, resolution, resolution]; halos are read and written by the step kernel
itself through peer-mapped neighbour buffers.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lettuce_b200 as lt  # noqa: E402
from lettuce_b200 import slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolution", type=int, default=256, help="nodes per axis per GPU")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--checkpoint", default=None, help="write the global lattice here at the end (rank 0)")
    args = ap.parse_args()
    if "RANK" not in os.environ:
        sys.exit("one process per GPU: launch with `torchrun --nproc-per-node N --master-addr 127.0.0.1 " + sys.argv[0] + "`")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    ctx = lt.Context(f"cuda:{local}", dtype=torch.float32)
    n = args.resolution
    dec = slab.SlabDecomposition(nx_global=n * world, world=world, rank=rank)
    flow = slab.SlabTaylorGreenVortex(ctx, [n * world, n, n], 1600, 0.05, lt.D3Q19(), dec)
    energy = lt.ObservableReporter(slab.GlobalSum(lt.IncompressibleKineticEnergy(flow)), interval=100, out=None)
    sim = slab.SlabSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [energy],
                              lt.StreamingStrategy.PRE_STREAMING, dec)
    mlups = sim(args.steps) * world
    if args.checkpoint:
        sim.dump(args.checkpoint)
    if rank == 0:
        print(f"{world} GPUs: {mlups:.0f} MLUPS, E_kin(t_end) = {energy.out[-1][2]:.6f}")
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
