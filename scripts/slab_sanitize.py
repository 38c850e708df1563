"""Small multi-GPU slab run for compute-sanitizer (memcheck / racecheck / synccheck): one unmasked and one masked
case on tiny lattices, the same code paths as tests/slab_worker.py (peer-mapped halos, in-kernel lock step, sparse
kernel publishing the progress counters, fused reductions).

    compute-sanitizer --tool memcheck --target-processes all python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/slab_sanitize.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lettuce_b200 as lt  # noqa: E402
import slab_worker as sw  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S = lt.StreamingStrategy
    ok = sw.gpu_case(lt.D3Q19, [8 * world, 8, 32], "bgk", S.PRE_STREAMING, torch.float32, 5, rank, world, dev, every=1)
    ok = sw.gpu_case(lt.D2Q9, [8 * world, 64], "kbc", S.POST_STREAMING, torch.float32, 5, rank, world, dev, every=1) and ok
    ok = sw.gpu_obstacle_case(lt.D2Q9, [16 * world, 32], "bgk", S.POST_STREAMING, torch.float32, 6, rank, world, dev) and ok
    ok = sw.gpu_obstacle_case(lt.D3Q27, [8 * world, 16, 16], "trt", S.PRE_STREAMING, torch.float32, 5, rank, world, dev) and ok
    print(f"[sanitize] rank {rank}: ok={ok}", flush=True)
    torch.cuda.synchronize(dev)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
