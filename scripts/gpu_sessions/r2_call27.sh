#!/bin/bash
# round-2 GPU session 27 (2 GPUs): the final bench.py under torchrun as the driver launches it, reference arm under
# torchrun, and the host <-> device bandwidth of both GPUs alone / together (explains e2e at N > 1)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python scripts/pcie_probe.py 4 > $O/r2aa_pcie.json 2> $O/r2aa.err; cat $O/r2aa_pcie.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > $O/r2aa_bench_n2.json 2>> $O/r2aa.err; cat $O/r2aa_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2aa_bench_ref_n2.json 2>> $O/r2aa.err; cat $O/r2aa_bench_ref_n2.json
tail -5 $O/r2aa.err
