#!/bin/bash
# round-2 GPU session 23 (2 GPUs): slabs re-validated on the shipped sources
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -c "from lettuce_b200 import build; print('source digest', build.source_digest())" > $O/r2w_box.txt 2>&1
timeout 900 python -m pytest tests/test_slab.py -m gpu -q --timeout 800 > $O/r2w_slab_tests.log 2>&1; tail -n 4 $O/r2w_slab_tests.log
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 50 > $O/r2w_bench_n2.json 2> $O/r2w_bench.err; cat $O/r2w_bench_n2.json; tail -n 2 $O/r2w_bench.err
