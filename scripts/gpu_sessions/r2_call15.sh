#!/bin/bash
# round-2 GPU session 15 (8 GPUs): slab bit-exactness at N = 8, weak / strong scaling of C2, C3, C5
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/r2o_topo.txt 2>&1
timeout 600 python -m pytest tests/test_slab.py -m gpu -q --timeout 500 > $O/r2o_slab_tests.log 2>&1; tail -n 5 $O/r2o_slab_tests.log
timeout 600 $TR --master-port 29511 bench.py --gpus 8 > $O/r2o_bench_n8.json 2> $O/r2o_bench.err; cat $O/r2o_bench_n8.json
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --config c3 --no-e2e > $O/r2o_bench_c3_n8_weak.json 2>> $O/r2o_bench.err; cat $O/r2o_bench_c3_n8_weak.json
timeout 600 $TR --master-port 29513 scripts/bench_multi.py c3 > $O/r2o_c3_n8.json 2>> $O/r2o_bench.err; cat $O/r2o_c3_n8.json
timeout 600 $TR --master-port 29514 scripts/bench_multi.py c3 --strategy POST_STREAMING >> $O/r2o_c3_n8.json 2>> $O/r2o_bench.err; tail -n 1 $O/r2o_c3_n8.json
timeout 900 $TR --master-port 29515 scripts/bench_multi.py c5 > $O/r2o_c5_n8.json 2>> $O/r2o_bench.err; cat $O/r2o_c5_n8.json
tail -n 5 $O/r2o_bench.err
