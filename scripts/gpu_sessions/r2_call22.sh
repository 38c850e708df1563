#!/bin/bash
# round-2 GPU session 22 (4 GPUs): the N = 4 points of the scaling curves
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 4 --no-e2e > $O/r2v_bench_n4.json 2> $O/r2v_bench.err; cat $O/r2v_bench_n4.json
timeout 600 $TR --master-port 29512 bench.py --gpus 4 --config c3 --no-e2e > $O/r2v_bench_c3_n4_weak.json 2>> $O/r2v_bench.err; cat $O/r2v_bench_c3_n4_weak.json
timeout 600 $TR --master-port 29513 scripts/bench_multi.py c3 > $O/r2v_c3_n4.json 2>> $O/r2v_bench.err; cat $O/r2v_c3_n4.json
timeout 900 $TR --master-port 29515 scripts/bench_multi.py c5 > $O/r2v_c5_n4.json 2>> $O/r2v_bench.err; cat $O/r2v_c5_n4.json
tail -n 3 $O/r2v_bench.err
