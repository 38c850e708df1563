#!/bin/bash
# round-2 GPU session 24: the pushing step on the staged machinery (halo warp, overlapped tiles)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -x > $O/r2x_tests.log 2>&1; tail -n 8 $O/r2x_tests.log
timeout 600 python scripts/soak_staged_kernel.py 1500 > $O/r2x_soak.jsonl 2>&1; cat $O/r2x_soak.jsonl
for tma in 0 1; do
  echo "{\"sweep\": \"POST step by step (no lazy batches), LBM_B200_TMA=$tma\"}" >> $O/r2x_post.jsonl
  LBM_B200_LAZY_POST_MIN=0 LBM_B200_TMA=$tma timeout 600 python scripts/bench_configs.py c2 c3 c4 c5 --small >> $O/r2x_post.jsonl 2>&1
done
cat $O/r2x_post.jsonl
