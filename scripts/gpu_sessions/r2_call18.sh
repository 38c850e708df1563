#!/bin/bash
# round-2 GPU session 18: outlets on different axes (nested neighbour look-ups), effect on the masked configurations
set -x
mkdir -p gpurun_out
O=gpurun_out
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2r_gpu_tests.log 2>&1; tail -n 8 $O/r2r_gpu_tests.log
timeout 900 python scripts/bench_configs.py c4 c5 ebb --small > $O/r2r_configs.jsonl 2>&1; cat $O/r2r_configs.jsonl
