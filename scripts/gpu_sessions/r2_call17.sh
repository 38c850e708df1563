#!/bin/bash
# round-2 GPU session 17: reports of every step as one library call (lbm_step_moments_n)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > $O/r2q_tests.log 2>&1; tail -n 6 $O/r2q_tests.log
timeout 900 python scripts/bench_reporter_overhead.py > $O/r2q_reporter_overhead.jsonl 2>&1; cat $O/r2q_reporter_overhead.jsonl
