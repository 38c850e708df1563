#!/bin/bash
# round-2 GPU session 19: fused reductions in the staged kernel
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > $O/r2s_tests.log 2>&1; tail -n 6 $O/r2s_tests.log
timeout 900 python scripts/bench_reporter_overhead.py --small > $O/r2s_reporter_overhead.jsonl 2>&1; cat $O/r2s_reporter_overhead.jsonl
timeout 900 python bench.py --config c3 --quick --no-cpu > $O/r2s_bench_c3.json 2> $O/r2s_bench.err; cat $O/r2s_bench_c3.json; tail -n 3 $O/r2s_bench.err
