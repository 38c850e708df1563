#!/bin/bash
# round-2 GPU session 6: first runs of the TMA-staged kernel
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 120 -x -k "tma" > $O/r2f_tma_tests.log 2>&1
tail -30 $O/r2f_tma_tests.log
if grep -q "failed\|error" $O/r2f_tma_tests.log; then
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 250 -x -k "tma_staged_kernel_is_bit_identical and D3Q19 and PRE" > $O/r2f_tma_memcheck.log 2>&1
  tail -40 $O/r2f_tma_memcheck.log
fi
for tma in 0 1; do
  LBM_B200_TMA=$tma timeout 600 python scripts/bench_configs.py c2 c3 c4 c5 extra --small > $O/r2f_configs_tma$tma.jsonl 2>&1
  cat $O/r2f_configs_tma$tma.jsonl
done
LBM_B200_TMA=1 timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2f_bench_c3_tma1.json 2>> $O/r2f_bench.err; cat $O/r2f_bench_c3_tma1.json
LBM_B200_TMA=1 timeout 600 python bench.py --quick --no-cpu --no-e2e > $O/r2f_bench_tma1.json 2>> $O/r2f_bench.err; cat $O/r2f_bench_tma1.json
LBM_B200_TMA=0 timeout 600 python bench.py --quick --no-cpu --no-e2e > $O/r2f_bench_tma0.json 2>> $O/r2f_bench.err; cat $O/r2f_bench_tma0.json
tail -5 $O/r2f_bench.err
