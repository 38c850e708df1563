#!/bin/bash
# round-2 GPU session 7: TMA-staged kernel, CTAs per SM / stage sweeps
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 120 -x -k "tma" > $O/r2g_tma_tests.log 2>&1
tail -5 $O/r2g_tma_tests.log
for ctas in 1 2; do
  LBM_B200_TMA=1 LBM_B200_TMA_CTAS=$ctas timeout 600 python scripts/bench_configs.py c2 c3 c4 c5 extra --small --pre-only > $O/r2g_configs_ctas$ctas.jsonl 2>&1
  cat $O/r2g_configs_ctas$ctas.jsonl
done
LBM_B200_TMA=1 LBM_B200_TMA_CTAS=1 LBM_B200_TMA_STAGES=3 timeout 600 python scripts/bench_configs.py c2 c4 --small --pre-only > $O/r2g_configs_ctas1_st3.jsonl 2>&1; cat $O/r2g_configs_ctas1_st3.jsonl
LBM_B200_TMA=1 timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2g_bench_c3_tma1.json 2>> $O/r2g_bench.err; cat $O/r2g_bench_c3_tma1.json
LBM_B200_TMA=1 timeout 600 python bench.py --quick --no-cpu --no-e2e > $O/r2g_bench_tma1.json 2>> $O/r2g_bench.err; cat $O/r2g_bench_tma1.json
tail -5 $O/r2g_bench.err
