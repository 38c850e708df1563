#!/bin/bash
# TMA-staged kernel with the compact halo layout: tests, D3Q27 3 vs 4 stages, ncu of the KBC kernel
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 120 -x -k "tma" > $O/r2i_tma_tests.log 2>&1
tail -5 $O/r2i_tma_tests.log
: > $O/r2i_sweep.jsonl
for st in 3 4; do
  echo "{\"sweep\": \"ctas=1 stages=$st\"}" >> $O/r2i_sweep.jsonl
  LBM_B200_TMA=1 LBM_B200_TMA_STAGES=$st timeout 300 python scripts/bench_configs.py c3 c5 extra --small --pre-only >> $O/r2i_sweep.jsonl 2>&1
  LBM_B200_TMA=1 LBM_B200_TMA_STAGES=$st timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2i_bench_c3_st$st.json 2>> $O/r2i_bench.err; cat $O/r2i_bench_c3_st$st.json
done
cat $O/r2i_sweep.jsonl
LBM_B200_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:step_tma --launch-skip 6 --launch-count 1 \
    -o $O/r2i_kbc_tma -f python scripts/bench_configs.py c3 --small --pre-only > $O/r2i_ncu_kbc.log 2>&1; tail -2 $O/r2i_ncu_kbc.log
