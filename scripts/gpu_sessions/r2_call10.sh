#!/bin/bash
# TMA-staged kernel with dynamic tile claiming: tests, sweep, bench, ncu
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 120 -x -k "tma" > $O/r2j_tma_tests.log 2>&1
tail -5 $O/r2j_tma_tests.log
: > $O/r2j_sweep.jsonl
for ctas in 1 2; do for st in 2 3 4 5; do
  echo "{\"sweep\": \"ctas=$ctas stages=$st\"}" >> $O/r2j_sweep.jsonl
  LBM_B200_TMA=1 LBM_B200_TMA_CTAS=$ctas LBM_B200_TMA_STAGES=$st timeout 300 python scripts/bench_configs.py c2 c3 c4 --small --pre-only >> $O/r2j_sweep.jsonl 2>&1
done; done
echo "{\"sweep\": \"defaults\"}" >> $O/r2j_sweep.jsonl
LBM_B200_TMA=1 timeout 300 python scripts/bench_configs.py c2 c3 c4 c5 extra --small --pre-only >> $O/r2j_sweep.jsonl 2>&1
cat $O/r2j_sweep.jsonl
LBM_B200_TMA=1 timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2j_bench_c3.json 2>> $O/r2j_bench.err; cat $O/r2j_bench_c3.json
LBM_B200_TMA=1 timeout 600 python bench.py --quick --no-cpu --no-e2e > $O/r2j_bench_c2.json 2>> $O/r2j_bench.err; cat $O/r2j_bench_c2.json
LBM_B200_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:step_tma --launch-skip 6 --launch-count 1 \
    -o $O/r2j_kbc_tma -f python scripts/bench_configs.py c3 --small --pre-only > $O/r2j_ncu_kbc.log 2>&1; tail -2 $O/r2j_ncu_kbc.log
