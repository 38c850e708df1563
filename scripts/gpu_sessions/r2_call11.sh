#!/bin/bash
# TMA-staged kernel after the in-place race fix: repeated tests, mismatch probe, defaults, bench, ncu
set -x
mkdir -p gpurun_out
O=gpurun_out
for k in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 120 -k "tma" 2>&1 | tail -3; done > $O/r2k_tma_tests.log 2>&1
cat $O/r2k_tma_tests.log
python scripts/probes/tma_mismatch.py 12,32,256 5 40 2>&1 | tail -5
python scripts/probes/tma_mismatch.py 5,70,512 5 40 2>&1 | tail -5
: > $O/r2k_sweep.jsonl
echo "{\"sweep\": \"defaults\"}" >> $O/r2k_sweep.jsonl
LBM_B200_TMA=1 timeout 300 python scripts/bench_configs.py c2 c3 c4 c5 extra --small --pre-only >> $O/r2k_sweep.jsonl 2>&1
echo "{\"sweep\": \"D3Q27 3 stages\"}" >> $O/r2k_sweep.jsonl
LBM_B200_TMA=1 LBM_B200_TMA_STAGES=3 timeout 300 python scripts/bench_configs.py c3 --small --pre-only >> $O/r2k_sweep.jsonl 2>&1
cat $O/r2k_sweep.jsonl
LBM_B200_TMA=1 timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2k_bench_c3.json 2>> $O/r2k_bench.err; cat $O/r2k_bench_c3.json
LBM_B200_TMA=1 timeout 600 python bench.py --quick --no-cpu --no-e2e > $O/r2k_bench_c2.json 2>> $O/r2k_bench.err; cat $O/r2k_bench_c2.json
LBM_B200_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:step_tma --launch-skip 6 --launch-count 1 \
    -o $O/r2k_kbc_tma -f python scripts/bench_configs.py c3 --small --pre-only > $O/r2k_ncu_kbc.log 2>&1; tail -2 $O/r2k_ncu_kbc.log
