#!/bin/bash
# round-2 GPU session 5: plain/extras kernel split + contraction-free packed arithmetic: full tests, configs, ncu
set -x
mkdir -p gpurun_out
O=gpurun_out
python scripts/measure_peak.py > $O/r2e_peak.json 2>&1; cat $O/r2e_peak.json
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" > $O/r2e_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2e_gpu_tests.log 2>&1
tail -12 $O/r2e_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2e_smoke.log 2>&1; tail -3 $O/r2e_smoke.log
timeout 900 python scripts/bench_configs.py c2 c3 c4 c5 --small > $O/r2e_configs.jsonl 2>&1; cat $O/r2e_configs.jsonl
LBM_B200_LANES=1 timeout 600 python scripts/bench_configs.py c3 --small > $O/r2e_c3_lanes1.jsonl 2>&1; cat $O/r2e_c3_lanes1.jsonl
timeout 600 python bench.py --config c3 --quick --no-cpu --no-e2e > $O/r2e_bench_c3.json 2>> $O/r2e_bench.err; cat $O/r2e_bench_c3.json
timeout 600 python bench.py --quick --no-cpu > $O/r2e_bench.json 2>> $O/r2e_bench.err; cat $O/r2e_bench.json
# ncu: full section set with source counters (lineinfo variant of the same sources)
export LBM_B200_LIB=$PWD/variants/liblbm_b200_li.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 6 --launch-count 1 \
    -o $O/r2e_kbc_lanes2 -f python scripts/bench_configs.py c3 --small > $O/r2e_ncu_kbc.log 2>&1; tail -2 $O/r2e_ncu_kbc.log
LBM_B200_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 6 --launch-count 1 \
    -o $O/r2e_kbc_lanes1 -f python scripts/bench_configs.py c3 --small > $O/r2e_ncu_kbc1.log 2>&1; tail -2 $O/r2e_ncu_kbc1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 30 --launch-count 1 \
    -o $O/r2e_c4 -f python scripts/bench_configs.py c4 > $O/r2e_ncu_c4.log 2>&1; tail -2 $O/r2e_ncu_c4.log
ls -la $O/*.ncu-rep
