#!/bin/bash
# round-2 GPU session 12: final single-GPU validation of the round's kernels (tests, smoke, bench, ncu evidence)
set -x
mkdir -p gpurun_out
O=gpurun_out
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" > $O/r2l_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > $O/r2l_gpu_tests.log 2>&1
tail -25 $O/r2l_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2l_smoke.log 2>&1; tail -3 $O/r2l_smoke.log
timeout 900 python bench.py > $O/r2l_bench.json 2> $O/r2l_bench.err; cat $O/r2l_bench.json
timeout 900 python bench.py --config c3 > $O/r2l_bench_c3.json 2>> $O/r2l_bench.err; cat $O/r2l_bench_c3.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2l_bench_ref.json 2>> $O/r2l_bench.err; cat $O/r2l_bench_ref.json
tail -5 $O/r2l_bench.err
timeout 900 python scripts/bench_configs.py c1 c2 c3 c4 c5 extra > $O/r2l_configs.jsonl 2>&1; cat $O/r2l_configs.jsonl
# ncu: DRAM bytes of the two bench kernels at the bench size (for roofline.traffic), launch list of a short bench run
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
    --clock-control none -k regex:"step_kernel|step_tma" --launch-skip 4 --launch-count 2 --csv --log-file $O/r2l_dram_c2_512_pre.csv \
    python bench.py --steps 3 --warmup 3 --quick --no-cpu --no-e2e > $O/r2l_ncu_c2.log 2>&1; tail -4 $O/r2l_dram_c2_512_pre.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
    --clock-control none -k regex:"step_kernel|step_tma" --launch-skip 4 --launch-count 2 --csv --log-file $O/r2l_dram_c3_512_pre.csv \
    python bench.py --config c3 --steps 3 --warmup 3 --quick --no-cpu --no-e2e > $O/r2l_ncu_c3.log 2>&1; tail -4 $O/r2l_dram_c3_512_pre.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2l_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --quick --no-cpu > $O/r2l_ncu_launches.log 2>&1; tail -3 $O/r2l_ncu_launches.log
