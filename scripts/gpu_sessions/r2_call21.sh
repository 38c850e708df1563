#!/bin/bash
# round-2 GPU session 21: final single-GPU evidence on the shipped sources (tests, smoke, bench lines, configurations)
set -x
mkdir -p gpurun_out
O=gpurun_out
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" > $O/r2u_box.txt 2>&1
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader >> $O/r2u_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=6 > $O/r2u_gpu_tests.log 2>&1
tail -n 16 $O/r2u_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2u_smoke.log 2>&1; tail -n 2 $O/r2u_smoke.log
timeout 900 python bench.py > $O/r2u_bench.json 2> $O/r2u_bench.err; cat $O/r2u_bench.json
timeout 900 python bench.py --config c3 > $O/r2u_bench_c3.json 2>> $O/r2u_bench.err; cat $O/r2u_bench_c3.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2u_bench_ref.json 2>> $O/r2u_bench.err; cat $O/r2u_bench_ref.json
tail -n 5 $O/r2u_bench.err
timeout 900 python scripts/bench_configs.py c1 c2 c3 c4 c5 extra > $O/r2u_configs.jsonl 2>&1; cat $O/r2u_configs.jsonl
timeout 900 python scripts/soak_staged_kernel.py 3000 > $O/r2u_soak.jsonl 2>&1; cat $O/r2u_soak.jsonl
# ncu: full section set of the two bench kernels at the bench size (one launch each)
timeout 900 ncu --set full --clock-control none -k regex:"step_kernel|step_tma" --launch-skip 4 --launch-count 1 -o $O/r2u_c2_512 -f \
    python bench.py --steps 3 --warmup 3 --quick --no-cpu --no-e2e > $O/r2u_ncu_c2.log 2>&1; tail -n 2 $O/r2u_ncu_c2.log
timeout 900 ncu --set full --clock-control none -k regex:"step_kernel|step_tma" --launch-skip 4 --launch-count 1 -o $O/r2u_c3_512 -f \
    python bench.py --config c3 --steps 3 --warmup 3 --quick --no-cpu --no-e2e > $O/r2u_ncu_c3.log 2>&1; tail -n 2 $O/r2u_ncu_c3.log
