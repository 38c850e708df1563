#!/bin/bash
# round-2 GPU session 26: final single-GPU validation on the shipped sources (all GPU tests, smoke, the three contract
# lines incl. the native_gpu_reference and c3_512 legs)
set -x
mkdir -p gpurun_out
O=gpurun_out
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" > $O/r2z_box.txt 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $O/r2z_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 --durations=6 > $O/r2z_gpu_tests.log 2>&1
tail -12 $O/r2z_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2z_smoke.log 2>&1; tail -2 $O/r2z_smoke.log
timeout 600 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err; cat $O/r2z_bench.json
timeout 400 python bench.py --config c3 > $O/r2z_bench_c3.json 2>> $O/r2z_bench.err; cat $O/r2z_bench_c3.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2z_bench_ref.json 2>> $O/r2z_bench.err; cat $O/r2z_bench_ref.json
tail -5 $O/r2z_bench.err
