#!/bin/bash
# round-2 GPU session 2: new unified step kernel (packed two-node lanes, fused moments, PDL chaining)
set -x
mkdir -p gpurun_out
O=gpurun_out
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" > $O/r2b_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/r2b_gpu_tests.log 2>&1
tail -40 $O/r2b_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2b_smoke.log 2>&1; tail -3 $O/r2b_smoke.log
for lanes in 1 2; do
  LBM_B200_LANES=$lanes timeout 900 python scripts/bench_configs.py c2 c3 c4 c5 extra --small > $O/r2b_lanes$lanes.jsonl 2>&1
  cat $O/r2b_lanes$lanes.jsonl
done
LBM_B200_PDL=0 timeout 600 python scripts/bench_configs.py c1 c4 > $O/r2b_nopdl.jsonl 2>&1; cat $O/r2b_nopdl.jsonl
timeout 600 python scripts/bench_configs.py c1 > $O/r2b_c1.jsonl 2>&1; cat $O/r2b_c1.jsonl
timeout 900 python bench.py > $O/r2b_bench.json 2> $O/r2b_bench.err; cat $O/r2b_bench.json; tail -5 $O/r2b_bench.err
timeout 600 python bench.py --config c3 --quick --no-cpu > $O/r2b_bench_c3.json 2>> $O/r2b_bench.err; cat $O/r2b_bench_c3.json
# ncu: full section set of the KBC step kernel (two nodes per thread, PRE and POST at 256^3) and of the masked D2Q9 step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 6 --launch-count 1 \
    -o $O/r2b_kbc_lanes2 -f python scripts/bench_configs.py c3 --small > $O/r2b_ncu_kbc.log 2>&1
LBM_B200_LANES=1 timeout 900 ncu --set full --clock-control none -k regex:step_kernel --launch-skip 6 --launch-count 1 \
    -o $O/r2b_kbc_lanes1 -f python scripts/bench_configs.py c3 --small > $O/r2b_ncu_kbc1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:step_kernel --launch-skip 30 --launch-count 1 \
    -o $O/r2b_c4 -f python scripts/bench_configs.py c4 > $O/r2b_ncu_c4.log 2>&1
ls -la $O/*.ncu-rep
