#!/bin/bash
# round-2 GPU session 4 (re-entry): validate the current code end to end, then lanes A/B and the launch list
set -x
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; free -g | head -2; nproc; python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))"; } > $O/r2d_box.txt 2>&1
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print('stamp        ', open(build.STAMP).read())" >> $O/r2d_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2d_gpu_tests.log 2>&1
tail -40 $O/r2d_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $O/r2d_smoke.log 2>&1; tail -3 $O/r2d_smoke.log
timeout 900 python bench.py > $O/r2d_bench.json 2> $O/r2d_bench.err; cat $O/r2d_bench.json; tail -5 $O/r2d_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2d_bench_ref.json 2>> $O/r2d_bench.err; cat $O/r2d_bench_ref.json
timeout 600 python bench.py --config c3 --quick --no-cpu > $O/r2d_bench_c3.json 2>> $O/r2d_bench.err; cat $O/r2d_bench_c3.json
for lanes in 1 2; do
  LBM_B200_LANES=$lanes timeout 900 python scripts/bench_configs.py c2 c3 c4 c5 extra --small > $O/r2d_lanes$lanes.jsonl 2>&1
  cat $O/r2d_lanes$lanes.jsonl
done
timeout 600 python scripts/bench_configs.py c1 > $O/r2d_c1.jsonl 2>&1; cat $O/r2d_c1.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2d_launches.csv \
    python bench.py --size 256 --steps 3 --warmup 1 --quick --no-cpu > $O/r2d_ncu_bench.log 2>&1
tail -3 $O/r2d_ncu_bench.log
