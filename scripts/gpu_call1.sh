#!/bin/bash
# round-2 GPU session 1: validate the final code + packaging, A/B the masked-kernel modes, baseline numbers
set -x
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; free -g | head -2; nproc; python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))"; } > $O/r2a_box.txt 2>&1
python -c "from lettuce_b200 import build; print('source digest', build.source_digest()); print(open(build.STAMP).read())" >> $O/r2a_box.txt 2>&1
LBM_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2a_gpu_tests.log 2>&1
tail -30 $O/r2a_gpu_tests.log
for m in 1 3; do
  LBM_B200_MASKED_MODE=$m timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ebb.py tests/test_gpu_properties.py -m gpu -q --timeout 600 \
     -k "obstacle or ebb or known_answers or no_streaming or lazy or broadcast or mask" > $O/r2a_gpu_tests_mode$m.log 2>&1
  tail -5 $O/r2a_gpu_tests_mode$m.log
done
timeout 300 python __graft_entry__.py --smoke > $O/r2a_smoke.log 2>&1; tail -3 $O/r2a_smoke.log
timeout 900 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; cat $O/r2a_bench.json; tail -5 $O/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2a_bench_ref.json 2>> $O/r2a_bench.err; cat $O/r2a_bench_ref.json
timeout 600 python bench.py --config c3 --quick > $O/r2a_bench_c3.json 2>> $O/r2a_bench.err; cat $O/r2a_bench_c3.json
for m in 1 2 3; do
  LBM_B200_MASKED_MODE=$m timeout 600 python scripts/bench_configs.py c4 c5 --small > $O/r2a_masked_mode$m.jsonl 2>&1
  cat $O/r2a_masked_mode$m.jsonl
done
timeout 600 python scripts/bench_configs.py c1 extra > $O/r2a_configs.jsonl 2>&1; cat $O/r2a_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2a_launches.csv \
    python bench.py --size 256 --steps 3 --warmup 1 --quick --no-cpu > $O/r2a_ncu_bench.log 2>&1
tail -3 $O/r2a_ncu_bench.log
