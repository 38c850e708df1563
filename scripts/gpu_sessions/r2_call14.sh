#!/bin/bash
# round-2 GPU session 14 (2 GPUs): slabs with the staged kernel on the interior planes
set -x
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_slab.py -m gpu -q --timeout 800 > $O/r2n_slab_tests.log 2>&1; tail -30 $O/r2n_slab_tests.log
timeout 900 $TR --master-port 29517 bench.py --gpus 2 --config c3 --no-e2e > $O/r2n_bench_c3_n2_weak.json 2> $O/r2n_bench.err; cat $O/r2n_bench_c3_n2_weak.json
timeout 900 $TR --master-port 29513 scripts/bench_multi.py c3 --energy > $O/r2n_c3_n2.json 2>> $O/r2n_bench.err; cat $O/r2n_c3_n2.json
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 30 > $O/r2n_bench_n2_512.json 2>> $O/r2n_bench.err; cat $O/r2n_bench_n2_512.json
tail -5 $O/r2n_bench.err
for tool in memcheck racecheck; do
  LBM_B200_TMA=1 timeout 600 compute-sanitizer --tool $tool --target-processes all --log-file $O/r2n_sanitizer_$tool.%p.log \
      $TR --master-port 2952${#tool} scripts/slab_sanitize.py > $O/r2n_sanitize_$tool.out 2>&1
  tail -n 4 $O/r2n_sanitize_$tool.out
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/r2n_sanitizer_$tool.*.log | sort | uniq -c
done
