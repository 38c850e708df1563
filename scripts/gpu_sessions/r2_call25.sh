#!/bin/bash
# round-2 GPU session 25: the reference's OWN generated CUDA kernel (use_native=True, prebuilt by
# baseline/build_native.py) next to the engine: parity test, bench line with the native_gpu_reference leg, ncu of it
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -q --timeout 200 -k "generated or reference_simulation" > $O/r2y_tests.log 2>&1; tail -n 6 $O/r2y_tests.log
timeout 400 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err; cat $O/r2y_bench.json; tail -n 3 $O/r2y_bench.err
for s in PRE_STREAMING POST_STREAMING; do
  timeout 120 python bench.py --leg native_gpu_reference --size 256 --strategy $s >> $O/r2y_native_256.jsonl 2>> $O/r2y_bench.err
done
cat $O/r2y_native_256.jsonl
timeout 300 ncu --set full --clock-control none -k regex:lettuce_kernel --launch-skip 3 --launch-count 1 -o $O/r2y_ref_native_256 -f \
  python bench.py --leg native_gpu_reference --size 256 > $O/r2y_ncu.log 2>&1; tail -n 3 $O/r2y_ncu.log
