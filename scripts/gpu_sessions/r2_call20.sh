#!/bin/bash
# round-2 GPU session 20: compute-sanitizer on the TMA-staged kernel (single GPU)
set -x
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file $O/r2t_sanitizer_$tool.log \
      python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 800 -x \
      -k "tma_staged_kernel_is_bit_identical and (D3Q27 or D2Q9) or tma_staged_kernel_with_boundaries or fused_step_moments_equal_the_reductions and res5" \
      > $O/r2t_sanitize_$tool.out 2>&1
  tail -n 3 $O/r2t_sanitize_$tool.out
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/r2t_sanitizer_$tool.log | sort | uniq -c
done
