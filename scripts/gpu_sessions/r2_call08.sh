#!/bin/bash
# TMA-staged kernel: stages x CTAs-per-SM sweep
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2h_sweep.jsonl
for ctas in 1 2; do for st in 2 3 4 5 6 8; do
  echo "{\"sweep\": \"ctas=$ctas stages=$st\"}" >> $O/r2h_sweep.jsonl
  LBM_B200_TMA=1 LBM_B200_TMA_CTAS=$ctas LBM_B200_TMA_STAGES=$st timeout 300 python scripts/bench_configs.py c2 c3 c4 --small --pre-only >> $O/r2h_sweep.jsonl 2>&1
done; done
cat $O/r2h_sweep.jsonl
