#!/bin/bash
# round-2 GPU session 16: automatic lazy POST batches for the staged kernel, reporter overhead, D2Q9 KBC policy
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 > $O/r2p_parity.log 2>&1; tail -n 6 $O/r2p_parity.log
timeout 600 python scripts/bench_configs.py c3 > $O/r2p_c3.jsonl 2>&1; cat $O/r2p_c3.jsonl
for tma in 0 1; do LBM_B200_TMA=$tma timeout 300 python scripts/bench_configs.py kbc2d > $O/r2p_kbc2d_tma$tma.jsonl 2>&1; cat $O/r2p_kbc2d_tma$tma.jsonl; done
timeout 900 python scripts/bench_reporter_overhead.py > $O/r2p_reporter_overhead.jsonl 2>&1; cat $O/r2p_reporter_overhead.jsonl
