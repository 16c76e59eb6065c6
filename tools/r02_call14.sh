#!/bin/bash
# round 2, call 14 (1 GPU): persistent chain with the next item's load issued before the epilogue + L2 prefetch of the partial sums; cfg sweep
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_elliptic.py tests/test_golden.py -q > $O/r02c14_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02c14_tests.log
for P in 128 96 64; do timeout 120 python tools/time_ops.py elliptic $P 30 2>&1 | head -1; done | tee $O/r02c14_time_elliptic.jsonl
for cfg in 0 1 2 3; do for stg in 0 6000; do
  echo "cfg=$cfg stagger=$stg $(SB200_PERSIST_CFG=$cfg SB200_STAGGER=$stg timeout 120 python tools/time_ops.py elliptic 128 30 2>&1 | head -1)"
done; done | tee $O/r02c14_persist_cfgs.log
