#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for cfg in 1 2; do for stg in 3000 9000 12000 16000 24000; do
  echo "cfg=$cfg stagger=$stg $(SB200_PERSIST_CFG=$cfg SB200_STAGGER=$stg timeout 120 python tools/time_ops.py elliptic 128 40 2>&1 | head -1)"
done; done | tee $O/r02c15_persist_stagger.log
