#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for B in 2 4 8 16; do
  echo "blocks_per_sm=$B: $(SB200_KSP_BLOCKS_PER_SM=$B timeout 120 python tools/ksp_once.py 1 | tail -1)" | tee -a $O/r02c20_ksp_blocks.log
  echo "blocks_per_sm=$B (2nd): $(SB200_KSP_BLOCKS_PER_SM=$B timeout 120 python tools/ksp_once.py 1 | tail -1)" | tee -a $O/r02c20_ksp_blocks.log
done
