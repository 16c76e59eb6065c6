#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for CH in 8 16 32; do for D in 0 1; do
  echo "CH=$CH DESC=$D: $(SB200_KSP_MDOT_CH=$CH SB200_KSP_DESC=$D timeout 120 python tools/ksp_once.py 1 | tail -1)" | tee -a $O/r02c9_ksp_variants.log
done; done
