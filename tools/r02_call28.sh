#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for W in 16 12 8; do echo "eo_warps=$W"; SB200_EO_WARPS=$W timeout 200 python tools/time_ops.py stokes 128 20 | head -2; done | tee $O/r02c28_eo_warps.log
