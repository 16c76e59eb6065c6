#!/bin/bash
# round 2, call 19 (1 GPU): the K-split GEMM variant (diagnostic library built with -DSB200_KSPLIT=2) against the production library
set -u
O=gpurun_out; mkdir -p $O
SB200_ABLATE_LIB=1 timeout 600 python -m pytest tests/test_gpu_elliptic.py -q > $O/r02c19_tests_ks.log 2>&1; echo "tests (ksplit lib) exit $?"; tail -2 $O/r02c19_tests_ks.log
for lib in 0 1; do for cfg in 1 2; do for P in 128 96 64; do
  echo "ksplit_lib=$lib cfg=$cfg $(SB200_ABLATE_LIB=$lib SB200_PERSIST_CFG=$cfg timeout 120 python tools/time_ops.py elliptic $P 40 2>&1 | head -1)"
done; done; done | tee $O/r02c19_ksplit.log
