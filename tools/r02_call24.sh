#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stokes_slab.py tests/test_gpu_slab.py tests/test_zz5_gpu_saddle_slab.py -q -x > $O/r02c24_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02c24_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for PE in 1 0; do SB200_SLAB_PEER_EPILOGUE=$PE timeout 300 $TR --master-port 29518 tests/dist/dist_stokes.py 24 128 2>/dev/null | grep -E "check|bench" | cut -c1-230; done
timeout 300 $TR --master-port 29519 tests/dist/dist_saddle.py 32 2>/dev/null | grep check | cut -c1-200
