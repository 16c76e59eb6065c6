#!/bin/bash
# round 2, call 23 (2 GPUs): slab Stokes with producer-fused forward pushes and the peer-store epilogue of the pencil derivative
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stokes_slab.py tests/test_gpu_slab.py tests/test_zz5_gpu_saddle_slab.py tests/test_gpu_stokes.py tests/test_gpu_ksp.py -q -x > $O/r02c23_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r02c23_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29517 tools/stokes_slab_profile.py > $O/r02_stokes_slab_profile_n2_v3.txt 2> $O/r02c23_prof.err; echo "profile exit $?"; cat $O/r02_stokes_slab_profile_n2_v3.txt
timeout 300 $TR --master-port 29518 tests/dist/dist_stokes.py 24 128 2>/dev/null | grep -E "check|bench"
SB200_SLAB_PRODUCER_PUSH=0 SB200_SLAB_PEER_EPILOGUE=0 timeout 300 $TR --master-port 29518 tests/dist/dist_stokes.py 24 128 2>/dev/null | grep -E "check|bench"
timeout 300 $TR --master-port 29511 tests/dist/dist_check.py 32 64 2>/dev/null | grep check
timeout 300 $TR --master-port 29519 tests/dist/dist_saddle.py 32 2>/dev/null | grep check
