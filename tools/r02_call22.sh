#!/bin/bash
# round 2, call 22 (2 GPUs): slab Stokes with the local axes on a side stream, trace fused into the flux kernel, fused pressure pad on the slab
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stokes_slab.py tests/test_zz5_gpu_saddle_slab.py tests/test_gpu_stokes.py -q -x > $O/r02c22_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r02c22_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29517 tools/stokes_slab_profile.py > $O/r02_stokes_slab_profile_n2_after.txt 2> $O/r02c22_prof.err; echo "profile exit $?"; cat $O/r02_stokes_slab_profile_n2_after.txt
for SS in 0 1; do
  SB200_STOKES_SIDE_STREAM=$SS timeout 300 $TR --master-port 29518 tests/dist/dist_stokes.py 24 128 2>/dev/null | grep -E "check|bench"
done
timeout 300 $TR --master-port 29519 tests/dist/dist_saddle.py 32 2>/dev/null | grep check
