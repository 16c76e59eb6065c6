#!/bin/bash
# round 2, call 25 (1 GPU): single-GPU Stokes with the pressure preparation on the side stream
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stokes.py tests/test_zz4_gpu_optins.py tests/test_zz1_gpu_saddle.py tests/test_gpu_solvers.py tests/test_golden.py tests/test_gpu_stokes_slab.py -q -x > $O/r02c25_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02c25_tests.log
for SS in 1 0; do echo "side_stream=$SS"; SB200_STOKES_SIDE_STREAM=$SS timeout 200 python tools/time_ops.py stokes 128 20; done | tee $O/r02c25_time_stokes.log
for SS in 1 0; do echo "side_stream=$SS"; SB200_STOKES_SIDE_STREAM=$SS timeout 200 python tools/time_ops.py stokes 32 20; done | tee -a $O/r02c25_time_stokes.log
