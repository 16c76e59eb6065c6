#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stokes.py tests/test_golden.py tests/test_zz4_gpu_optins.py -q -x -k "not full_size" > $O/r02c32_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02c32_tests.log
timeout 200 python tools/time_ops.py stokes 128 20 | tee $O/r02c32_time_stokes.log
