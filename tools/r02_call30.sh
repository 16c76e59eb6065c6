#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stokes.py tests/test_golden.py tests/test_gpu_solvers.py tests/test_zz3_gpu_drivers.py tests/test_gpu_stokes_slab.py -q -x > $O/r02c30_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02c30_tests.log
timeout 200 python tools/time_ops.py stokes 128 20 | tee $O/r02c30_time_stokes.log
