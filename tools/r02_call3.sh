#!/bin/bash
# round 2, call 3 (1 GPU): generalised even-odd kernel + fused scatters (Stokes and the generic elliptic path): parity, then timings
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_elliptic.py tests/test_golden.py tests/test_gpu_stokes.py tests/test_zz4_gpu_optins.py -q > $O/r02c3_tests_a.log 2>&1; echo "tests A exit $?"; tail -6 $O/r02c3_tests_a.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_stokes.py::test_full_size_128 > $O/r02c3_tests_all.log 2>&1; echo "all gpu tests exit $?"; tail -4 $O/r02c3_tests_all.log
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02c3_time_stokes128.jsonl 2>&1
timeout 200 python bench.py --child p_sweep > $O/r02c3_p_sweep.json 2> $O/r02c3_p_sweep.err
timeout 300 python tools/stokes_once.py > $O/r02c3_plain_stokes.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c3_launches_stokes.csv python tools/stokes_once.py > $O/r02c3_ncu_stokes.log 2>&1
cat $O/r02c3_time_stokes128.jsonl
