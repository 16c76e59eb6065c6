#!/bin/bash
# round 2, call 2 (1 GPU): generalised even-odd kernel + fused Stokes scatters: parity first, memcheck of the new modes, then timings
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_stokes.py tests/test_zz4_gpu_optins.py -x -q > $O/r02c2_tests_a.log 2>&1; echo "tests A exit $?"; tail -4 $O/r02c2_tests_a.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_stokes.py -x -q -k "9, 7, 6 or 8, 6 or 33" > $O/r02c2_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 $O/r02c2_memcheck.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_stokes.py::test_full_size_128 > $O/r02c2_tests_all.log 2>&1; echo "all gpu tests exit $?"; tail -4 $O/r02c2_tests_all.log
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02c2_time_stokes128.jsonl 2>&1
timeout 200 python bench.py --child p_sweep > $O/r02c2_p_sweep.json 2> $O/r02c2_p_sweep.err
timeout 300 python tools/stokes_once.py > $O/r02c2_plain_stokes.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c2_launches_stokes.csv python tools/stokes_once.py > $O/r02c2_ncu_stokes.log 2>&1
cat $O/r02c2_time_stokes128.jsonl
