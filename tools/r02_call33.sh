#!/bin/bash
# round 2, call 33 (1 GPU): persistent chain at every P % 16 == 0 up to 160 (short last tile batches guarded)
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_elliptic.py tests/test_golden.py -q > $O/r02c33_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r02c33_tests.log
for P in 48 80 96 112 128 144 160; do timeout 120 python tools/time_ops.py elliptic $P 20 2>/dev/null | head -1; done | tee $O/r02c33_time_elliptic.jsonl
