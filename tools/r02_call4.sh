#!/bin/bash
# round 2, call 4 (1 GPU): axis-0 finishing job with 16-byte term loads; parity, timings, the bench line with the stokes object
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cheb.py tests/test_gpu_elliptic.py tests/test_golden.py tests/test_gpu_stokes.py tests/test_zz4_gpu_optins.py tests/test_zz1_gpu_saddle.py tests/test_gpu_solvers.py -q > $O/r02c4_tests_a.log 2>&1; echo "tests A exit $?"; tail -6 $O/r02c4_tests_a.log
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02c4_time_stokes128.jsonl 2>&1
timeout 300 python tools/stokes_once.py > $O/r02c4_plain_stokes.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c4_launches_stokes.csv python tools/stokes_once.py > $O/r02c4_ncu_stokes.log 2>&1
timeout 900 python bench.py --no-cpu-baseline --no-ksp > $O/r02c4_bench.json 2> $O/r02c4_bench.err; echo "bench exit $?"
cat $O/r02c4_time_stokes128.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c4_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('stokes')))
print(d['value'], d['parity'])
PY
