#!/bin/bash
# round 2, call 6 (1 GPU): KSP lookahead + fused Hessenberg update: parity, the whole GPU suite, KSP launch list, ksp child of the bench
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ksp.py tests/test_gpu_solvers.py tests/test_zz1_gpu_saddle.py -q > $O/r02c6_tests_ksp.log 2>&1; echo "ksp tests exit $?"; tail -4 $O/r02c6_tests_ksp.log
timeout 300 python tools/ksp_once.py 1 > $O/r02c6_plain_ksp.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c6_launches_ksp.csv python tools/ksp_once.py 1 > $O/r02c6_ncu_ksp.log 2>&1
tail -1 $O/r02c6_plain_ksp.log
timeout 400 python bench.py --child ksp > $O/r02c6_ksp.json 2> $O/r02c6_ksp.err; echo "ksp child exit $?"
python -c "import json; d=json.loads(open('$O/r02c6_ksp.json').read().strip().splitlines()[-1]); [print(k, json.dumps(v)) for k,v in d.items()]"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02c6_tests_all.log 2>&1; echo "all gpu tests exit $?"; tail -4 $O/r02c6_tests_all.log
