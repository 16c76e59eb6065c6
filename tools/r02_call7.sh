#!/bin/bash
# round 2, call 7 (1 GPU): KSP kernels (single-group multi-dot, descending MAXPY), P = 96 persistent chain, fused pressure pad, crop_sum
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ksp.py tests/test_gpu_elliptic.py tests/test_gpu_stokes.py tests/test_zz4_gpu_optins.py tests/test_golden.py tests/test_gpu_solvers.py -q > $O/r02c7_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r02c7_tests.log
timeout 400 python bench.py --child ksp > $O/r02c7_ksp.json 2> $O/r02c7_ksp.err; echo "ksp child exit $?"
python -c "import json; d=json.loads(open('$O/r02c7_ksp.json').read().strip().splitlines()[-1]); [print(k, json.dumps(v)) for k,v in d.items() if 'fgmres' in k or 'jacobi' in k]"
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02c7_time_stokes128.jsonl 2>&1; cat $O/r02c7_time_stokes128.jsonl
timeout 300 python bench.py --child p_sweep > $O/r02c7_p_sweep.json 2> $O/r02c7_p_sweep.err
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r02c7_p_sweep.json')):
    if r.get('op','').startswith('MatMult_Elliptic') or 'error' in r or 'truncated' in r: print(json.dumps(r))
PY
