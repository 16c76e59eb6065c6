#!/bin/bash
# round 2, call 18 (1 GPU): the whole GPU suite and the full bench line on the current code
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_gputests_final.log 2>&1; echo "gpu tests exit $?"; tail -4 $O/r02_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r02_smoke.log
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['parity']['ok'], d['roofline']['frac'], d['roofline']['peak'], d['e2e']['value'], d['cpu_baseline']['value'])
print(json.dumps(d['stokes'])[:400])
print(json.dumps(d['ksp'].get('fgmres30_cycle_128')))
print(len(d['p_sweep']), [r for r in d['p_sweep'] if 'error' in r or 'truncated' in r])
r=json.loads(open('gpurun_out/r02_bench_ref.json').read().strip().splitlines()[-1]); print(r['value'], r['config']==d['config'] or 'config differs (value_l2_warm key only?)')
PY
