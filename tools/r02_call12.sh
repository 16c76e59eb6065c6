#!/bin/bash
# round 2, call 12 (1 GPU): slab saddle solve tests (emulated ranks; the multi-process script at world 1), KSP with PDL-chained vector kernels
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_zz5_gpu_saddle_slab.py tests/test_gpu_ksp.py -q > $O/r02c12_tests.log 2>&1; echo "tests exit $?"; tail -12 $O/r02c12_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 tests/dist/dist_saddle.py 32 > $O/r02c12_dist_saddle_n1.jsonl 2> $O/r02c12_dist_saddle_n1.err; echo "dist_saddle world 1 exit $?"; cat $O/r02c12_dist_saddle_n1.jsonl; tail -3 $O/r02c12_dist_saddle_n1.err
timeout 400 python bench.py --child ksp > $O/r02c12_ksp.json 2> $O/r02c12_ksp.err; echo "ksp child exit $?"
python -c "import json; d=json.loads(open('$O/r02c12_ksp.json').read().strip().splitlines()[-1]); [print(k, json.dumps(v)) for k,v in d.items() if 'fgmres' in k or 'jacobi' in k or 'config1' in k]"
