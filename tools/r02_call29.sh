#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench n2 exit $?"
python -c "import json; d=json.loads(open('$O/r02_bench_n2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['parity']['ok'], d['e2e']['value'], d['value_l2_warm']); print(json.dumps(d.get('stokes'))[:500])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | cut -c1-300
