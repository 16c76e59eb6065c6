#!/bin/bash
# round 2, call 17 (8 GPUs): the round-2 kernels at 4 and 8 real processes: bench lines (elliptic + Stokes, parity fields), slab saddle solve, slab parity
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench n8 exit $?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench n4 exit $?"
for f in n8 n4; do python -c "import json; d=json.loads(open('$O/r02_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['parity']['ok'], d['parity']['rel'], d['e2e']['value']); print(json.dumps(d.get('stokes'))[:600])"; done
timeout 300 $TR --nproc-per-node 8 --master-port 29514 tests/dist/dist_check.py 32 64 128 > $O/r02_dist_check_n8.jsonl 2> $O/r02_dist_check_n8.err; echo "dist_check n8 exit $?"; cat $O/r02_dist_check_n8.jsonl
timeout 300 $TR --nproc-per-node 8 --master-port 29515 tests/dist/dist_saddle.py 16 > $O/r02_dist_saddle_n8.jsonl 2> $O/r02_dist_saddle_n8.err; echo "dist_saddle n8 exit $?"; cat $O/r02_dist_saddle_n8.jsonl
CUDA_VISIBLE_DEVICES=0,1 timeout 300 $TR --nproc-per-node 2 --master-port 29516 tests/dist/dist_saddle.py 32 > $O/r02_dist_saddle_n2.jsonl 2> $O/r02_dist_saddle_n2.err; echo "dist_saddle n2 exit $?"; cat $O/r02_dist_saddle_n2.jsonl
tail -n 3 $O/r02_bench_n8.err
