#!/bin/bash
# round 2, call 8 (2 GPUs): slab MatMult after the stage kernel's carve-out change: bench N=2 merged / two-phase, timelines of both
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for M in 1 0; do
  SB200_SLAB_MERGED=$M timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-stokes > $O/r02c8_bench_n2_merged$M.json 2> $O/r02c8_bench_n2_merged$M.err; echo "bench merged=$M exit $?"
  python -c "import json; d=json.loads(open('$O/r02c8_bench_n2_merged$M.json').read().strip().splitlines()[-1]); print('merged=$M', d['ms_per_step'], d['value'], d['config']['value_l2_warm'], d['parity']['ok'])"
  SB200_SLAB_MERGED=$M SB200_TL_EPOCH=35 SB200_ABLATE_LIB=1 SB200_XFLAGS=64 timeout 300 $TR --master-port 29513 tools/slab_timeline.py 128 > $O/r02c8_timeline_n2_merged$M.jsonl 2> $O/r02c8_timeline_n2_merged$M.err; echo "timeline merged=$M exit $?"
  cat $O/r02c8_timeline_n2_merged$M.jsonl
done
