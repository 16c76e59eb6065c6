#!/bin/bash
# usage: tools/slab_ablate.sh NGPUS  -- times the slab MatMult with parts of the exchange switched off (results invalid)
N=${1:-2}
for XF in 0 16 12 28 32 60; do
  echo "XFLAGS=$XF"
  SB200_XFLAGS=$XF timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['value_l2_warm'])"
done
