#!/bin/bash
# round 2, call 1: recover the numbers round 1's truncated bench tail lost (per-P sweep, Stokes opt-ins), persist cfg variants
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/r02_bench_start.json 2> $O/r02_bench_start.err; echo "bench exit $?"
for cfg in 0 1 2 3; do
  for stg in 0 6000; do
    echo "{\"cfg\": $cfg, \"stagger\": $stg}" >> $O/r02_persist_cfgs.jsonl
    SB200_PERSIST_CFG=$cfg SB200_STAGGER=$stg timeout 120 python tools/time_ops.py elliptic 128 20 >> $O/r02_persist_cfgs.jsonl 2>&1
  done
done
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02_time_stokes128.jsonl 2>&1
tail -c 1500 $O/r02_bench_start.json
