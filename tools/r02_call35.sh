#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-ksp --no-stokes"
timeout 200 $B > $O/r02_bench_plain.json 2> $O/r02_bench_plain.err && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_ncu_bench.log 2>&1
echo "launch list exit $?"; grep -c persist_kernel $O/r02_launches_bench.csv
