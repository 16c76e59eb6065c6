#!/bin/bash
# round 2, call 1 (1 GPU): the whole GPU suite at HEAD, the bench line (with p_sweep / ksp children), Stokes timings and launch list
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/r02_gputests.log
timeout 900 python bench.py > $O/r02_bench_start.json 2> $O/r02_bench_start.err; echo "bench exit $?"
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02_time_stokes128.jsonl 2>&1
timeout 300 python tools/time_ops.py elliptic 128 20 > $O/r02_time_elliptic128.jsonl 2>&1
timeout 300 python tools/stokes_once.py > $O/r02_plain_stokes.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_stokes.csv python tools/stokes_once.py > $O/r02_ncu_stokes.log 2>&1
cat $O/r02_time_stokes128.jsonl $O/r02_time_elliptic128.jsonl
head -c 1500 $O/r02_bench_start.json
