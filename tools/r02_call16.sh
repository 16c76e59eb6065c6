#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for cfg in 3 1; do
  SB200_PERSIST_CFG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 4 -c 2 -f -o $O/r02_prof_persist_cfg$cfg python tools/mm_once.py 128 > $O/r02_ncu_persist_cfg$cfg.log 2>&1
  python tools/ncu_keys.py $O/r02_prof_persist_cfg$cfg.ncu-rep > $O/r02_persist_cfg${cfg}_keys.txt 2>&1
  grep -E "gpu__time_duration|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|registers|long_scoreboard|stall:wait|short_scoreboard|issue_active" $O/r02_persist_cfg${cfg}_keys.txt
done
