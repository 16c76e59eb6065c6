#!/bin/bash
# round 2, call 10 (1 GPU): ncu evidence - launch list of the bench command, full captures of the persistent chain (128, 96), the
# even-odd kernel per P (48, 96, 128, 129) and the Stokes step; key metrics extracted on the box
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-ksp --no-stokes"
timeout 300 $B > $O/r02_bench_plain.json 2> $O/r02_bench_plain.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_ncu_bench.log 2>&1
echo "launch list exit $?"
for P in 128 96; do
  timeout 120 python tools/mm_once.py $P > $O/r02_mm_once_$P.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 4 -c 2 -f -o $O/r02_prof_persist_$P python tools/mm_once.py $P > $O/r02_ncu_persist_$P.log 2>&1
  python tools/ncu_keys.py $O/r02_prof_persist_$P.ncu-rep > $O/r02_persist_${P}_keys.txt 2>&1
done
: > $O/r02_perP_keys.txt
for P in 48 96 128 129; do for AX in 0 2; do
  timeout 120 python tools/cheb_once.py $P $AX > /dev/null 2>&1 && \
  timeout 600 ncu --set full --clock-control none -k regex:eo_deriv_kernel -s 2 -c 1 -f -o $O/r02_prof_eo_${P}_$AX python tools/cheb_once.py $P $AX > $O/r02_ncu_eo.log 2>&1
  echo "=== ChebMult P=$P axis=$AX" >> $O/r02_perP_keys.txt
  python tools/ncu_keys.py $O/r02_prof_eo_${P}_$AX.ncu-rep >> $O/r02_perP_keys.txt 2>&1
  [ "$P$AX" != "1280" ] && rm -f $O/r02_prof_eo_${P}_$AX.ncu-rep
done; done
timeout 120 python tools/stokes_once.py > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none -k regex:"eo_deriv|vv_flux|pad_|crop_|reduce_order" -s 9 -c 16 -f -o $O/r02_prof_stokes python tools/stokes_once.py > $O/r02_ncu_stokes_full.log 2>&1
python tools/ncu_keys.py $O/r02_prof_stokes.ncu-rep > $O/r02_stokes_keys.txt 2>&1
rm -f $O/r02_prof_stokes.ncu-rep
ls -la $O | tail -30
