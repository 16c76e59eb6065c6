#!/bin/bash
# Everything that round 1 left unmeasured, in ONE GPU call (expect 25-40 box minutes; the per-step timeouts add up to more, so give
# the call room and read gpurun_out/r02_steps.log to see how far it got):
#   gpurun --timeout 3000 -- 'bash tools/round2_first_call.sh'
# Steps 1-2 alone (tests + bench, ~12 minutes) are the part that must not be skipped: STEPS=2 bash tools/round2_first_call.sh
# Outputs land in gpurun_out/ (r02_*): copy the summaries worth keeping into profiles/.
# Order: the things whose numbers matter most first, so a call cut short still leaves them behind; every step is time-boxed and
# a failing step does not stop the rest.
set -u
mkdir -p gpurun_out
O=gpurun_out
step() { echo "=== $1" | tee -a $O/r02_steps.log; shift; ( "$@" ) >> $O/r02_steps.log 2>&1; echo "    exit $?" | tee -a $O/r02_steps.log; }

# 1. late GPU tests first (never run on a GPU in round 1), then the whole suite
step "late GPU tests" timeout 900 python -m pytest tests/test_zz1_gpu_saddle.py tests/test_zz2_gpu_reference_api2.py tests/test_zz3_gpu_drivers.py tests/test_zz4_gpu_optins.py -q
step "full GPU suite" timeout 1200 python -m pytest tests -m gpu -x -q

# 2. the bench line (with the per-P sweep and the KSP metric in child processes) and the reference arm
timeout 600 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench exit $?" >> $O/r02_steps.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_ref.json 2>> $O/r02_steps.log

[ "${STEPS:-9}" -le 2 ] && { tail -3 $O/r02_steps.log; exit 0; }
# 3. per-launch time lists (never a bench value): the Stokes step and the native saddle-point PC
timeout 300 python tools/stokes_once.py > $O/r02_plain_stokes.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_stokes.csv python tools/stokes_once.py > $O/r02_ncu_stokes.log 2>&1
timeout 300 python tools/saddle_once.py > $O/r02_plain_saddle.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches_saddle.csv python tools/saddle_once.py > $O/r02_ncu_saddle.log 2>&1

# 4. one full capture of the Stokes kernels (the 52 % item of round 1): the 14 launches of the SECOND StokesMatMult of
#    tools/stokes_once.py (of the kernels matching the filter, StokesFunction launches 13 and the first StokesMatMult 14: skipped)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'eo_deriv|vv_flux|pad_nodes|crop|reduce_order' -s 27 -c 14 -o $O/r02_prof_stokes python tools/stokes_once.py > $O/r02_ncu_full_stokes.log 2>&1
# 5. the device assembly of the preconditioning matrices at 128^3 (not timed in round 1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fd_assemble' -c 4 -o $O/r02_prof_fd python tools/p_sweep.py 1 > $O/r02_ncu_full_fd.log 2>&1
# 6. a Stokes linear solve at 128^3 with nothing leaving the device (Jacobi on MatVVPC applied on the device): wall time of the
#    native executable, outer iteration count; 20^3 with the host ILU(2) stand-in beside it
( time timeout 900 apps/stokes -exact 2 -cont0 1 -dim 128,128,128 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_rtol 1e-6 -ksp_max_it 60 \
    -vel_pc_type jacobi -svel_pc_type jacobi -ksp_monitor ) > $O/r02_stokes128_device_jacobi.log 2>&1
( time timeout 900 apps/stokes -exact 2 -cont0 1 -dim 128,128,128 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_rtol 1e-6 -ksp_max_it 60 \
    -vel_pc_type jacobi -svel_pc_type jacobi -ksp_monitor -sb200_trace_divergence 1 -sb200_fold_pressure 1 ) > $O/r02_stokes128_device_jacobi_switches.log 2>&1
( time timeout 300 apps/stokes -exact 2 -cont0 1 -dim 20,20,20 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_rtol 1e-10 -ksp_max_it 400 \
    -vel_pc_factor_levels 2 -svel_pc_factor_levels 2 -ksp_monitor ) > $O/r02_stokes20_ilu2.log 2>&1
( time timeout 300 apps/stokes -exact 2 -cont0 1 -dim 20,20,20 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_rtol 1e-10 -ksp_max_it 400 \
    -vel_pc_factor_levels 2 -svel_pc_factor_levels 2 -ksp_monitor -sb200_graph 1 ) > $O/r02_stokes20_ilu2_graph.log 2>&1
( time timeout 600 apps/elliptic -dim 128,128,128 -exact 2 -ksp_rtol 1e-10 -ksp_max_it 2000 -pc_type jacobi -ksp_monitor ) > $O/r02_elliptic128_device_jacobi.log 2>&1
# 7. (only on a multi-GPU call: gpurun --gpus 2|4|8) the slab-partitioned Stokes shells with and without the two evaluation switches
NG=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$NG" -ge 2 ]; then
  for opts in "" "trace,fold"; do
    SB200_STOKES_OPTS=$opts timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 \
      tests/dist/dist_stokes.py 24 128 >> $O/r02_stokes_slab_n$NG.jsonl 2>> $O/r02_steps.log
  done
fi
# 8. memcheck of the kernels written after round 1's last GPU run (vecops, crop_trace, the FOLD variants) at small sizes
step "memcheck saddle" timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/saddle_once.py 16
step "memcheck stokes switches" timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_zz4_gpu_optins.py -x -q -k "16"
tail -3 $O/r02_steps.log
