#!/bin/bash
# round 2, call 5 (2 GPUs): TMA tensor-store pencil epilogue: emulated-rank parity, 2-process parity, bench N=2 with / without, timeline
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_stokes_slab.py tests/test_gpu_ksp.py -q > $O/r02c5_tests_emul.log 2>&1; echo "emulated slab tests exit $?"; tail -3 $O/r02c5_tests_emul.log
timeout 600 $TR --master-port 29511 tests/dist/dist_check.py 32 64 128 > $O/r02c5_dist_check_n2.jsonl 2> $O/r02c5_dist_check_n2.err; echo "dist_check exit $?"; cat $O/r02c5_dist_check_n2.jsonl
for B in 0 1; do
  SB200_SLAB_BULK=$B timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-stokes > $O/r02c5_bench_n2_bulk$B.json 2> $O/r02c5_bench_n2_bulk$B.err; echo "bench bulk=$B exit $?"
  python -c "import json; d=json.loads(open('$O/r02c5_bench_n2_bulk$B.json').read().strip().splitlines()[-1]); print('bulk=$B', d['ms_per_step'], d['value'], d['config']['value_l2_warm'], d['parity'], d['e2e']['value'])"
done
for B in 0 1; do
  SB200_SLAB_BULK=$B SB200_ABLATE_LIB=1 SB200_XFLAGS=64 timeout 300 $TR --master-port 29513 tools/slab_timeline.py 128 > $O/r02c5_timeline_n2_bulk$B.jsonl 2> $O/r02c5_timeline_n2_bulk$B.err; echo "timeline bulk=$B exit $?"
done
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02c5_bench_n2_full.json 2> $O/r02c5_bench_n2_full.err; echo "bench full exit $?"
python -c "import json; d=json.loads(open('$O/r02c5_bench_n2_full.json').read().strip().splitlines()[-1]); print(json.dumps(d.get('stokes')))"
timeout 300 python tools/time_ops.py stokes 128 10 > $O/r02c5_time_stokes128.jsonl 2>&1; cat $O/r02c5_time_stokes128.jsonl
