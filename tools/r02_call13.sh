#!/bin/bash
# round 2, call 13 (2 GPUs): slab saddle solve - emulated test, multi-process suite at world 2, the script at 32^3
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_zz5_gpu_saddle_slab.py -q > $O/r02c13_tests_saddle.log 2>&1; echo "saddle tests exit $?"; tail -5 $O/r02c13_tests_saddle.log
timeout 900 python -m pytest tests/test_dist_multi.py -q > $O/r02c13_tests_dist.log 2>&1; echo "dist tests exit $?"; tail -5 $O/r02c13_tests_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist/dist_saddle.py 32 > $O/r02c13_dist_saddle_n2.jsonl 2> $O/r02c13_dist_saddle_n2.err; echo "dist_saddle exit $?"; cat $O/r02c13_dist_saddle_n2.jsonl; tail -3 $O/r02c13_dist_saddle_n2.err
