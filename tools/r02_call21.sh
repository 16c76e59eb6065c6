#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/stokes_slab_profile.py > $O/r02_stokes_slab_profile_n2.txt 2> $O/r02_stokes_slab_profile_n2.err; echo "exit $?"; cat $O/r02_stokes_slab_profile_n2.txt; tail -3 $O/r02_stokes_slab_profile_n2.err
