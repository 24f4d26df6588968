#!/bin/bash
N=${1:-2}; shift
mkdir -p gpurun_out
nproc > gpurun_out/r02_mgp${N}.log; free -g | head -2 >> gpurun_out/r02_mgp${N}.log
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/probe_front_mr.py "$@" >> gpurun_out/r02_mgp${N}.log 2> gpurun_out/r02_mgp${N}.err
echo "rc=$?" >> gpurun_out/r02_mgp${N}.log
