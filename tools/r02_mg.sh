#!/bin/bash
# N-rank run (gpurun --gpus N): the 2-rank GPU tests over NCCL (N >= 2), then the bench line as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mg${N}_smi.log 2>&1
if [ "$2" = "tests" ]; then
  timeout 600 python -m pytest tests/test_multi_rank_gpu.py -x -q -m gpu > gpurun_out/r02_mg${N}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_mg${N}_tests.log
fi
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_mg${N}_bench.json 2> gpurun_out/r02_mg${N}_bench.err
echo "bench rc=$?" >> gpurun_out/r02_mg${N}_bench.err
