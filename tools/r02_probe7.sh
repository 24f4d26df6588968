#!/bin/bash
mkdir -p gpurun_out
MOIP_DEBUG_GEN=1 PROBE_SPLIT=48 PROBE_WORKERS=12 timeout 45 python tools/probe_synth.py kp:4:40 > gpurun_out/r02_p7_hang.log 2>&1
echo "rc=$?" >> gpurun_out/r02_p7_hang.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_p7_bench.json 2> gpurun_out/r02_p7_bench.err
echo "bench rc=$?" >> gpurun_out/r02_p7_bench.err
