#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02_p14_tests.log
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -15 > $T; echo "pytest rc=$?" >> $T
L=gpurun_out/r02_p14.log; : > $L
echo "nproc $(nproc)" >> $L
run() { echo "== $*" >> $L; env "$@" timeout 120 python tools/probe_synth.py $SPEC 2>&1 | grep -v "watchdog\|Traceback\|File \|^    \|CUDA error" | head -8 >> $L; }
SPEC=kp:4:25
run MOIP_CHAIN_DEBUG=1 MOIP_CHAIN_STATS=1
run MOIP_CHAIN=0
SPEC=kp:4:40
run PROBE_SPLIT=4 PROBE_WORKERS=16 MOIP_CHAIN_STATS=1
run PROBE_SPLIT=4 PROBE_WORKERS=16 MOIP_CHAIN=0
run PROBE_SPLIT=4 PROBE_WORKERS=24
run PROBE_SPLIT=4 PROBE_WORKERS=24 MOIP_SYNC=block
SPEC=ap:3:30
run PROBE_SPLIT=24 PROBE_WORKERS=24 MOIP_SYNC=block
