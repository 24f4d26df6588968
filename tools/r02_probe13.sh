#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02_p13_tests.log
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "synthetic or stealing or bruteforce" 2>&1 | tail -15 > $T; echo "pytest rc=$?" >> $T
L=gpurun_out/r02_p13.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 120 python tools/probe_synth.py $SPEC 2>&1 | grep -v "watchdog\|Traceback\|File \|^    \|CUDA error" | head -8 >> $L; }
SPEC=ap:3:17
run MOIP_CHAIN_DEBUG=1
SPEC=ap:3:20
run MOIP_CHAIN_STATS=1
run MOIP_CHAIN=0
SPEC=ap:3:30
run PROBE_SPLIT=16 PROBE_WORKERS=16
run PROBE_SPLIT=16 PROBE_WORKERS=16 MOIP_CHAIN=0
run PROBE_SPLIT=24 PROBE_WORKERS=24
run PROBE_SPLIT=32 PROBE_WORKERS=32
run PROBE_SPLIT=24 PROBE_WORKERS=24 MOIP_SYNC=block
