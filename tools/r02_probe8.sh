#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "pool or steal or example" > gpurun_out/r02_p8_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p8_tests.log
MOIP_NO_STEAL=1 MOIP_WATCHDOG=6 PROBE_SPLIT=48 PROBE_WORKERS=12 timeout 30 python tools/probe_synth.py kp:4:40 > gpurun_out/r02_p8_hang.log 2>&1
echo "rc=$?" >> gpurun_out/r02_p8_hang.log
L=gpurun_out/r02_p8.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 100 python tools/probe_synth.py $SPEC >> $L 2>&1; echo "rc=$?" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=12 PROBE_WORKERS=12
run PROBE_SPLIT=24 PROBE_WORKERS=12
run PROBE_SPLIT=6 PROBE_WORKERS=12
run PROBE_SPLIT=16 PROBE_WORKERS=16
run PROBE_SPLIT=24 PROBE_WORKERS=12 MOIP_NO_STEAL=1
SPEC=kp:4:40
run PROBE_SPLIT=12 PROBE_WORKERS=12
run PROBE_SPLIT=24 PROBE_WORKERS=24
run PROBE_SPLIT=4 PROBE_WORKERS=16
run PROBE_SPLIT=24 PROBE_WORKERS=24 MOIP_NO_STEAL=1
