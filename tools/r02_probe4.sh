#!/bin/bash
# round-2 probe 4: GPU test suite on the new K2 / cache / point store, then the two headline fronts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_p4_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p4_tests.log
L=gpurun_out/r02_p4.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 150 python tools/probe_synth.py $SPEC 2>&1 | grep -v "^moip_b200:" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=32 PROBE_WORKERS=16 MOIP_KERNEL_TIMING=1
run PROBE_SPLIT=32 PROBE_WORKERS=16
run PROBE_SPLIT=24 PROBE_WORKERS=12
SPEC=kp:4:40
run PROBE_SPLIT=32 PROBE_WORKERS=16 MOIP_KERNEL_TIMING=1
run PROBE_SPLIT=32 PROBE_WORKERS=16
run PROBE_SPLIT=48 PROBE_WORKERS=24
echo done
