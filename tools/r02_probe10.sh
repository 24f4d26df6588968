#!/bin/bash
# round-2 probe 10: 4 host cores per GPU (what the 8-GPU box offers: 32 cores) emulated with taskset on one GPU
mkdir -p gpurun_out
L=gpurun_out/r02_p10.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 100 taskset -c 0-3 python tools/probe_synth.py $SPEC 2>&1 | grep -v "B&B rounds\|stage " >> $L; echo "rc=$?" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=4 PROBE_WORKERS=4 MOIP_SYNC=spin
run PROBE_SPLIT=8 PROBE_WORKERS=8 MOIP_SYNC=block
run PROBE_SPLIT=12 PROBE_WORKERS=12 MOIP_SYNC=block
run PROBE_SPLIT=16 PROBE_WORKERS=16 MOIP_SYNC=block
run PROBE_SPLIT=24 PROBE_WORKERS=24 MOIP_SYNC=block
run PROBE_SPLIT=12 PROBE_WORKERS=12 MOIP_SYNC=spin
SPEC=kp:4:40
run PROBE_SPLIT=4 PROBE_WORKERS=4 MOIP_SYNC=spin
run PROBE_SPLIT=4 PROBE_WORKERS=12 MOIP_SYNC=block
run PROBE_SPLIT=4 PROBE_WORKERS=24 MOIP_SYNC=block
