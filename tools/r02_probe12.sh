#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_p12.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 120 python tools/probe_synth.py $SPEC 2>&1 | grep -v "watchdog\|Traceback\|File \|^    \|^  stage" | head -12 >> $L; }
SPEC=ap:3:20
run MOIP_CHAIN_DEBUG=1
run MOIP_CHAIN_DEBUG=1 MOIP_BB_LEVELS=1
run MOIP_CHAIN_DEBUG=1 MOIP_K1_OCC=1
SPEC=ap:3:16
run MOIP_CHAIN_DEBUG=1
SPEC=ap:3:17
run MOIP_CHAIN_DEBUG=1
