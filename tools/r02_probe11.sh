#!/bin/bash
# round-2 probe 11: chained B&B rounds (bbchain.h) -- correctness on the synthetic goldens, then time-to-front against the host loop
mkdir -p gpurun_out
T=gpurun_out/r02_p11_tests.log
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "synthetic or stealing or bruteforce" > $T 2>&1; echo "pytest rc=$?" >> $T
L=gpurun_out/r02_p11.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 120 python tools/probe_synth.py $SPEC 2>&1 | grep -v watchdog >> $L; echo "rc=$?" >> $L; }
SPEC=ap:3:12
run MOIP_CHAIN_STATS=1
run MOIP_CHAIN=0
SPEC=ap:3:30
run PROBE_SPLIT=16 PROBE_WORKERS=16 MOIP_CHAIN_STATS=1
run PROBE_SPLIT=16 PROBE_WORKERS=16 MOIP_CHAIN=0
run PROBE_SPLIT=24 PROBE_WORKERS=24
run PROBE_SPLIT=32 PROBE_WORKERS=32
