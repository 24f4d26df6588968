#!/bin/bash
# round-2 probe 2: do the short helper kernels (K2/K4, one CTA per node) cost K1 capacity?  3AP n=30 front, 24 strips, 12 contexts
mkdir -p gpurun_out
L=gpurun_out/r02_p2.log; : > $L
run() { echo "== $*" >> $L; env "$@" PROBE_SPLIT=24 PROBE_WORKERS=12 timeout 120 python tools/probe_synth.py ap:3:30 >> $L 2>&1; }
run MOIP_KERNEL_TIMING=1
run MOIP_X=0
run MOIP_AUX_GRID=32
run MOIP_AUX_GRID=12
run MOIP_AUX_GRID=4
run MOIP_AUX_CARVEOUT=52
run MOIP_AUX_CARVEOUT=100
run MOIP_AUX_CARVEOUT=52 MOIP_AUX_GRID=12
echo "== kp40" >> $L
MOIP_KERNEL_TIMING=1 PROBE_SPLIT=24 PROBE_WORKERS=12 timeout 120 python tools/probe_synth.py kp:4:40 >> $L 2>&1
MOIP_AUX_GRID=12 PROBE_SPLIT=24 PROBE_WORKERS=12 timeout 120 python tools/probe_synth.py kp:4:40 >> $L 2>&1
echo done
