#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_p18.log
echo "nproc $(nproc)" > $L
MOIP_EXPLODE_LOG=100000 MOIP_STRIP_TIMELINE=1 timeout 45 python tools/probe_front_mr.py ap3_30_1:192:24:MOIP_WINDOWS=8 >> $L 2> gpurun_out/r02_p18.err
echo "rc=$?" >> $L
grep "IP past" gpurun_out/r02_p18.err | head -20 >> $L
