#!/bin/bash
# round-2 probe 16: why the 192-box run of probe 15 did not finish -- watchdog on a short run
mkdir -p gpurun_out
L=gpurun_out/r02_p16.log
echo "nproc $(nproc)" > $L
MOIP_WATCHDOG=12 MOIP_CHAIN_STATS=1 timeout 50 python tools/probe_front_mr.py ap3_30_1:192:24:MOIP_WINDOWS=8 >> $L 2> gpurun_out/r02_p16.err
echo "rc=$?" >> $L
grep -c watchdog gpurun_out/r02_p16.err >> $L
grep "watchdog" gpurun_out/r02_p16.err | tail -24 >> $L
grep -v "watchdog" gpurun_out/r02_p16.err | tail -5 >> $L
