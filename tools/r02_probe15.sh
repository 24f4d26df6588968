#!/bin/bash
# round-2 probe 15: boxes (strips x windows) on one GPU -- the IP count of a 192-way cut without needing 8 GPUs
mkdir -p gpurun_out
L=gpurun_out/r02_p15.log
echo "nproc $(nproc)" > $L
timeout 900 python tools/probe_front_mr.py ap3_30_1:24:24 ap3_30_1:192:24:MOIP_WINDOWS=1 ap3_30_1:192:24:MOIP_WINDOWS=8 ap3_30_1:192:24:MOIP_WINDOWS=16 ap3_30_1:96:24:MOIP_WINDOWS=8 ap3_30_1:48:24:MOIP_WINDOWS=4 \
  kp4_40_1:4:24 kp4_40_1:32:24:MOIP_WINDOWS=1 kp4_40_1:32:24:MOIP_WINDOWS=4 kp4_40_1:96:24:MOIP_WINDOWS=8 kp4_40_1:192:24:MOIP_WINDOWS=16 >> $L 2> gpurun_out/r02_p15.err
echo "rc=$?" >> $L
tail -5 gpurun_out/r02_p15.err >> $L
