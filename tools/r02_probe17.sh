#!/bin/bash
# round-2 probe 17: boxes with postponement (node budget on a box's cold first subproblem), one process per configuration
mkdir -p gpurun_out
L=gpurun_out/r02_p17.log
echo "nproc $(nproc)" > $L
for spec in "$@"; do
  MOIP_WATCHDOG=20 timeout 70 python tools/probe_front_mr.py $spec >> $L 2> gpurun_out/r02_p17.err
  rc=$?
  if [ $rc -ne 0 ]; then echo "$spec rc=$rc" >> $L; grep watchdog gpurun_out/r02_p17.err | grep -v "strip -1" | tail -4 | cut -c1-250 >> $L; grep -v watchdog gpurun_out/r02_p17.err | tail -3 >> $L; fi
done
