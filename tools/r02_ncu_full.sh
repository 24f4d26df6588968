#!/bin/bash
# one ncu --set full capture of the headline kernel (K1, plain batch instantiation) of the default bench command
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1_reg -s 3 -c 1 -f -o gpurun_out/r02_k1_reg_final python bench.py --no-fronts --cpu-sample 16 --steps 2 --warmup 3 > gpurun_out/r02_ncu_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02_ncu_full.log
