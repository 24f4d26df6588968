#!/bin/bash
# round-2 ncu evidence (one B200; run AFTER the same commands have exited 0 without ncu):
#  1. launch list of the bench command (gpu__time_duration per launch)          -> gpurun_out/r02_launches_bench.csv
#  2. one --set full capture of the batch K1 kernel in that command              -> gpurun_out/r02_k1_reg_batch.ncu-rep
#  3. launch list of a front (3AP n=15, one context: fused rounds, K3 scans)     -> gpurun_out/r02_launches_front_ap15.csv
#  4. one --set full capture of the FUSED K1 instantiation during that front     -> gpurun_out/r02_k1_reg_fused.ncu-rep
mkdir -p gpurun_out
B="python bench.py --no-fronts --cpu-sample 16 --steps 2 --warmup 3"
timeout 300 $B > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_prof_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_reg -s 3 -c 1 -f -o gpurun_out/r02_k1_reg_batch $B > gpurun_out/r02_prof_ncu2.log 2>&1
timeout 120 python tools/probe_synth.py ap:3:15 > gpurun_out/r02_prof_front_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_front_ap15.csv python tools/probe_synth.py ap:3:15 > gpurun_out/r02_prof_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_reg -s 2000 -c 1 -f -o gpurun_out/r02_k1_reg_fused python tools/probe_synth.py ap:3:15 > gpurun_out/r02_prof_ncu4.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.txt 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_front_ap15.csv > gpurun_out/r02_launches_front_ap15_summary.txt 2>&1
echo done
