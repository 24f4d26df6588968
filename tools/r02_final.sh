#!/bin/bash
# round-2 final check on one B200: the whole GPU suite, the default bench line, the ncu launch list of a chained front
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_final_tests.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/r02_final_tests.log
timeout 330 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?" >> gpurun_out/r02_final_bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_final_launches.csv python tools/probe_synth.py ap:3:12 > gpurun_out/r02_final_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches_summary.txt 2>&1
