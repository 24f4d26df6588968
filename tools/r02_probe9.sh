#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu > gpurun_out/r02_p9_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p9_tests.log
L=gpurun_out/r02_p9.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 100 python tools/probe_synth.py $SPEC 2>&1 | grep -v watchdog >> $L; echo "rc=$?" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=16 PROBE_WORKERS=16
run PROBE_SPLIT=16 PROBE_WORKERS=16 MOIP_KERNEL_TIMING=1
run PROBE_SPLIT=12 PROBE_WORKERS=12
SPEC=kp:4:40
run PROBE_SPLIT=16 PROBE_WORKERS=16
run PROBE_SPLIT=4 PROBE_WORKERS=16
run PROBE_SPLIT=48 PROBE_WORKERS=12 MOIP_NO_STEAL=1 MOIP_WATCHDOG=10
timeout 300 python bench.py --no-fronts --steps 3 --warmup 3 --cpu-sample 64 --workload kp40 > gpurun_out/r02_p9_bench_kp40.json 2> gpurun_out/r02_p9_bench_kp40.err
timeout 300 python bench.py --no-fronts --steps 3 --warmup 3 --cpu-sample 64 > gpurun_out/r02_p9_bench_ap30.json 2> gpurun_out/r02_p9_bench_ap30.err
