#!/bin/bash
# round-2 probe 6: fused round templated; 2-rank GPU tests; per-stage B&B statistics; the 24-worker knapsack hang
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py tests/test_multi_rank_gpu.py tests/test_synergistic.py -x -q -m gpu > gpurun_out/r02_p6_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p6_tests.log
L=gpurun_out/r02_p6.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 100 python tools/probe_synth.py $SPEC >> $L 2>&1; echo "rc=$?" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=24 PROBE_WORKERS=12 MOIP_PROFILE_ROUNDS=1
SPEC=kp:4:40
run PROBE_SPLIT=32 PROBE_WORKERS=16 MOIP_PROFILE_ROUNDS=1
run PROBE_SPLIT=48 PROBE_WORKERS=12
run PROBE_SPLIT=24 PROBE_WORKERS=24
run PROBE_SPLIT=40 PROBE_WORKERS=20
timeout 300 python bench.py --no-fronts --steps 3 --warmup 3 --cpu-sample 64 > gpurun_out/r02_p6_bench.json 2> gpurun_out/r02_p6_bench.err
echo done
