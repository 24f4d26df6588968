#!/bin/bash
# round-2 probe 5: fused round (propagate -> LP -> round/verify in one CTA) on the register-resident K1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu > gpurun_out/r02_p5_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_p5_tests.log
L=gpurun_out/r02_p5.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 150 python tools/probe_synth.py $SPEC 2>&1 | grep -v "B&B rounds" >> $L; }
SPEC=ap:3:30
run PROBE_SPLIT=24 PROBE_WORKERS=12 MOIP_KERNEL_TIMING=1
run PROBE_SPLIT=24 PROBE_WORKERS=12
run PROBE_SPLIT=32 PROBE_WORKERS=16
run PROBE_SPLIT=48 PROBE_WORKERS=24
run PROBE_SPLIT=24 PROBE_WORKERS=12 MOIP_FUSED_ROUND=0
SPEC=kp:4:40
run PROBE_SPLIT=48 PROBE_WORKERS=24
timeout 300 python bench.py --no-fronts --steps 3 --warmup 3 --cpu-sample 64 > gpurun_out/r02_p5_bench.json 2> gpurun_out/r02_p5_bench.err
echo done
