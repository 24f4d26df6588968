#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02_p19_tests.log
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -6 > $T; echo "pytest rc=$?" >> $T
bash tools/r02_probe17.sh "$@"
mv gpurun_out/r02_p17.log gpurun_out/r02_p19.log
