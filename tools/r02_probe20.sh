#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu.py -x -q -m gpu -k "k1_" 2>&1 | tail -3 > gpurun_out/r02_p20_tests.log
MOIP_K1_BATCH_FARKAS=0 timeout 150 python bench.py --no-fronts --cpu-sample 16 --steps 3 --warmup 3 > gpurun_out/r02_p20_farkas0.json 2> gpurun_out/r02_p20.err
MOIP_K1_BATCH_FARKAS=1 timeout 150 python bench.py --no-fronts --cpu-sample 16 --steps 3 --warmup 3 > gpurun_out/r02_p20_farkas1.json 2>> gpurun_out/r02_p20.err
