#!/bin/bash
# experiment sweep over launch configurations of the fast K1 kernel (needs a MOIP_K1_EXPERIMENT build)
for cfg in 0 1 2 3 4; do
  for occ in 0; do
    echo "cfg=$cfg (0:128x4 1:128x5 2:128x6 3:256x2 4:256x3)"
    MOIP_K1_CFG=$cfg python tools/run_k1_once.py ap30 8192 1000
  done
done
echo "cfg=0 carveout 100"; MOIP_K1_CFG=0 MOIP_K1_CARVEOUT_PCT=100 python tools/run_k1_once.py ap30 8192 1000
echo "cfg=1 occ4";  MOIP_K1_CFG=1 MOIP_K1_OCC=4 python tools/run_k1_once.py ap30 8192 1000
echo "cfg=0 occ3";  MOIP_K1_CFG=0 MOIP_K1_OCC=3 python tools/run_k1_once.py ap30 8192 1000
echo "cfg=0 norm1"; MOIP_K1_CFG=0 MOIP_NORM_EVERY=1 python tools/run_k1_once.py ap30 8192 1000
echo "cfg=0 norm8"; MOIP_K1_CFG=0 MOIP_NORM_EVERY=8 python tools/run_k1_once.py ap30 8192 1000
