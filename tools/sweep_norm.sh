#!/bin/bash
# restart-test cadence / check cadence sweep for K1 (converged runs, 3AP30, B=16384)
for ne in 4 8 16; do for ce in 32 64; do
  echo -n "norm_every=$ne check_every=$ce: "; MOIP_NORM_EVERY=$ne MOIP_CHECK_EVERY=$ce python tools/run_k1_once.py ap30 16384 0
done; done
