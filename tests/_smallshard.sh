#!/bin/bash
# 1-GPU emulation of the strong-scaling shards: persistent kernel on 1/2, 1/8 of cfg2, by resident CTAs per SM
cd "$(dirname "$0")/.."
for scale in 1.0 0.5 0.125; do
  for ctas in 3 2 1; do
    BB_PERSIST=256 BB_STEPK_CTAS=$ctas QSCALE=$scale QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
  done
done
BB_PERSIST=0 QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
