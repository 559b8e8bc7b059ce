#!/bin/bash
# 1-GPU emulation of the 8-GPU strong-scaling shard (1/8 of cfg2): synchronisation variants of the in-kernel tail
cd "$(dirname "$0")/.."
for tune in 0 1 2 3 4 7; do
  echo "tune=$tune"
  BB_PERSIST=256 BB_STEPK_TUNE=$tune QSCALE=0.125 QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
done
for gs in 16 24; do
  echo "gsize=$gs tune=3"
  BB_PERSIST=256 BB_STEPK_TUNE=3 BB_STEPK_GSIZE=$gs QSCALE=0.125 QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
done
echo "full tune=0 / tune=7"
BB_PERSIST=256 BB_STEPK_TUNE=0 QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
BB_PERSIST=256 BB_STEPK_TUNE=7 QN=512 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
