#!/bin/bash
# development helper (GPU box): time tuning builds of the library (BB_LIB_PATH) on cfg2, K = 8
cd "$(dirname "$0")/.."
for lib in "$@"; do
  for mode in BB_PERSIST=0 BB_PERSIST=256; do
    echo "== $lib $mode"
    env BB_STEPK_MIN_OCC=2 BB_LIB_PATH=$PWD/barbay.jl_b200/$lib $mode QK=8 timeout 300 python tests/_quickbench.py 2>&1 | tail -1
  done
done
