#!/bin/bash
# development helper (GPU box): time tuning builds of the library (BB_LIB_PATH) on cfg2, K = 8
# usage: _variants.sh lib.so[:ENV=VAL,ENV=VAL] ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  lib=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  echo "== $lib $envs"
  env BB_STEPK_MIN_OCC=2 BB_LIB_PATH=$PWD/barbay.jl_b200/$lib $envs QK=${QK:-8} timeout 300 python tests/_quickbench.py 2>&1 | tail -1
done
