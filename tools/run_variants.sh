#!/bin/bash
# Development helper (GPU box): time each tuning variant with tests/_quickbench.py
# usage: tools/run_variants.sh "K list" name1 name2 ...
ks=$1; shift
mkdir -p gpurun_out
for n in "$@"; do
  for k in $ks; do
    echo -n "$n K=$k " | tee -a gpurun_out/variants.log
    BB_LIB_PATH=$PWD/barbay.jl_b200/csrc/build/variants/$n.so QK=$k python tests/_quickbench.py 2>&1 | tail -1 | tee -a gpurun_out/variants.log
  done
done
