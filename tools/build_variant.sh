#!/bin/bash
# Development helper: build a tuning variant of the library that carries only the fp32 (and optionally fp64)
# T=5 column kernels, compiled with extra -D flags:
#   tools/build_variant.sh NAME "-DBB_X=1 ..." [double]  ->  barbay.jl_b200/csrc/build/variants/NAME.so
# (selected at run time with BB_LIB_PATH; never used by tests or bench.py)
set -e
cd "$(dirname "$0")/../barbay.jl_b200/csrc"
name=$1; flags=$2; dbl=$3
d=build/variants/$name; mkdir -p $d
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
for u in bb_capi bb_engine_f32 bb_engine_f64 bb_registry; do $NV $flags -c $u.cu -o $d/$u.o 2> $d/$u.log & done
$NV -DBB_REAL=float -DBB_NT=5 $flags -c bb_inst.cu -o $d/inst_f5.o 2> $d/inst_f5.log &
objs="$d/bb_capi.o $d/bb_engine_f32.o $d/bb_engine_f64.o $d/bb_registry.o build/bb_layout.o $d/inst_f5.o"
if [ -n "$dbl" ]; then $NV -DBB_REAL=double -DBB_NT=5 $flags -c bb_inst.cu -o $d/inst_d5.o 2> $d/inst_d5.log & objs="$objs $d/inst_d5.o"; fi
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs -ldl
grep -A2 "pass2_kernelIfLi5ELi1ELb0ELb0ELb0ELb1" $d/inst_f5.log | grep -E "Used|spill" | head -3
