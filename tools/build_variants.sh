#!/bin/bash
# Comparison builds of the Q = 2 tile kernels: tools/build_variants.sh "<name>=<nvcc flags> ..."  ->  build_variants/libsbmbp_<name>.so
# (run after `make`; select one with SBMBP_LIB=build_variants/libsbmbp_<name>.so; tools/variant_ab.sh times them).  engine.cu and the
# Q = 2 instantiations are rebuilt with the flags (tile geometry lives in both), the rest is linked from the regular build.
# The directory is git-ignored but travels to the GPU box.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/sbm-bp_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin g++"
mkdir -p $ROOT/build_variants
for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    $NV $flags -DINST_QT=2 -Xptxas -v -c $C/inst.cu -o $C/build/var_${name}_inst_2.o 2> $C/build/ptxas_var_${name}.log &
    $NV $flags -c $C/engine.cu -o $C/build/var_${name}_engine.o &
    wait
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build_variants/libsbmbp_${name}.so $C/build/var_${name}_engine.o $C/build/graph.o \
        $C/build/var_${name}_inst_2.o $C/build/inst_4.o $C/build/inst_8.o $C/build/inst_16.o $C/build/inst_32.o -lpthread
    grep -A2 "bp_sweep_pipe_kernelIdLi2ELb0" $C/build/ptxas_var_${name}.log | grep "Used\|spill" | sed "s/^/$name: /"
done
