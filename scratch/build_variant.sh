#!/bin/bash
# usage: scratch/build_variant.sh <name> <extra nvcc flags...>  -> scratch/exp/libgennet_<name>.so (synth.cu rebuilt with the flags)
set -e
NAME=$1; shift
cd /root/repo/gennet_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DGN_QUICK "$@" -Xptxas -v -c synth.cu -o ../../scratch/exp/synth_$NAME.o 2> ../../scratch/exp/synth_$NAME.ptxas.log
OBJS=$(ls *.o | grep -v '^synth.o$')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scratch/exp/libgennet_$NAME.so $OBJS ../../scratch/exp/synth_$NAME.o -lcudart -lcuda
grep -A1 "synth_kernelILi12ELi0ELi1" ../../scratch/exp/synth_$NAME.ptxas.log | grep -E "registers|spill" | head -4
