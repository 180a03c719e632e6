#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG ..." [SOURCE] : recompile SOURCE (default score_pair.cu) with extra flags, link with the stock objects
# into hgrnet_b200/lib/var_NAME.so (select it with HGR_LIB=...).  Kernel experiments only.
set -e
name=$1; flags=$2; src=${3:-score_pair.cu}; stem=${src%.cu}
L=hgrnet_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -I include -c hgrnet_b200/csrc/$src -o $L/var_$name.o
objs=$(ls $L/*.o | grep -v "var_\|$stem.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $L/var_$name.so $objs $L/var_$name.o
rm $L/var_$name.o
echo built $L/var_$name.so
