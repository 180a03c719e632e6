#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG ..." : recompile score_pair.cu with extra flags, link with the stock objects
# into hgrnet_b200/lib/var_NAME.so (select it with HGR_LIB=...).  Kernel experiments only.
set -e
name=$1; flags=$2
L=hgrnet_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -I include -c hgrnet_b200/csrc/score_pair.cu -o $L/var_$name.o
objs=$(ls $L/*.o | grep -v "var_\|score_pair.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $L/var_$name.so $objs $L/var_$name.o
rm $L/var_$name.o
echo built $L/var_$name.so
