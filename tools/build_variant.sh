#!/bin/bash
# Debug: link a variant of the library with extra nvcc flags for gemm.cu only -> controllora_b200/libclb_<name>.so (select it with CLB_LIB=...)
#   tools/build_variant.sh NAME "-DFLAG1 -DFLAG2"
set -e
cd "$(dirname "$0")/../controllora_b200"
mkdir -p build_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -c csrc/gemm.cu -o build_var/gemm_$1.o
objs=$(ls build/*.o | grep -v gemm.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libclb_$1.so build_var/gemm_$1.o $objs -lcudart
echo libclb_$1.so
