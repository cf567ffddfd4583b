#!/bin/bash
# usage: tools/build_variant.sh <name> "<extra nvcc flags for gemm_tcgen05.cu>"   -> open_flamingo_b200/libofk_<name>.so
set -e
cd "$(dirname "$0")/../open_flamingo_b200"
name=$1; flags=$2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c csrc/gemm_tcgen05.cu -o csrc/build/gemm_tcgen05_$name.o
objs=$(ls csrc/build/*.o | grep -v "gemm_tcgen05")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libofk_$name.so $objs csrc/build/gemm_tcgen05_$name.o
rm -f csrc/build/gemm_tcgen05_$name.o
ls -la libofk_$name.so
