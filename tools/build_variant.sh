#!/bin/bash
# Builds a variant of the product library with extra -D flags for pruned.cu (kernel experiments):
#   tools/build_variant.sh <name> [-DPR_WARPS=4 ...]  ->  autopas_b200/csrc/build/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../autopas_b200/csrc"
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr -DAPB_BUILD -I../../include -I. "$@" -c pruned.cu -o build/variants/pruned_$name.o
objs=$(ls build/*.o | grep -v pruned.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/pruned_$name.o -lcudart
echo built build/variants/lib_$name.so
