#!/bin/bash
# build_variant.sh <name> <extra nvcc defines...>  ->  gsrast_b200/variants/lib_<name>.so
set -e
name=$1; shift
cd "$(dirname "$0")/../gsrast_b200/csrc"
mkdir -p ../variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -shared -o ../variants/lib_$name.so preprocess.cu binning.cu radix_sort.cu bin_expand.cu blend.cu forward.cu views.cu ply.cu -lcudart
echo built $name
