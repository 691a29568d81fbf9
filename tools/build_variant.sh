#!/bin/bash
# build a tuning variant of librt3.so with extra -D flags:  tools/build_variant.sh <name> [-DRT3_...=..]...
# result: rendertoy3c_b200/variants/librt3_<name>.so  (select at run time with RT3_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p rendertoy3c_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-strict-aliasing -shared -lcudart -ldl "$@" \
  -o rendertoy3c_b200/variants/librt3_$name.so rendertoy3c_b200/csrc/rt3_lib.cu
