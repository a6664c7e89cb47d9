#!/bin/bash
# Builds an A/B variant of the library: tools/build_variant.sh <name> [-DMACRO ...]  ->  gpurun_out/../variants/lib_<name>.so
# Select it at run time with POASTA_B200_LIB=variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
src=poasta_b200/csrc
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-pthread "$@" \
  -shared -o variants/lib_$name.so $src/align.cu $src/capi_graph.cpp
echo variants/lib_$name.so
