#!/bin/bash
# usage: tools/build_variant.sh <suffix> [-DFLAG=...]   -> vfmseg_b200/lib/libvfm_<suffix>.so (experiment builds of the same sources)
set -e
cd "$(dirname "$0")/.."
sfx=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o vfmseg_b200/lib/libvfm_${sfx}.so vfmseg_b200/csrc/api.cu
