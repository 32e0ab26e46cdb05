#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh <name> [-DMACRO ...]  -> gpurun_out/variants/lib_<name>.so
set -e
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC --shared -cudart shared "$@" \
  -o variants/lib_${name}.so buffer_b200/csrc/mutual_nn.cu buffer_b200/csrc/mutual_nn_tc.cu buffer_b200/csrc/ransac.cu buffer_b200/csrc/refine.cu buffer_b200/csrc/extras.cu buffer_b200/csrc/api.cu
