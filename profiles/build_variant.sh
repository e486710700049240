#!/bin/bash
# Build an experiment variant of the library next to the product one:
#   profiles/build_variant.sh <name> [extra nvcc flags...]   ->  leven_b200/lib/libleven_b200.<name>.so
# selected at run time with LVN_LIB_VARIANT=<name> (leven_b200/compute.py).  A/B runs inside one gpurun call.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
cd leven_b200/csrc
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o ../lib/libleven_b200.$name.so \
    api.cu kernels_chunk.cu kernels_csg.cu kernels_util.cu seam.cu simplify.cu clipmap_update.cu
