#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=..]...   -> dnet_b200/lib/ab/libdnet_b200_NAME.so (A/B timing experiments)
set -e
name=$1; shift
mkdir -p dnet_b200/lib/ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -cudart static "$@" \
  -o dnet_b200/lib/ab/libdnet_b200_$name.so dnet_b200/csrc/dn_api.cu
echo dnet_b200/lib/ab/libdnet_b200_$name.so
