#!/bin/bash
# A/B builds of the library: scripts/build_variant.sh NAME [-DSRBD_K3_...=..]  ->  build/libsrbd_NAME.so  (SRBD_LIB=... selects it)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  --expt-relaxed-constexpr "$@" srbd-nmpc-solver_b200/csrc/capi.cu -o build/libsrbd_$name.so
echo built build/libsrbd_$name.so
