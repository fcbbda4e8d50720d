#!/bin/bash
# Build the product library (CUDA kernels + C ABI) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
mkdir -p piranha_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -Iinclude -diag-suppress 186 ${PK_NVCC_EXTRA} -o piranha_b200/lib/libpiranha_b200.so piranha_b200/csrc/pk_abi.cu
