#!/bin/bash
# Builds the in-tree CUDA library for sm_100a (B200).  The .so is git-ignored but travels with gpurun.
set -e
cd "$(dirname "$0")"
mkdir -p dapol_b200/lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  ${DAPOL_PTXAS_V:+-Xptxas -v} --shared -o dapol_b200/lib/libdapol_b200.so dapol_b200/csrc/dapol_lib.cu "$@"
echo "built dapol_b200/lib/libdapol_b200.so"
