#!/bin/bash
# Kernel-variant experiments: tools/build_variant.sh NAME [-DDAPOL_PAD_MINB=5 ...] -> dapol_b200/lib/var_NAME.so
# (dapol_lib.cu recompiled with the extra flags, linked with the current dapol_rp / dapol_proof objects).
# Select at run time with DAPOL_B200_LIB=dapol_b200/lib/var_NAME.so.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v"
$NVCC $FLAGS "$@" -c -o build/var_$name.o dapol_b200/csrc/dapol_lib.cu > build/var_$name.log 2>&1
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -o dapol_b200/lib/var_$name.so build/var_$name.o build/dapol_rp.o build/dapol_proof.o
echo "built dapol_b200/lib/var_$name.so"
