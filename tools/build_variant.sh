#!/bin/bash
# Kernel-variant experiments: tools/build_variant.sh NAME TU [-DDAPOL_PAD_MINB=4 ...] -> dapol_b200/lib/var_NAME.so
# (TU = dapol_lib | dapol_merge | dapol_rp | dapol_proof recompiled with the extra flags, linked with the current objects of the others).
# Select at run time with DAPOL_B200_LIB=dapol_b200/lib/var_NAME.so.
set -e
cd "$(dirname "$0")/.."
name=$1; tu=$2; shift 2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# variants are throw-away measurement builds: -split-compile for speed (not reproducible, see build.sh)
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -split-compile 0 -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v"
$NVCC $FLAGS "$@" -c -o build/var_$name.o dapol_b200/csrc/$tu.cu > build/var_$name.log 2>&1
objs=""
for t in dapol_lib dapol_merge dapol_rp dapol_proof; do
  if [ $t = $tu ]; then objs="$objs build/var_$name.o"; else objs="$objs build/$t.o"; fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -o dapol_b200/lib/var_$name.so $objs
echo "built dapol_b200/lib/var_$name.so"
