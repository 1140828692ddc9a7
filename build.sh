#!/bin/bash
# Builds the in-tree CUDA library for sm_100a (B200).  The .so is git-ignored but travels with gpurun.
set -e
cd "$(dirname "$0")"
mkdir -p dapol_b200/lib build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden ${DAPOL_PTXAS_V:+-Xptxas -v}"
pids=()
for tu in dapol_lib dapol_rp dapol_proof; do
  $NVCC $FLAGS -c -o build/$tu.o dapol_b200/csrc/$tu.cu "$@" > build/$tu.log 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
cat build/dapol_lib.log build/dapol_rp.log build/dapol_proof.log
[ $rc -eq 0 ] || { echo "build failed"; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -o dapol_b200/lib/libdapol_b200.so build/dapol_lib.o build/dapol_rp.o build/dapol_proof.o
echo "built dapol_b200/lib/libdapol_b200.so"
