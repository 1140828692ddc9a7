#!/bin/bash
# Builds the in-tree CUDA library for sm_100a (B200).  The .so is git-ignored but travels with gpurun.
# A translation unit is recompiled only when it, a header, or this script is newer than its object (FORCE=1: all).
set -e
cd "$(dirname "$0")"
# DAPOL_VARIANT=name: a kernel-variant build (extra nvcc flags as arguments) into build/var_name/ and dapol_b200/lib/var_name.so,
# selected at run time with DAPOL_B200_LIB=dapol_b200/lib/var_name.so; the default library is untouched.
BUILD=build${DAPOL_VARIANT:+/var_$DAPOL_VARIANT}
OUT=dapol_b200/lib/${DAPOL_VARIANT:+var_$DAPOL_VARIANT.so}
OUT=${OUT%/}; [ -n "$DAPOL_VARIANT" ] || OUT=dapol_b200/lib/libdapol_b200.so
mkdir -p dapol_b200/lib $BUILD
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# No -split-compile by default: nvcc 12.9 gives one of two different SASS images of the same source from run to run with it (measured:
# profiles/r02_variants.txt 10), and one of them exposed a miscompile; without it the build is reproducible (5 min instead of 3).
# DAPOL_FAST_BUILD=1 turns it back on for development builds.
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo ${DAPOL_FAST_BUILD:+-split-compile 0} -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden ${DAPOL_PTXAS_V:+-Xptxas -v}"
newest_hdr=$(ls -t dapol_b200/csrc/*.cuh dapol_b200/csrc/*.h dapol_b200/csrc/*.inc include/*.h build.sh | head -1)
pids=(); tus=()
for tu in dapol_lib dapol_merge dapol_rp dapol_proof dapol_shard; do
  o=$BUILD/$tu.o
  if [ -n "$FORCE" ] || [ -n "$*" ] || [ ! -f $o ] || [ dapol_b200/csrc/$tu.cu -nt $o ] || [ "$newest_hdr" -nt $o ]; then
    $NVCC $FLAGS -c -o $o.tmp dapol_b200/csrc/$tu.cu "$@" > $BUILD/$tu.log 2>&1 && mv $o.tmp $o &
    pids+=($!); tus+=($tu)
  fi
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
for tu in "${tus[@]}"; do cat $BUILD/$tu.log; done
[ $rc -eq 0 ] || { echo "build failed"; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -o $OUT $BUILD/dapol_lib.o $BUILD/dapol_merge.o $BUILD/dapol_rp.o $BUILD/dapol_proof.o $BUILD/dapol_shard.o -ldl
echo "built $OUT (recompiled: ${tus[*]:-nothing})"
