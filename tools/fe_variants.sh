#!/bin/bash
# Generates several mul/sqr carry-chain variants, compiles the microbenchmark for each and runs it (GPU box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/fev
run() { # name, env...
  name=$1; shift
  env "$@" DAPOL_FE_OUT=/tmp/fev/$name.inc python tools/gen_fe_mul.py 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDAPOL_FE_GEN_INC="\"/tmp/fev/$name.inc\"" -o /tmp/fev/$name tools/fe_microbench.cu 2>&1 | grep -i error
  /tmp/fev/$name "$name" | tee -a gpurun_out/fe_variants.txt
}
: > gpurun_out/fe_variants.txt
run base_chainfold DAPOL_FE_FOLD=chain
run plainfold DAPOL_FE_FOLD=plain
run plainfold_diag DAPOL_FE_FOLD=plain DAPOL_FE_SQR_DIAG=plain
run mul2 DAPOL_FE_PLAIN_MUL=2:0,5:1
run mul4 DAPOL_FE_PLAIN_MUL=1:0,3:1,5:0,7:1
run mul6 DAPOL_FE_PLAIN_MUL=1:0,2:1,3:0,4:1,5:0,6:1
run mul8 DAPOL_FE_PLAIN_MUL=1:0,1:1,3:0,3:1,5:0,5:1,7:0,7:1
run mul10 DAPOL_FE_PLAIN_MUL=1:0,1:1,2:0,3:0,3:1,4:1,5:0,5:1,7:0,7:1
run mul14 DAPOL_FE_PLAIN_MUL=1:0,1:1,2:0,2:1,3:0,3:1,4:0,4:1,5:0,5:1,6:0,6:1,7:0,7:1
run sq2 DAPOL_FE_PLAIN_SQR=1:1,3:1
run sq4 DAPOL_FE_PLAIN_SQR=0:0,1:1,2:0,3:1
run sq6 DAPOL_FE_PLAIN_SQR=0:0,0:1,1:1,2:0,3:1,4:0
run sq4_diag DAPOL_FE_PLAIN_SQR=0:0,1:1,2:0,3:1 DAPOL_FE_SQR_DIAG=plain
run sqall DAPOL_FE_PLAIN_SQR=0:0,0:1,1:0,1:1,2:0,2:1,3:0,3:1,4:0,4:1,5:0,5:1,6:0,6:1
