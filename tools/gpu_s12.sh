#!/bin/bash
# two-lane prover: parity + throughput against the single-lane build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
for V in default; do
  lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
  echo "rp $V"
  DAPOL_B200_LIB=$lib RP_WINDOWS=0 timeout 600 python tools/rp_probe.py 64x1x16384 64x1x16384 64x1x131072 64x1x131072 64x32x2048 64x32x2048 64x32x8192 64x32x8192 2> gpurun_out/s12_rp_$V.err | tee gpurun_out/s12_rp_$V.txt | cut -c1-300
done
