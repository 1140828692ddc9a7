#!/bin/bash
# Session D, call 5: occupancy variants (launch bounds, batch size) for the node kernels and the range-proof MSM kernels.
mkdir -p gpurun_out
for V in default minb4 minb45 nb32 minb5; do
  lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
  DAPOL_B200_LIB=$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 \
    > gpurun_out/s8_bench_${V}.json 2> gpurun_out/s8_bench_${V}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/s8_bench_${V}.json")); print("$V", round(d["value"]/1e6,2), {k: round(v,2) for k,v in d["phase_ms"].items()}, d["root"])
except Exception as e: print("$V failed", e)
P
done
for V in default rpminb4; do
  lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
  echo "rp $V"
  DAPOL_B200_LIB=$lib RP_WINDOWS=16,0 timeout 600 python tools/rp_probe.py 64x1x16384 64x32x2048 2> gpurun_out/s8_rp_$V.err | tee gpurun_out/s8_rp_$V.txt | cut -c1-330
done
