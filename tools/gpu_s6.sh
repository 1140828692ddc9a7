#!/bin/bash
# Session D, call 2: shared-memory staging of the table look-ups: variants x comb windows, range-proof probe, parity.
mkdir -p gpurun_out
for V in default stage inl; do
  for W in 15 24; do
    lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
    DAPOL_B200_LIB=$lib timeout 300 python bench.py --steps 4 --warmup 2 --comb-window $W --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 \
      > gpurun_out/s6_bench_${V}_w$W.json 2> gpurun_out/s6_bench_${V}_w$W.err
    python - <<P
import json
try:
    d = json.load(open("gpurun_out/s6_bench_${V}_w$W.json")); print("$V W=$W", round(d["value"]/1e6,2), {k: round(v,2) for k,v in d["phase_ms"].items()}, d["root"])
except Exception as e: print("$V W=$W failed", e)
P
  done
done
for V in default rpstage; do
  lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
  echo "rp $V"
  DAPOL_B200_LIB=$lib RP_WINDOWS=12,16 timeout 600 python tools/rp_probe.py 64x1x16384 64x32x512 2> gpurun_out/s6_rp_$V.err | tee gpurun_out/s6_rp_$V.txt | cut -c1-330
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/s6_pytest_gpu.txt
