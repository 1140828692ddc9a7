#!/bin/bash
# Session D, call 6: parity after the Straus verifier + launch-bound changes, bench, range-proof probe, C3 tool smoke on one GPU.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/s9_pytest_gpu.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s9_bench_n1.json 2> gpurun_out/s9_bench_n1.err; tail -3 gpurun_out/s9_bench_n1.err
python - <<P
import json
d = json.load(open("gpurun_out/s9_bench_n1.json")); print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"]); print(d.get("range_proofs"))
P
for V in default rpinl4; do
  lib=dapol_b200/lib/var_$V.so; [ $V = default ] && lib=dapol_b200/lib/libdapol_b200.so
  echo "rp $V"
  DAPOL_B200_LIB=$lib RP_WINDOWS=0 timeout 600 python tools/rp_probe.py 64x1x16384 64x1x131072 64x32x2048 2> gpurun_out/s9_rp_$V.err | tee gpurun_out/s9_rp_$V.txt | cut -c1-330
done
timeout 900 python tools/c3_all.py 20 32 0 4096 8192 > gpurun_out/s9_c3_1gpu.json 2> gpurun_out/s9_c3_1gpu.err; cat gpurun_out/s9_c3_1gpu.json; tail -3 gpurun_out/s9_c3_1gpu.err
