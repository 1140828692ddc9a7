#!/bin/bash
# Session D, call 1: parity of the split merge + wide comb windows, window sweep, full bench, range-proof probe, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/s5_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s5_pytest_gpu.txt
for W in 15 20 22 24; do
  timeout 300 python bench.py --steps 5 --warmup 3 --comb-window $W --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 \
    > gpurun_out/s5_bench_w$W.json 2> gpurun_out/s5_bench_w$W.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/s5_bench_w$W.json")); print("W=$W", d["value"], d["phase_ms"], d["e2e"]["value"])
except Exception as e: print("W=$W failed", e)
P
done
timeout 600 python bench.py > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err; tail -3 gpurun_out/s5_bench_n1.err
RP_WINDOWS=12,16 timeout 600 python tools/rp_probe.py 64x1x16384 64x32x512 > gpurun_out/s5_rp_probe.txt 2> gpurun_out/s5_rp_probe.err
cat gpurun_out/s5_rp_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/s5_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/s5_ncu_bench.log 2>&1
for k in k_pad k_compress_internal k_merge_sum; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/s5_$k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/s5_ncu_$k.log 2>&1
  ncu -i gpurun_out/s5_$k.ncu-rep --page raw --csv > gpurun_out/s5_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/s5_$k.ncu-rep --page source --csv > gpurun_out/s5_${k}_source.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -30
