#!/bin/bash
# Session D, call 4: full bench line, range-proof probe, launch list + ncu captures of the new k_pad / p10.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s7_bench_n1.json 2> gpurun_out/s7_bench_n1.err; tail -3 gpurun_out/s7_bench_n1.err
python - <<P
import json
d = json.load(open("gpurun_out/s7_bench_n1.json")); print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d.get("range_proofs"), d.get("cpu_baseline"))
P
RP_WINDOWS=12,16,0 timeout 600 python tools/rp_probe.py 64x1x16384 64x32x512 64x32x2048 2> gpurun_out/s7_rp.err | tee gpurun_out/s7_rp.txt | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/s7_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/s7_ncu_bench.log 2>&1
for k in k_pad k_compress_internal; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/s7_$k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/s7_ncu_$k.log 2>&1
  ncu -i gpurun_out/s7_$k.ncu-rep --page raw --csv > gpurun_out/s7_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/s7_$k.ncu-rep --page source --csv > gpurun_out/s7_${k}_source.csv 2>/dev/null
done
for spec in "m32 3" "m1 13"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rp_p10 -s $2 -c 1 -o gpurun_out/s7_p10_$1 -f \
    python tools/rp_min.py > gpurun_out/s7_ncu_p10_$1.log 2>&1
  ncu -i gpurun_out/s7_p10_$1.ncu-rep --page raw --csv > gpurun_out/s7_p10_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/s7_p10_$1.ncu-rep --page source --csv > gpurun_out/s7_p10_$1_source.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
