#!/bin/bash
# Round 2, GPU call A (one GPU): smoke, full GPU parity suite (incl. the full-size goldens), bench N=1 with the new legs
# (range-proof roofline + CPU baselines, C1 latency), the CPU arm, ncu launch list + full captures of the final kernels.
mkdir -p gpurun_out
P=gpurun_out/r02a
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > ${P}_gpu.txt; nproc >> ${P}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
timeout 1200 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file ${P}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > ${P}_ncu_bench.log 2>&1
for k in k_pad k_compress_internal k_merge_sum k_leaf; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o ${P}_$k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > ${P}_ncu_$k.log 2>&1
  ncu -i ${P}_$k.ncu-rep --page raw --csv > ${P}_${k}_raw.csv 2>/dev/null
done
# range-proof kernels: the m = 32 shape of config 3 (k_rp_p10 with inlined products, Straus verifier k_rp_v1) and m = 1
for shape in 64x32x512 64x1x8192; do
  for k in k_rp_p10 k_rp_v1; do
    skip=0; [ $k = k_rp_p10 ] && skip=7   # the probe's 2-proof warm-up batch launches k_rp_p10 six times: capture round 2 of the real batch
    RP_WINDOWS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o ${P}_${k}_${shape} -f \
      python tools/rp_probe.py $shape > ${P}_ncu_${k}_${shape}.log 2>&1
    ncu -i ${P}_${k}_${shape}.ncu-rep --page raw --csv > ${P}_${k}_${shape}_raw.csv 2>/dev/null
  done
done
rm -f gpurun_out/*.ncu-rep
python - <<P
import json
d = json.load(open("${P}_bench_n1.json"))
print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print(json.dumps(d.get("range_proofs"))[:3000]); print(d.get("cpu_baseline")); print(d.get("c1"))
P
