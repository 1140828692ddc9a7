#!/bin/bash
# Round 2, GPU call T (one GPU): FINAL library -- smoke, the whole GPU parity suite, bench N=1 (all legs), the CPU arm, ncu launch list of the bench.
mkdir -p gpurun_out
P=gpurun_out/r02t
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > ${P}_gpu.txt; nproc >> ${P}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
timeout 1200 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${P}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > ${P}_ncu_bench.log 2>&1
python - <<PY
import json
d = json.loads([l for l in open("${P}_bench_n1.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print(json.dumps(d.get("c1")))
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), rp[k]["roofline"]["frac"], [(b["group"], round(b["verify_per_s"])) for b in rp[k]["verify_batched_bucket_method"]])
print(d.get("cpu_baseline")); print(rp.get("cpu_baseline"))
PY
