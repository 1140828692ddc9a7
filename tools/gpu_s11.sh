#!/bin/bash
# Session D, call 9: full parity run, the bench line with the CPU arm, launch list, C5 sweep, C3 per-chunk timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s11_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/s11_bench_n1.json 2> gpurun_out/s11_bench_n1.err; tail -3 gpurun_out/s11_bench_n1.err
python - <<P
import json
d = json.load(open("gpurun_out/s11_bench_n1.json")); print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]); print(d.get("range_proofs")); print(d.get("cpu_baseline"))
P
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s11_bench_ref.json 2> gpurun_out/s11_bench_ref.err; cat gpurun_out/s11_bench_ref.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s11_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/s11_ncu_bench.log 2>&1
timeout 900 python tools/c5_verify_sweep.py 20 > gpurun_out/s11_c5.jsonl 2> gpurun_out/s11_c5.err; cat gpurun_out/s11_c5.jsonl | cut -c1-400; tail -2 gpurun_out/s11_c5.err
timeout 900 python tools/c3_all.py 20 32 0 8192 32768 > gpurun_out/s11_c3_1gpu.json 2> gpurun_out/s11_c3_1gpu.err; cat gpurun_out/s11_c3_1gpu.json | cut -c1-900; tail -2 gpurun_out/s11_c3_1gpu.err
