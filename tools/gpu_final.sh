#!/bin/bash
# end-of-session check: smoke, the full GPU parity suite, the bench line with the CPU arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/final_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -2 gpurun_out/final_bench_n1.err
python - <<P
import json
d = json.load(open("gpurun_out/final_bench_n1.json")); print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]); print(d.get("range_proofs")); print(d.get("cpu_baseline"))
P
