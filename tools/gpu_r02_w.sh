#!/bin/bash
# Round 2, GPU call W (one GPU): final library with the 26-bit comb window instantiated; bench with its new defaults (tree leg on the 26-bit
# tables, range-proof leg on its own context) -- smoke, the whole GPU parity suite, bench N=1.
mkdir -p gpurun_out
P=gpurun_out/r02w
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 900 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -5 ${P}_bench_n1.err
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
python - <<PY
import json
d = json.loads([l for l in open("${P}_bench_n1.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["config"]["comb_window"])
print(json.dumps(d.get("c1")))
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), rp[k]["roofline"]["frac"], rp[k]["roofline"]["rangeproof_window"], [(b["group"], round(b["verify_per_s"])) for b in rp[k]["verify_batched_bucket_method"]])
print(d.get("cpu_baseline"))
PY
