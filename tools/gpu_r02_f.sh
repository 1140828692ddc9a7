#!/bin/bash
# Round 2, GPU call F (one GPU): Karatsuba field product -- parity (whole GPU suite incl. the new batch-proof and id tests),
# bench N=1, microbench of the field operations.
mkdir -p gpurun_out
P=gpurun_out/r02f
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
timeout 1200 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
python - <<E
import json
d = json.loads([l for l in open("${P}_bench_n1.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["roofline"]["fe_mul_Gop_s"], d["roofline"]["fe_sq_Gop_s"], d["e2e"]["value"], d["gpu_launches"])
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, {x: round(rp[k][x], 1) for x in ("prove_per_s", "verify_per_s", "e2e_prove_plus_verify_per_s")}, rp[k]["roofline"]["frac"], rp[k]["roofline"]["whole_prover_frac"])
print(rp.get("cpu_baseline")); print(d.get("cpu_baseline")); print(d.get("c1"))
E
