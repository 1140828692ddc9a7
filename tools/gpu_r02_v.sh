#!/bin/bash
# Round 2, GPU call V (one GPU): the REPRODUCIBLE build of the final sources (no -split-compile) -- smoke, the whole GPU parity suite, bench N=1
# (all legs), CPU arm; then the 26-bit comb window experiment (variant library) against 24 bits.
mkdir -p gpurun_out
P=gpurun_out/r02v
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
timeout 1200 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err
for w in 24 26; do
  DAPOL_B200_LIB=dapol_b200/lib/var_w26.so timeout 600 python bench.py --comb-window $w --no-cpu-baseline --no-c1 --rp-singles 0 --rp-aggregates 0 --steps 6 > ${P}_bench_w$w.json 2> ${P}_bench_w$w.err; tail -2 ${P}_bench_w$w.err
done
python - <<PY
import json
def last(f): return json.loads([l for l in open(f) if l.startswith("{")][-1])
d = last("${P}_bench_n1.json")
print(round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print(json.dumps(d.get("c1")))
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), rp[k]["roofline"]["frac"], [(b["group"], round(b["verify_per_s"])) for b in rp[k]["verify_batched_bucket_method"]])
print(d.get("cpu_baseline"))
for w in (24, 26):
    try:
        x = last("${P}_bench_w%d.json" % w); print("comb window", w, round(x["value"]/1e6,2), {k: round(v,2) for k,v in x["phase_ms"].items()}, x["roofline"]["frac"], x["root"])
    except Exception as e:
        print("comb window", w, "failed", e)
PY
