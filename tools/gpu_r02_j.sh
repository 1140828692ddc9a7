#!/bin/bash
# Round 2, GPU call J (one GPU): parity of the new rows (no -x: every failure is listed), the verifier head-to-head after the
# bucket-balance fixes, generator-table window 12 (17.7 GB) vs automatic (110 GB) at m = 32, single-proof latency split (K = 1).
mkdir -p gpurun_out
P=gpurun_out/r02j
timeout 1500 python -m pytest tests/test_gpu_blake2b.py tests/test_gpu_update.py tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py \
  tests/test_gpu_batch_proof.py tests/test_gpu_persist.py -m gpu -q 2>&1 | tail -30 | tee ${P}_pytest_gpu.txt
timeout 900 python bench.py --no-cpu-baseline --no-c1 --steps 3 > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
timeout 900 python tools/c5_verify_sweep.py 16 0,16,256,4096 > ${P}_c5_groups.jsonl 2> ${P}_c5_groups.err; tail -3 ${P}_c5_groups.err
RP_WINDOWS=12,0 COMB_WINDOW=15 timeout 900 python tools/rp_probe.py 64x32x2048 64x16x1 64x1x1 > ${P}_rp_probe.txt 2> ${P}_rp_probe.err; tail -3 ${P}_rp_probe.err
cat ${P}_rp_probe.txt
python - <<PY
import json
d = json.loads([l for l in open("${P}_bench_n1.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), d["phase_ms"])
rp = d.get("range_proofs") or {}
for k in ("n64_m1", "n64_m32"):
    if k in rp:
        print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]),
              [(b["group"], round(b["verify_per_s"]), b["fallbacks"], b["all_verified"], round(b["kernel_class_ms"]["verifier"], 2)) for b in rp[k]["verify_batched_bucket_method"]])
for l in open("${P}_c5_groups.jsonl"):
    d = json.loads(l); print(d["nbits"], d["proofs_per_gpu"], d["verify_group"], round(d["verifies_per_s"]), d["verdicts_exact"], d["reverified_per_call"])
PY
