#!/bin/bash
# Round 2, GPU call I (one GPU): parity of the new rows -- Blake2b 64-byte digests, Dapol::update, batched verification by the
# bucket method -- plus the proof paths they touch; bench N=1 with the verifier head-to-head; the 128-byte table-slot variant of
# the tree kernels; C5 sweep with 1 % bad proofs per verifier mode.
mkdir -p gpurun_out
P=gpurun_out/r02i
timeout 1500 python -m pytest tests/test_gpu_blake2b.py tests/test_gpu_update.py tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py \
  tests/test_gpu_batch_proof.py tests/test_gpu_persist.py -m gpu -x -q 2>&1 | tail -12 | tee ${P}_pytest_gpu.txt
timeout 900 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; tail -3 ${P}_bench_n1.err
DAPOL_B200_LIB=dapol_b200/lib/var_pad128.so timeout 600 python bench.py --no-cpu-baseline --no-c1 --rp-singles 0 --rp-aggregates 0 \
  > ${P}_bench_n1_pad128.json 2> ${P}_bench_n1_pad128.err; tail -3 ${P}_bench_n1_pad128.err
timeout 900 python tools/c5_verify_sweep.py 16 0,16,64,1024 > ${P}_c5_groups.jsonl 2> ${P}_c5_groups.err; tail -3 ${P}_c5_groups.err
python - <<PY
import json
for f in ("${P}_bench_n1.json", "${P}_bench_n1_pad128.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, round(d["value"]/1e6,2), d["phase_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
    rp = d.get("range_proofs") or {}
    for k in ("n64_m1", "n64_m32"):
        if k in rp:
            print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), [(b["group"], round(b["verify_per_s"]), b["fallbacks"], b["all_verified"]) for b in rp[k]["verify_batched_bucket_method"]])
for l in open("${P}_c5_groups.jsonl"):
    d = json.loads(l); print(d["nbits"], d["proofs_per_gpu"], d["verify_group"], round(d["verifies_per_s"]), d["verdicts_exact"], d["reverified_per_call"])
PY
