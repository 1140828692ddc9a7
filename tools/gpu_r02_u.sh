#!/bin/bash
# Round 2, GPU call U (2 GPUs): the N = 2 bench line of the final library, launched as the driver launches it.
mkdir -p gpurun_out
P=gpurun_out/r02u
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 4 --warmup 3 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; tail -3 ${P}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > ${P}_bench_n2_ref.json 2> ${P}_bench_n2_ref.err; tail -2 ${P}_bench_n2_ref.err
python - <<PY
import json
d = json.loads([l for l in open("${P}_bench_n2.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phase_ms"].items()}, d["e2e"]["value"], d["gpu_launches"], d.get("scaling"), d["n_gpus"])
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), [(b["group"], round(b["verify_per_s"])) for b in rp[k]["verify_batched_bucket_method"]])
r = json.loads([l for l in open("${P}_bench_n2_ref.json") if l.startswith("{")][-1]); print(r["impl"], r["value"], r["n_gpus"])
PY
