#!/bin/bash
# Round 2, GPU call X (2 GPUs): the N = 2 bench line with the final bench flow (tree context released before the range-proof leg), as the driver launches it.
mkdir -p gpurun_out
P=gpurun_out/r02x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 3 --warmup 3 --rp-singles 4096 --rp-aggregates 256 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; tail -4 ${P}_bench_n2.err
python - <<PY
import json
d = json.loads([l for l in open("${P}_bench_n2.json") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["phase_ms"].items()}, round(d["e2e"]["value"]/1e6,2), d["n_gpus"], d["config"]["comb_window"], d["config"]["parallelism"][:60])
rp = d["range_proofs"]
for k in ("n64_m1", "n64_m32"):
    print(k, round(rp[k]["prove_per_s"]), round(rp[k]["verify_per_s"]), rp[k]["all_verified"])
PY
