#!/bin/bash
# Round 2, GPU call C (2 GPUs): the one-call sharded build over the library's own NCCL communicator -- root equality with the
# one-GPU build, proofs verified (GPU + oracle sample), and the N=2 bench line (one-call vs the staged round-1 path).
mkdir -p gpurun_out
P=gpurun_out/r02c
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
df -h /tmp /dev/shm > ${P}_disk.txt 2>&1; free -g >> ${P}_disk.txt; nproc >> ${P}_disk.txt
python tools/northstar.py 16 24 0 1024 2048 4 > ${P}_ns_1gpu.json 2> ${P}_ns_1gpu.err; tail -2 ${P}_ns_1gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29611 tools/northstar.py 16 24 0 1024 2048 4 > ${P}_ns_2gpu.json 2> ${P}_ns_2gpu.err; tail -3 ${P}_ns_2gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29612 tools/northstar.py 18 30 1 1024 1024 2 > ${P}_ns_2gpu_splitting.json 2> ${P}_ns_2gpu_splitting.err; tail -3 ${P}_ns_2gpu_splitting.err
python - <<E
import json
a, b = (json.load(open("${P}_ns_%s.json" % s)) for s in ("1gpu", "2gpu"))
print("root 1gpu == root 2gpu:", a["root"] == b["root"], a["root"][:16], "| all verified:", a["all_verified"], b["all_verified"], "| oracle:", a["oracle_verified_sample"], b["oracle_verified_sample"])
print(json.dumps(b)[:1500])
E
timeout 600 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 6 --warmup 3 --rp-singles 0 --rp-aggregates 0 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; tail -2 ${P}_bench_n2.err
timeout 600 $TR --nproc-per-node 2 --master-port 29614 bench.py --gpus 2 --steps 6 --warmup 3 --rp-singles 0 --rp-aggregates 0 --staged-sharding > ${P}_bench_n2_staged.json 2> ${P}_bench_n2_staged.err
python - <<E
import json
for s in ("", "_staged"):
    d = json.load(open("${P}_bench_n2%s.json" % s))
    print(s or "one-call", round(d["value"] / 1e6, 2), "M leaves/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"] / 1e6, 2), d["root"])
E
