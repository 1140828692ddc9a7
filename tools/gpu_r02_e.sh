#!/bin/bash
# Round 2, GPU call E (8 GPUs): N=8 bench line (one-call sharded build vs the staged round-1 path), BASELINE config 3
# (all 2^20 inclusion proofs, m = 32) and the north-star target for real: 2^24 users / height 40, every inclusion proof
# (m = 64, 3,655 B) generated, written out and verified.
mkdir -p gpurun_out
P=gpurun_out/r02e
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
export NORTHSTAR_OUT=/dev/shm
nvidia-smi --query-gpu=index,name,memory.total --format=csv > ${P}_gpus.txt; nproc >> ${P}_gpus.txt; df -h /dev/shm /tmp >> ${P}_gpus.txt
timeout 600 $TR --master-port 29621 bench.py --gpus 8 --steps 6 --warmup 3 --rp-singles 0 --rp-aggregates 0 > ${P}_bench_n8.json 2> ${P}_bench_n8.err; tail -2 ${P}_bench_n8.err
timeout 600 $TR --master-port 29622 bench.py --gpus 8 --steps 6 --warmup 3 --rp-singles 0 --rp-aggregates 0 --staged-sharding > ${P}_bench_n8_staged.json 2> ${P}_bench_n8_staged.err
timeout 600 $TR --master-port 29623 tools/northstar.py 20 32 0 8192 0 8 > ${P}_c3_all_users.json 2> ${P}_c3_all_users.err; tail -2 ${P}_c3_all_users.err
timeout 1500 $TR --master-port 29624 tools/northstar.py 24 40 0 4096 0 8 > ${P}_northstar.json 2> ${P}_northstar.err; tail -3 ${P}_northstar.err
python - <<E
import json
def last(f):
    try:
        return json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
for s in ("", "_staged"):
    d = last("${P}_bench_n8%s.json" % s)
    if "value" in d:
        print(s or "one-call", round(d["value"] / 1e6, 2), "M leaves/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"] / 1e6, 2), d["root"])
    else:
        print(s, d)
for f in ("c3_all_users", "northstar"):
    d = last("${P}_%s.json" % f)
    print(f, json.dumps({k: d.get(k) for k in ("proofs", "all_verified", "tampered_rejected", "bytes_written", "sink", "tree_build_s_e2e", "tree_phase_ms_rank0", "shard0_root_equals_oracle_golden",
                                               "oracle_verified_sample", "prove_per_s", "verify_per_s", "proofs_per_s_wall", "prove_write_verify_wall_s", "total_wall_s_build_plus_proofs", "rangeproof_window", "error")}))
E
