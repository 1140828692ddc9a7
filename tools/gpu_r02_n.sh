#!/bin/bash
# Round 2, GPU call N (8 GPUs), final library: N=8 bench line (tree leg), and BASELINE config 3 -- all 2^20 inclusion proofs (m = 32) proved,
# written and verified -- with the per-proof verifier and with the group verifier (bucket method, G = 256).
mkdir -p gpurun_out
P=gpurun_out/r02n
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
export NORTHSTAR_OUT=/dev/shm
nvidia-smi --query-gpu=index,name,memory.total --format=csv > ${P}_gpus.txt; nproc >> ${P}_gpus.txt
VERIFY_GROUP=256 timeout 600 $TR --master-port 29632 tools/northstar.py 20 32 0 8192 0 8 > ${P}_c3_all_users_g256.json 2> ${P}_c3_all_users_g256.err; tail -2 ${P}_c3_all_users_g256.err
python - <<PY
import json
def last(f):
    try:
        return json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
d = {"skipped": "the N = 8 bench line is r02e (same tree code) and the driver's own scaling run"}
if "value" in d:
    print(round(d["value"] / 1e6, 2), "M leaves/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"] / 1e6, 2), d["root"])
else:
    print(d)
d = last("${P}_c3_all_users_g256.json")
print(json.dumps({k: d.get(k) for k in ("proofs", "all_verified", "tampered_rejected", "verify_group", "verify_fallbacks_rank0", "bytes_written", "tree_build_s_e2e", "oracle_verified_sample",
                                        "prove_per_s", "verify_per_s", "proofs_per_s_wall", "prove_write_verify_wall_s", "rank0_chunk_s_prove_write_verify", "error")}))
PY
