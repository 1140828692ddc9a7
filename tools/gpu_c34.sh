#!/bin/bash
# 8-GPU run of BASELINE configs[3] (2^24 users, height 40, leaf ranges sharded, root gather) and configs[2] (inclusion proofs
# for all 2^20 users of a height-32 tree, proved and verified, sharded).   bash tools/gpu_c34.sh [N=8]
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/c34_gpus.txt
k=0; m=$N; while [ $m -gt 1 ]; do m=$((m / 2)); k=$((k + 1)); done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# C4: one tree of 2^24 users, height 40, over N GPUs (2^24 / N users per GPU)
timeout 900 $TR --master-port 29521 bench.py --gpus $N --users-log2 $((24 - k)) --height $((40 - k)) --steps 4 --warmup 3 --rp-singles 0 --rp-aggregates 0 \
  > gpurun_out/c4_n$N.json 2> gpurun_out/c4_n$N.err
tail -c 400 gpurun_out/c4_n$N.err
# the same tree on ONE GPU (130 GB of node store with the build-time half points): the root must be identical
timeout 900 python bench.py --users-log2 24 --height 40 --comb-window 15 --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 \
  > gpurun_out/c4_single.json 2> gpurun_out/c4_single.err
tail -c 400 gpurun_out/c4_single.err
python - <<PY
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"error": str(e)}
a, b = last("gpurun_out/c4_n$N.json"), last("gpurun_out/c4_single.json")
print("C4 $N GPUs:", a.get("value"), a.get("ms_per_step"), a.get("phase_ms"), "e2e", (a.get("e2e") or {}).get("value"), "root", a.get("root"))
print("C4 1 GPU  :", b.get("value"), b.get("ms_per_step"), b.get("phase_ms"), "root", b.get("root"), "EQUAL" if a.get("root") and a.get("root") == b.get("root") else "DIFFERENT/UNKNOWN")
PY
# C3: all 2^20 inclusion proofs, both policies
for pol in 0 1; do
  timeout 900 $TR --master-port 2952$((2 + pol)) tools/c3_all.py 20 32 $pol 8192 > gpurun_out/c3_all_p$pol.json 2> gpurun_out/c3_all_p$pol.err
  tail -n 1 gpurun_out/c3_all_p$pol.json; tail -c 300 gpurun_out/c3_all_p$pol.err
done
