#!/bin/bash
# Weak-scaling run on N GPUs of one box: bench.py at 1, 2, .. N GPUs back to back + root equality of the sharded tree with the
# single-GPU build of the same liabilities.   bash tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 8 --warmup 3 \
      > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -c 600 gpurun_out/scale_n$n.err
done
# the same tree as the N-GPU run, on one GPU: 2^20 * N users at height 32 + log2 N  (root must be identical)
k=0; m=$N; while [ $m -gt 1 ]; do m=$((m / 2)); k=$((k + 1)); done
timeout 600 python bench.py --users-log2 $((20 + k)) --height $((32 + k)) --steps 1 --warmup 1 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/scale_single_tree.json 2> gpurun_out/scale_single_tree.err
python - <<PY
import json
a = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/scale_single_tree.json").read().strip().splitlines()[-1])
print("sharded root", a["root"], "single-GPU root", b["root"], "EQUAL" if a["root"] == b["root"] else "DIFFERENT")
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, "GPUs:", round(d["value"] / 1e6, 2), "M leaves/s, e2e", round(d["e2e"]["value"] / 1e6, 2), d["phase_ms"], d.get("range_proofs", {}).get("n64_m1", {}).get("prove_plus_verify_per_s"))
    except Exception as e:
        pass
PY
