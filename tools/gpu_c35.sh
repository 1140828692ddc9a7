#!/bin/bash
# 8-GPU box: config 3 with the final prover (all 2^20 inclusion proofs proved + verified) and config 5 (verification sweep) at 8 / 4 / 2 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29531 tools/c3_all.py 20 32 0 8192 > gpurun_out/c3b_all_p0.json 2> gpurun_out/c3b_all_p0.err
tail -n 1 gpurun_out/c3b_all_p0.json | cut -c1-1100; tail -c 300 gpurun_out/c3b_all_p0.err
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port 2954$n tools/c5_verify_sweep.py 20 > gpurun_out/c5_n$n.jsonl 2> gpurun_out/c5_n$n.err
  python - <<P
import json
for l in open("gpurun_out/c5_n$n.jsonl"):
    try:
        d = json.loads(l); print($n, d["nbits"], d["proofs_per_gpu"], round(d["verifies_per_s"]/1e6, 3), "M/s", d["verdicts_exact"])
    except Exception: pass
P
  tail -c 200 gpurun_out/c5_n$n.err
done
