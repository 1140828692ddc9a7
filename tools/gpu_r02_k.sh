#!/bin/bash
# Round 2, GPU call K (one GPU): small-batch prover (split MSMs, table rounds only) -- parity of every small-K shape, single-proof
# latency, C1 leg; smoke with both verifier modes; C3-shaped sample (2^20 users / H = 32, 32768 proofs) per verifier mode.
mkdir -p gpurun_out
P=gpurun_out/r02k
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 1500 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_gpu_batch_proof.py tests/test_golden.py -m gpu -q 2>&1 | tail -15 | tee ${P}_pytest_gpu.txt
RP_WINDOWS=0 COMB_WINDOW=15 timeout 600 python tools/rp_probe.py 64x16x1 64x1x1 64x32x1 64x64x1 64x16x8 64x32x64 > ${P}_rp_probe.txt 2> ${P}_rp_probe.err; tail -3 ${P}_rp_probe.err
cat ${P}_rp_probe.txt
timeout 900 python bench.py --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 --steps 3 > ${P}_bench_c1.json 2> ${P}_bench_c1.err; tail -3 ${P}_bench_c1.err
python -c "
import json
d = json.loads([l for l in open('${P}_bench_c1.json') if l.startswith('{')][-1]); print(json.dumps(d['c1']))"
for G in 0 256; do
  VERIFY_GROUP=$G timeout 900 python tools/northstar.py 20 32 0 8192 32768 8 > ${P}_c3_sample_g$G.json 2> ${P}_c3_sample_g$G.err; tail -2 ${P}_c3_sample_g$G.err
  python -c "
import json
d = json.loads([l for l in open('${P}_c3_sample_g$G.json') if l.startswith('{')][-1])
print($G, d['all_verified'], d['tampered_rejected'], d['oracle_verified_sample'], round(d['prove_per_s']), round(d['verify_per_s']), d['rank0_chunk_s_prove_write_verify'], d['verify_fallbacks_rank0'])"
done
