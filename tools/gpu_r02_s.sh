#!/bin/bash
# Round 2, GPU call S (one GPU): packed small-shape MSM kernels (lanes per MSM) vs a warp per MSM, same box, back to back; then parity with forced packing.
mkdir -p gpurun_out
P=gpurun_out/r02s
run() { env "$@" RP_WINDOWS=0 COMB_WINDOW=15 timeout 600 python tools/rp_probe.py 64x1x16384 64x1x65536 64x2x16384 32x1x32768 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['nbits'], d['m'], d['k'], 'prove', d['prove_ms'], round(d['proofs_per_s']), 'verify', round(d['verifies_per_s']), d['all_ok'])"; }
for rep in 1 2; do
  echo "== warp per MSM (pack 0)"; run DAPOL_RP_PACK_LANES=0
  echo "== pack 8"; run DAPOL_RP_PACK_LANES=8
  echo "== pack 16"; run DAPOL_RP_PACK_LANES=16
  echo "== pack 4"; run DAPOL_RP_PACK_LANES=4
done 2>&1 | tee ${P}_pack_ab.txt
timeout 900 python -m pytest tests/test_gpu_rangeproof.py -m gpu -q 2>&1 | tail -4 | tee ${P}_pytest_gpu.txt
