#!/bin/bash
# Round 2, GPU call R (one GPU): same-box A/B of the prover changes at the bench shapes -- compress pass per point vs per proof,
# packed small-shape MSM kernels at 0 / 4 / 8 / 16 lanes -- then parity of the range-proof tests with the packed kernels.
mkdir -p gpurun_out
P=gpurun_out/r02r
run() { env "$@" RP_WINDOWS=0 COMB_WINDOW=15 timeout 600 python tools/rp_probe.py 64x1x16384 64x1x65536 64x2x8192 64x32x2048 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['m'], d['k'], 'prove', d['prove_ms'], round(d['proofs_per_s']), 'verify', round(d['verifies_per_s']))"; }
for rep in 1 2; do
  echo "== compress per point, pack 0"; run DAPOL_RP_PACK_LANES=0
  echo "== compress per proof (old arrangement), pack 0"; run DAPOL_RP_COMPRESS_SEQ=1 DAPOL_RP_PACK_LANES=0
  echo "== pack 8"; run DAPOL_RP_PACK_LANES=8
  echo "== pack 16"; run DAPOL_RP_PACK_LANES=16
  echo "== pack 4"; run DAPOL_RP_PACK_LANES=4
done 2>&1 | tee ${P}_ab.txt
timeout 900 python -m pytest tests/test_gpu_rangeproof.py tests/test_golden.py -m gpu -q 2>&1 | tail -4 | tee ${P}_pytest_gpu.txt
