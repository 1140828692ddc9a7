#!/bin/bash
# Round 2, GPU call Q (one GPU): prover points compressed side by side (k_rp_compress_pts), verifier's fixed-base MSM split for small
# batches -- parity of the range-proof / inclusion / batch-proof tests, single-proof latency, throughput at the bench shapes.
mkdir -p gpurun_out
P=gpurun_out/r02q
timeout 1500 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_gpu_batch_proof.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
RP_WINDOWS=0 COMB_WINDOW=15 timeout 600 python tools/rp_probe.py 64x16x1 64x1x1 64x32x1 64x64x1 64x1x16384 64x32x2048 > ${P}_rp_probe.txt 2> ${P}_rp_probe.err; tail -3 ${P}_rp_probe.err
cat ${P}_rp_probe.txt
