#!/bin/bash
# Session D, call 8: hybrid inner-product rounds: parity + throughput of the aggregated shapes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s10_pytest_rp.txt
RP_WINDOWS=0 timeout 900 python tools/rp_probe.py 64x32x2048 64x16x2048 64x64x1024 64x1x16384 2> gpurun_out/s10_rp.err | tee gpurun_out/s10_rp.txt | cut -c1-330
tail -2 gpurun_out/s10_rp.err
timeout 900 python tools/c3_all.py 20 32 0 8192 16384 > gpurun_out/s10_c3_1gpu.json 2> gpurun_out/s10_c3_1gpu.err; cat gpurun_out/s10_c3_1gpu.json | cut -c1-700; tail -2 gpurun_out/s10_c3_1gpu.err
