#!/bin/bash
# GPU session C: range-proof parity tests + throughput probe.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rangeproof.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_rp.txt
timeout 900 python tools/rp_probe.py 2>&1 | tee gpurun_out/rp_probe.txt
