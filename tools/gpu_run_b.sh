#!/bin/bash
# GPU session B: parity tests + bench after a kernel change (short).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
