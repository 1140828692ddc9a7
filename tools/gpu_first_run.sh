#!/bin/bash
# First GPU session: parity tests, integer-pipe microbenchmark, comb-window sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python tools/microbench.py 2>&1 | tee gpurun_out/microbench.txt
