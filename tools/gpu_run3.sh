#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
bash tools/fe_variants.sh
