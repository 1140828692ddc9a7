#!/bin/bash
# GPU session A: parity tests, smoke, microbench, bench (N=1), reference arm, ncu launch list + full capture of k_pad.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 600 python tools/microbench.py 2>&1 | tee gpurun_out/microbench.txt
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
timeout 600 python bench.py --impl reference --steps 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --users-log2 18 --height 30 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pad -s 1 -c 1 -o gpurun_out/prof_k_pad \
    python bench.py --steps 1 --warmup 1 --users-log2 16 --height 28 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 bash tools/fe_variants.sh > gpurun_out/fe_variants.log 2>&1
ls -la gpurun_out
