#!/bin/bash
# GPU session 2: tests, smoke, bench (N=1), ncu launch list + one full capture of the padding kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.txt
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.txt
python bench.py --impl reference --steps 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --users-log2 18 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pad -s 1 -c 1 -o gpurun_out/prof_k_pad \
    python bench.py --steps 1 --warmup 1 --users-log2 16 --height 28 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
