#!/bin/bash
# GPU session: parity tests, bench line, comb-window sweep, ncu launch list + full captures of the node kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
[ -n "$SKIP_TESTS" ] || timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
for W in 13 14 15 16; do
  timeout 300 python bench.py --comb-window $W --steps 4 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/bench_W$W.json 2>> gpurun_out/bench_W.err
done
timeout 300 python tools/microbench.py > gpurun_out/microbench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --users-log2 18 --height 30 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/ncu_bench.log 2>&1
for K in k_pad k_leaf; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${K}_full -f \
    python bench.py --steps 1 --warmup 0 --users-log2 18 --height 30 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/ncu_$K.log 2>&1
  ncu -i gpurun_out/${K}_full.ncu-rep --page raw --csv > gpurun_out/${K}_full_raw.csv 2>/dev/null
done
# k_merge: the 2nd launch is a full-size level (skip the first = leaf level)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge -s 2 -c 1 -o gpurun_out/k_merge_full -f \
  python bench.py --steps 1 --warmup 0 --users-log2 18 --height 30 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/ncu_k_merge.log 2>&1
ncu -i gpurun_out/k_merge_full.ncu-rep --page raw --csv > gpurun_out/k_merge_full_raw.csv 2>/dev/null
# the .ncu-rep files (with source) are ~50 MB each: keep the csv pages only
for K in k_pad k_leaf k_merge; do ncu -i gpurun_out/${K}_full.ncu-rep --page source --csv > gpurun_out/${K}_full_source.csv 2>/dev/null; rm -f gpurun_out/${K}_full.ncu-rep; done
ls -la gpurun_out
