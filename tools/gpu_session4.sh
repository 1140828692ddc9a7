#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_golden.py tests/test_gpu_tree.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_rp.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/rp_probe.py 64x1x16384 64x8x2048 64x32x512 64x32x2048 > gpurun_out/rp_probe.txt 2> gpurun_out/rp_probe.err
for spec in "m32 3" "m1 13"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rp_p10 -s $2 -c 1 -o gpurun_out/k_rp_p10_$1 -f \
    python tools/rp_min.py > gpurun_out/ncu_rp_p10_$1.log 2>&1
  ncu -i gpurun_out/k_rp_p10_$1.ncu-rep --page raw --csv > gpurun_out/k_rp_p10_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/k_rp_p10_$1.ncu-rep --page source --csv > gpurun_out/k_rp_p10_$1_source.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -12
