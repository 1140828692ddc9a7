#!/bin/bash
# GPU session 3: range-proof parity on the new CTA reduction, inversion-batch sweep, range-proof probe, ncu of the prover.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rangeproof.py tests/test_gpu_inclusion.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_rp.txt
for v in b16 b20 b24 b32 b32m3; do
  DAPOL_B200_LIB=$PWD/dapol_b200/lib/var_$v.so timeout 300 python bench.py --steps 6 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/var_$v.json 2>> gpurun_out/var.err
done
timeout 600 python tools/rp_probe.py 64x1x16384 64x8x2048 64x32x512 > gpurun_out/rp_probe.txt 2> gpurun_out/rp_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/rp_launches.csv \
  python tools/rp_min.py > gpurun_out/ncu_rp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rp_p10 -s 20 -c 1 -o gpurun_out/k_rp_p10_full -f \
  python tools/rp_min.py > gpurun_out/ncu_rp_p10.log 2>&1
ncu -i gpurun_out/k_rp_p10_full.ncu-rep --page raw --csv > gpurun_out/k_rp_p10_full_raw.csv 2>/dev/null
ncu -i gpurun_out/k_rp_p10_full.ncu-rep --page source --csv > gpurun_out/k_rp_p10_full_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -15
