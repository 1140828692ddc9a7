#!/bin/bash
# GPU session 2: parity tests on the current library, kernel-variant sweep (launch bounds, comb prefetch, inversion batch),
# full bench line, C5 verification sweep, C3 inclusion-proof sample.
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
for v in base nopf minb3 minb5 b6 b12 b16; do
  lib=dapol_b200/lib/var_$v.so; [ $v = base ] && lib=dapol_b200/lib/libdapol_b200.so
  DAPOL_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 6 --no-cpu-baseline --rp-singles 0 --rp-aggregates 0 > gpurun_out/var_$v.json 2>> gpurun_out/var.err
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 900 python tools/c5_verify_sweep.py 20 > gpurun_out/c5_sweep.jsonl 2> gpurun_out/c5.err; tail -3 gpurun_out/c5.err
timeout 900 python tools/c3_inclusion.py 20 32 2048 > gpurun_out/c3_inclusion.jsonl 2> gpurun_out/c3.err; tail -3 gpurun_out/c3.err
ls -la gpurun_out | head -40
