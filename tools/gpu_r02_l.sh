#!/bin/bash
# Round 2, GPU call L (one GPU): k_pad with the comb's mixed additions inlined (no call-ABI moves in the hot loop), at 4 and 3 CTAs/SM,
# against the default build -- tree bench only, same box, back to back.
mkdir -p gpurun_out
P=gpurun_out/r02l
for v in default padinl4 padinl3 $EXTRA_VARIANTS; do
  lib=dapol_b200/lib/var_$v.so; [ $v = default ] && lib=dapol_b200/lib/libdapol_b200.so
  DAPOL_B200_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --no-c1 --rp-singles 0 --rp-aggregates 0 --steps 6 > ${P}_bench_$v.json 2> ${P}_bench_$v.err; tail -2 ${P}_bench_$v.err
  python -c "
import json
d = json.loads([l for l in open('${P}_bench_$v.json') if l.startswith('{')][-1])
print('$v', round(d['value']/1e6,2), {k: round(x,2) for k,x in d['phase_ms'].items()}, round(d['roofline']['frac'],3), d.get('gpu_root_matches'), d.get('root'))"
done
