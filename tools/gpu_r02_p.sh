#!/bin/bash
# Round 2, GPU call P (one GPU): parallel Merkle fold of the inclusion-proof verifier -- parity of every proof test, C1 latency, smoke.
mkdir -p gpurun_out
P=gpurun_out/r02p
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 1500 python -m pytest tests/test_gpu_inclusion.py tests/test_gpu_blake2b.py tests/test_gpu_update.py tests/test_gpu_ids.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_persist.py tests/test_golden.py -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
timeout 900 python bench.py --rp-singles 0 --rp-aggregates 0 --steps 3 > ${P}_bench_c1.json 2> ${P}_bench_c1.err; tail -3 ${P}_bench_c1.err
python -c "
import json
d = json.loads([l for l in open('${P}_bench_c1.json') if l.startswith('{')][-1]); print(json.dumps(d['c1'])); print(d['value'], d.get('gpu_root_matches'))"
VERIFY_GROUP=256 timeout 900 python tools/northstar.py 20 32 0 8192 32768 8 > ${P}_c3_sample_g256.json 2> ${P}_c3_sample_g256.err; tail -2 ${P}_c3_sample_g256.err
python -c "
import json
d = json.loads([l for l in open('${P}_c3_sample_g256.json') if l.startswith('{')][-1])
print(d['all_verified'], d['tampered_rejected'], d['oracle_verified_sample'], round(d['prove_per_s']), round(d['verify_per_s']), d['rank0_chunk_s_prove_write_verify'])"
