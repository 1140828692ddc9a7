#!/bin/bash
# Round 2, GPU call O (one GPU): the whole GPU parity suite + smoke on the final library (after the id / salt leaf-hash fix).
mkdir -p gpurun_out
P=gpurun_out/r02o
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${P}_smoke.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee ${P}_pytest_gpu.txt
