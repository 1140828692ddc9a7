"""Smallest useful range-proof run for ncu: one m = 32 batch (K = 256) and one m = 1 batch (K = 8192), prove + verify.
   python tools/rp_min.py [window=12]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dapol_b200 import Context
ctx = Context(0); ctx.set_rangeproof_window(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
seed = hashlib.sha256(b"x").digest()
for m, k in ((32, 256), (1, 8192)):
    rng = np.random.default_rng(m)
    vals = rng.integers(0, 1 << 63, size=(k, m), dtype=np.uint64)
    bl = rng.integers(0, 256, size=(k, m, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
    p = ctx.rangeproof_prove_batch(64, vals, bl, seed, np.arange(k, dtype=np.uint64), np.zeros(k, np.uint64))
    coms = np.stack([ctx.commit_batch(vals[:, j], bl[:, j]) for j in range(m)], axis=1)
    ok = ctx.rangeproof_verify_batch(64, m, p, coms)
    print(m, k, bool(ok.all()), p[0, :16].tobytes().hex())
