import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dapol_b200 import Context
ctx = Context(0); ctx.set_rangeproof_window(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
vals = np.array([[5], [9]], np.uint64); bl = np.zeros((2, 1, 32), np.uint8); bl[:, 0, 0] = 3
p = ctx.rangeproof_prove_batch(64, vals, bl, hashlib.sha256(b"x").digest(), [0, 1], [0, 0])
print(p[0, :32].tobytes().hex())
