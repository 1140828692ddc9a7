"""Batched range-proof verification probe for ncu captures: K valid proofs of one shape, verified per proof (Straus) and per group
(bucket method).   python tools/verify_probe.py [nbits x m x K = 64x1x16384] [modes = 256]
modes: comma list of group[:window_bits] (window_bits 0 / absent = the library's choice, log2(group * nv) - 3)"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dapol_b200 import Context

SEED = hashlib.sha256(b"dapol-b200").digest()
nbits, m, k = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64x1x16384").split("x"))
modes = [(int(a.split(":")[0]), int(a.split(":")[1]) if ":" in a else 0) for a in (sys.argv[2] if len(sys.argv) > 2 else "256").split(",")]
ctx = Context(0, 15)
rng = np.random.default_rng(1)
vals = rng.integers(0, 1 << (nbits - 1), size=(k, m), dtype=np.uint64)
bl = rng.integers(0, 256, size=(k, m, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
proofs = ctx.rangeproof_prove_batch(nbits, vals, bl, SEED, np.arange(k, dtype=np.uint64), np.zeros(k, np.uint64))
coms = np.stack([ctx.commit_batch(vals[:, j], bl[:, j]) for j in range(m)], axis=1)
out = {}
for mode, cbits in [(0, 0)] + modes:
    ctx.set_verify_mode(mode, cbits)
    ok = ctx.rangeproof_verify_batch(nbits, m, proofs, coms)
    ok = ctx.rangeproof_verify_batch(nbits, m, proofs, coms)
    t = ctx.rangeproof_last_kernel_times()
    out[f"group_{mode}_c{cbits}"] = {"all_ok": bool(ok.all()), "total_ms": round(t["total"], 3), "verifier_ms": round(t["verifier"], 3), "other_ms": round(t["other"], 3),
                            "verifies_per_s": round(k / t["total"] * 1e3)}
print(json.dumps({"nbits": nbits, "m": m, "k": k, **out}))
ctx.close()
