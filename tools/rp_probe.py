"""Range-proof throughput probe on the GPU box (prints JSON lines): prove / verify per shape and table window."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dapol_b200 import Context

SEED = hashlib.sha256(b"dapol-b200").digest()
shapes = [(64, 1, 8192), (64, 16, 1024), (64, 32, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
WINDOWS = [int(w) for w in os.environ.get("RP_WINDOWS", "12,8").split(",")]
for W in WINDOWS:
    ctx = Context(0, int(os.environ.get("COMB_WINDOW", "15")))
    ctx.set_rangeproof_window(W)
    for nbits, m, k in shapes:
        rng = np.random.default_rng(1)
        vals = rng.integers(0, 1 << 63, size=(k, m), dtype=np.uint64)
        bl = rng.integers(0, 256, size=(k, m, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
        streams = np.arange(k, dtype=np.uint64); bases = np.zeros(k, np.uint64)
        ctx.rangeproof_prove_batch(nbits, vals[:2], bl[:2], SEED, streams[:2], bases[:2])  # builds the tables
        tb = ctx.rangeproof_last_times()["table_build"]
        t0 = time.perf_counter(); proofs = ctx.rangeproof_prove_batch(nbits, vals, bl, SEED, streams, bases); t1 = time.perf_counter()
        pt = ctx.rangeproof_last_times()
        coms = np.stack([ctx.commit_batch(vals[:, j], bl[:, j]) for j in range(m)], axis=1)
        t2 = time.perf_counter(); ok = ctx.rangeproof_verify_batch(nbits, m, proofs, coms); t3 = time.perf_counter()
        vt = ctx.rangeproof_last_times()
        print(json.dumps({"W": W, "nbits": nbits, "m": m, "k": k, "all_ok": bool(ok.all()), "table_build_ms": round(tb, 1),
                          "prove_ms": {k_: round(v, 2) for k_, v in pt.items() if k_ != "table_build"}, "prove_wall_s": round(t1 - t0, 3),
                          "proofs_per_s": round(k / pt["total"] * 1e3, 1),
                          "verify_ms": {k_: round(v, 2) for k_, v in vt.items() if k_ != "table_build"}, "verify_wall_s": round(t3 - t2, 3),
                          "verifies_per_s": round(k / vt["total"] * 1e3, 1)}), flush=True)
    ctx.close()
