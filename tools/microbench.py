"""Integer-pipe peak + field-op throughput + comb-window sweep on the GPU box (prints JSON lines)."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dapol_b200 import Context, Dapol

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
ctx = Context(0)
res = {}
for v, name in enumerate(["imad_lo_Ginstr_s", "imad_wide_GMAC32_s", "madlo_madhi_cc_GMAC32_s"]):
    res[name] = round(ctx.imad_peak(v), 1)
res["fe_mul_Gop_s"] = round(ctx.fe_bench(0), 2)
res["fe_sq_Gop_s"] = round(ctx.fe_bench(1), 2)
print(json.dumps({"microbench": res}))
ctx.close()
H, n = 32, 1 << 18
rng = np.random.default_rng(1)
idx = np.unique(rng.integers(0, 1 << H, size=n, dtype=np.uint64))
n = len(idx)
vals = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
bl = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); bl[:, 31] &= 0x7F
for W in (8, 12, 13, 14, 15, 16):
    ctx = Context(0, W)
    best = None
    for rep in range(3):
        t = Dapol.new_blank(ctx, 0, H, H).build(idx, vals, bl, PAD_SEED)
        ms = ctx.last_build_times()
        if best is None or ms["total"] < best["total"]:
            best = ms
        nodes, pads = t.num_nodes, t.num_padding
        root = t.root_raw().com.hex()
        t.close()
    print(json.dumps({"W": W, "n": n, "nodes": nodes, "pads": pads, "ms": {k: round(v, 3) for k, v in best.items()},
                      "leaves_per_s": round(n / best["total"] * 1e3), "root": root[:16]}))
    ctx.close()
