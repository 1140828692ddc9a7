#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- full-size golden roots from the C oracle (oracle/c, OpenMP), committed as
tests/golden/full_size_golden.json and asserted by tests/test_gpu_fullsize.py and bench.py.

  C2: the exact bench.py workload (2^20 synthetic liabilities, height 32, D = blake3, bench seeds):
      root record, per-level node / padding counts, sha256 of every level's (com || hash) bytes.
  C4 shard: BASELINE config 4 (2^24 liabilities, height 40) split by the 3-bit leaf-index prefix as
      the 8-GPU build does; the oracle builds the height-37 subtree of prefix 0 with the padding
      blocks the single-tree creation order gives it (cref.Tree(level_base=...)).

Run here (no GPU): python tests/golden/gen_golden_full.py [c2] [c4]     (~3 min + ~12 min on 8 cores)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import AUDIT_SEED, PAD_SEED, synth_liabilities  # noqa: E402
from oracle import cref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "full_size_golden.json")


def level_pad_counts(idx_sorted: np.ndarray, height: int) -> np.ndarray:
    """counts[h] = padding nodes at level h of the tree over the sorted leaf indexes (h = 0 .. height)."""
    counts = np.zeros(height + 1, np.uint64)
    cur = idx_sorted.astype(np.uint64)
    for h in range(height, 0, -1):
        par = cur >> np.uint64(1)
        newp = np.ones(len(cur), bool)
        newp[1:] = par[1:] != par[:-1]
        n_par = int(newp.sum())
        counts[h] = 2 * n_par - len(cur)
        cur = par[newp]
    return counts


def tree_summary(t, height):
    root = t.root()
    levels = []
    for h in range(height + 1):
        lv = t.level(h)
        d = hashlib.sha256()
        d.update(lv["comc"].tobytes())
        d.update(lv["hash"].tobytes())
        levels.append({"nodes": int(len(lv["idx"])), "pads": int(lv["is_pad"].sum()), "sha256_com_hash": d.hexdigest()})
    return {"root": {"com": root["comc"].hex(), "hash": root["hash"].hex(), "value": root["v"], "blinding": root["r"].hex()},
            "num_pads": int(t.num_pads), "levels": levels}


def gen_c2(users_log2=20, H=32):
    n = 1 << users_log2
    iid, io, eid, eo, vals = synth_liabilities(n)
    t0 = time.time()
    rc, idx, bl, _ = cref.derive_leaves(0, iid, io, eid, eo, AUDIT_SEED, H)
    assert rc == 0
    order = np.argsort(idx, kind="stable")
    t = cref.Tree(0, H, idx[order], vals[order], bl[order], PAD_SEED, 0, os.cpu_count())
    out = tree_summary(t, H)
    out.update({"users_log2": users_log2, "height": H, "hash_id": 0, "oracle_seconds": round(time.time() - t0, 1),
                "sample_leaf_index_of": {str(p): int(idx[p]) for p in (0, 1, 12345, n - 1)}})
    return out


def gen_c4_shard(users_log2=24, H=40, k=3, prefix=0):
    n = 1 << users_log2
    iid, io, eid, eo, vals = synth_liabilities(n)
    t0 = time.time()
    rc, idx, bl, _ = cref.derive_leaves(0, iid, io, eid, eo, AUDIT_SEED, H)
    assert rc == 0
    order = np.argsort(idx, kind="stable")
    idx_s = idx[order]
    Hs = H - k
    shard_of = idx_s >> np.uint64(Hs)
    counts_all = np.zeros((1 << k, Hs + 1), np.uint64)
    for r in range(1 << k):
        counts_all[r] = level_pad_counts(idx_s[shard_of == r] & np.uint64((1 << Hs) - 1), Hs)
    from dapol_b200.sharded import shard_pad_bases
    level_base, top_base = shard_pad_bases(counts_all, prefix, 0)
    sel = order[shard_of == prefix]
    t = cref.Tree(0, Hs, idx[sel] & np.uint64((1 << Hs) - 1), vals[sel], bl[sel], PAD_SEED, 0, os.cpu_count(), level_base=level_base)
    out = tree_summary(t, Hs)
    out.update({"users_log2": users_log2, "height": H, "prefix_bits": k, "prefix": prefix, "shard_leaves": int(len(sel)),
                "pad_counts_all_shards": counts_all.tolist(), "level_base": [int(x) for x in level_base], "top_pad_base": int(top_base),
                "hash_id": 0, "oracle_seconds": round(time.time() - t0, 1)})
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4"]
    cref.build()
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    if "c2" in which:
        data["c2_2p20_h32"] = gen_c2()
        json.dump(data, open(OUT, "w"), indent=1)
        print("c2 done", data["c2_2p20_h32"]["root"]["com"][:16], data["c2_2p20_h32"]["oracle_seconds"], flush=True)
    if "c4" in which:
        data["c4_2p24_h40_shard0"] = gen_c4_shard()
        json.dump(data, open(OUT, "w"), indent=1)
        print("c4 shard done", data["c4_2p24_h40_shard0"]["root"]["com"][:16], data["c4_2p24_h40_shard0"]["oracle_seconds"], flush=True)
