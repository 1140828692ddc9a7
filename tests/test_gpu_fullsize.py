"""Full-size parity (BASELINE configs 2, 3, 4) against goldens the C oracle produced on the CPU
(tests/golden/gen_golden_full.py -> tests/golden/full_size_golden.json), through the C ABI:

  C2  2^20 users / height 32, the exact bench.py workload: root record, node and padding counts and the sha256 of
      every level's commitments and hashes equal the oracle's build;
  C3  1024 random inclusion proofs (aggregation 32: one m = 32 aggregated Bulletproof each) of that tree are verified by
      the ORACLE's verifier against the oracle's root;
  C4  2^24 users / height 40: the height-37 subtree of leaf-index prefix 0 (what GPU 0 of the 8-GPU build owns), built with
      the padding blocks the single-tree creation order assigns it, equals the oracle's subtree root and levels;
  m = 32 / m = 64 range-proof BYTES for 16 / 4 proofs (the oracle proves them on a thread pool).
"""
import concurrent.futures as cf
import hashlib
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "full_size_golden.json")))
SEED = hashlib.sha256(b"dapol-b200").digest()


def _level_digest(lv):
    d = hashlib.sha256()
    d.update(lv["comc"].tobytes())
    d.update(lv["hash"].tobytes())
    return d.hexdigest()


def _check_tree(level_fn, num_pads, g, H):
    assert num_pads == g["num_pads"]
    for h in range(H + 1):
        lv = level_fn(h)
        want = g["levels"][h]
        assert (len(lv["idx"]), int(lv["is_pad"].sum())) == (want["nodes"], want["pads"]), h
        assert _level_digest(lv) == want["sha256_com_hash"], f"level {h} differs from the oracle's"


@pytest.fixture(scope="module")
def c2():
    from bench import AUDIT_SEED, PAD_SEED, synth_liabilities
    from dapol_b200 import Context, Dapol
    g = GOLD["c2_2p20_h32"]
    c = Context(0)
    c.set_rangeproof_window(12)
    n, H = 1 << g["users_log2"], g["height"]
    t = Dapol.new(c, 0, synth_liabilities(n), AUDIT_SEED, H, 32, PAD_SEED)
    yield c, t, g, n, H
    t.close(); c.close()


def test_c2_root_and_every_level_equal_oracle_golden(c2):
    c, t, g, n, H = c2
    r = t.root_raw()
    assert (r.com.hex(), r.hash.hex(), r.value, r.blinding.hex()) == (g["root"]["com"], g["root"]["hash"], g["root"]["value"], g["root"]["blinding"])
    for pos, want in g["sample_leaf_index_of"].items():
        assert t.leaf_index_of(int(pos)) == want
    _check_tree(t.level, t.num_padding, g, H)


def test_c3_sample_of_inclusion_proofs_verified_by_oracle(c2, cref):
    """1024 random users of the 2^20: proofs generated on the GPU (Padding policy, aggregation 32), verified by the oracle's
    DapolProof::verify restatement against the ORACLE's root (the golden); a proof paired with another leaf is rejected."""
    c, t, g, n, H = c2
    root_com, root_hash = bytes.fromhex(g["root"]["com"]), bytes.fromhex(g["root"]["hash"])
    users = np.random.default_rng(20261017).choice(n, 1024, replace=False)
    picks = [t.leaf_index_of(int(u)) for u in users]
    proofs = t.generate_proofs(picks, hashlib.sha256(b"c3-sample").digest())
    paths = t.paths(picks)

    def check(k):
        lc, lh = paths["leaf_comc"][k].tobytes(), paths["leaf_hash"][k].tobytes()
        return cref.verify_inclusion(0, 0, proofs[k].serialize(), root_com, root_hash, lc, lh)
    with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        ok = list(ex.map(check, range(len(picks))))
    assert all(ok), f"{ok.count(False)} of {len(ok)} GPU proofs rejected by the oracle"
    other = paths["leaf_comc"][1].tobytes()
    assert not cref.verify_inclusion(0, 0, proofs[0].serialize(), root_com, root_hash, other, paths["leaf_hash"][0].tobytes())
    # and the GPU verifier agrees on the same batch, with one tampered proof
    from dapol_b200 import DapolProof, DapolProofNode
    leaves = [DapolProofNode(paths["leaf_comc"][k].tobytes(), paths["leaf_hash"][k].tobytes()) for k in range(len(picks))]
    bad = bytearray(proofs[5].serialize()); bad[100] ^= 4
    proofs[5] = DapolProof(bytes(bad), 0, 0)
    got = DapolProof.verify_many(c, DapolProofNode(root_com, root_hash), leaves, proofs)
    assert not got[5] and got.sum() == len(picks) - 1


def test_c4_shard0_subtree_equals_oracle_golden():
    """BASELINE config 4's leaf derivation over all 2^24 users, the 3-bit prefix split, the per-level padding bases and the
    height-37 subtree GPU 0 of the 8-GPU build owns: root record and every level equal the oracle's shard build."""
    import torch
    from bench import AUDIT_SEED, PAD_SEED, synth_liabilities
    from dapol_b200 import Context, CudaEngine
    from dapol_b200.sharded import shard_pad_bases
    g = GOLD["c4_2p24_h40_shard0"]
    n, H, k = 1 << g["users_log2"], g["height"], g["prefix_bits"]
    Hs = H - k
    c = Context(0)
    E = CudaEngine(c)
    iid, io, eid, eo, vals = synth_liabilities(n)
    audit, seedst, cand, blind = E.derive(0, H, n, E.to_dev(iid, np.uint8), E.to_dev(io, np.int64), E.to_dev(eid, np.uint8), E.to_dev(eo, np.int64),
                                          AUDIT_SEED)
    d_vals = E.to_dev(vals, np.int64)
    counts_all, shard0 = [], None
    for r in range(1 << k):
        idx, v, bl = E.assign(0, H, n, audit, seedst, cand, blind, d_vals, k, r, n // (1 << k) + n // 16)
        counts_all.append(E.pad_counts(Hs, idx))
        if r == g["prefix"]:
            shard0 = (idx, v, bl)
    assert np.array(counts_all).tolist() == g["pad_counts_all_shards"]
    level_base, top_base = shard_pad_bases(np.array(counts_all), g["prefix"], 0)
    assert [int(x) for x in level_base] == g["level_base"] and int(top_base) == g["top_pad_base"]
    assert len(shard0[0]) == g["shard_leaves"]
    h = E.build_shard(0, Hs, *shard0, PAD_SEED, level_base)
    rec = E.root_record(h)
    assert rec[128:160].tobytes().hex() == g["root"]["com"] and rec[160:192].tobytes().hex() == g["root"]["hash"]
    assert rec[192:224].tobytes().hex() == g["root"]["blinding"] and int.from_bytes(rec[224:232].tobytes(), "little") == g["root"]["value"]
    L = E.L
    import ctypes as C

    def level(hh):
        m = L.dapol_tree_level_size(h, hh)
        idx = np.zeros(m, np.uint64); cc = np.zeros((m, 32), np.uint8); hs = np.zeros((m, 32), np.uint8); pad = np.zeros(m, np.uint8)
        assert L.dapol_tree_level_copy(h, hh, idx.ctypes.data_as(C.c_void_p), None, None, cc.ctypes.data_as(C.c_void_p),
                                       hs.ctypes.data_as(C.c_void_p), pad.ctypes.data_as(C.c_void_p)) == 0
        return dict(idx=idx, comc=cc, hash=hs, is_pad=pad)
    _check_tree(level, L.dapol_tree_num_padding(h), g, Hs)
    E.destroy(h)
    del audit, seedst, cand, blind, d_vals, shard0
    torch.cuda.empty_cache()
    c.close()


@pytest.mark.parametrize("m,k", [(32, 16), (64, 4)])
def test_large_aggregate_proof_bytes_equal_oracle(cref, m, k):
    """Every proof of the batch, not only the first: byte-identical with the oracle's prover at m = 32 (the aggregated proof of a
    height-32 inclusion proof) and m = 64 (height 40), hybrid inner-product rounds included (auto window)."""
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(12)
    rnd = random.Random(m * 7 + k)
    vals = np.array([[rnd.randrange(1 << 64) for _ in range(m)] for _ in range(k)], np.uint64)
    vals[0, 0] = (1 << 64) - 1; vals[-1, -1] = 0
    bl = np.frombuffer(rnd.randbytes(32 * m * k), np.uint8).copy().reshape(k, m, 32)
    bl[:, :, 31] &= 0x7F
    bl[0, 1] = 0; bl[0, 1, 0] = 1  # the Padding policy's dummy party (0, Scalar::one())
    vals[0, 1] = 0
    streams = np.array([rnd.randrange(1 << 64) for _ in range(k)], np.uint64)
    bases = np.array([rnd.randrange(1 << 31) << 32 for _ in range(k)], np.uint64)
    proofs = c.rangeproof_prove_batch(64, vals, bl, SEED, streams, bases)

    def want(i):
        return cref.rp_prove([int(x) for x in vals[i]], [b.tobytes() for b in bl[i]], SEED, int(streams[i]), int(bases[i]), 64)
    with cf.ThreadPoolExecutor(min(k, os.cpu_count() or 4)) as ex:
        wants = list(ex.map(want, range(k)))
    for i in range(k):
        assert proofs[i].tobytes() == wants[i], i
    c.close()
