"""GPU parity tests (run on the B200 box): the CUDA tree build, called through the C ABI, against the
CPU oracle (oracle/c via ctypes) on the same seeded inputs -- bit-exact for every node of every level."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _inputs(H, n, seed, stride=False):
    rnd = random.Random(seed)
    if stride:  # benches/dapol.rs:160-175 layout: leaf i at i * 2^H / n
        idx = np.arange(n, dtype=np.uint64) * np.uint64((1 << H) // n)
    else:
        idx = np.array(sorted(rnd.sample(range(1 << H), n)) if H < 40 else sorted({rnd.randrange(1 << H) for _ in range(n)}), np.uint64)
    n = len(idx)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32)
    bl[:, 31] &= 0x7F  # Scalar::from_bits domain (possibly unreduced)
    return idx, vals, bl


def _assert_same_tree(gpu, oracle_tree, H):
    assert gpu.num_padding == oracle_tree.num_pads
    for h in range(H + 1):
        g, o = gpu.level(h), oracle_tree.level(h)
        assert len(g["idx"]) == len(o["idx"]), h
        for key in ("idx", "v", "comc", "hash", "is_pad"):
            assert (g[key] == o[key]).all(), (h, key)
        L = 2 ** 252 + 27742317777372353535851937790883648493
        gr = [int.from_bytes(x.tobytes(), "little") % L for x in g["r"]]
        orr = [int.from_bytes(x.tobytes(), "little") % L for x in o["r"]]
        assert gr == orr, h


def test_commit_batch_vs_oracle(ctx, cref):
    rnd = random.Random(11)
    vals = [0, 1, 5, 2 ** 64 - 1] + [rnd.randrange(1 << 64) for _ in range(300)]
    bls = [(1).to_bytes(32, "little"), (7).to_bytes(32, "little"), bytes(32), ((1 << 255) - 1).to_bytes(32, "little")]
    bls += [(rnd.randrange(1 << 255)).to_bytes(32, "little") for _ in range(300)]
    out = ctx.commit_batch(vals, np.frombuffer(b"".join(bls), np.uint8))
    for v, b, c in zip(vals, bls, out):
        assert c.tobytes() == cref.commit(v, b)
    assert out[1].tobytes().hex() != "" and ctx.commit_batch([5], np.frombuffer((7).to_bytes(32, "little"), np.uint8))[0].tobytes().hex() == \
        "84dcc85db7eef17103ea879c4900162127debe4b41a8f06012a25911292aff18"


@pytest.mark.parametrize("hash_id,H,n,stride", [
    (0, 6, 9, False), (1, 4, 4, False), (0, 10, 200, False), (0, 16, 1024, True), (0, 16, 1024, False),
    (1, 12, 300, False), (0, 3, 4, False), (0, 1, 1, False), (0, 1, 2, False), (0, 5, 1, False),
    (0, 32, 2048, True), (0, 40, 500, False), (0, 64, 100, False), (0, 0, 1, False)])
def test_tree_every_node_vs_oracle(ctx, cref, hash_id, H, n, stride):
    from dapol_b200 import Dapol
    idx, vals, bl = _inputs(H, n, 1000 + H * 7 + n, stride)
    if H == 0:
        idx = np.zeros(1, np.uint64)
    gpu = Dapol.new_blank(ctx, hash_id, H, H).build(idx, vals, bl, PAD_SEED, 5)
    ora = cref.Tree(hash_id, H, idx, vals, bl, PAD_SEED, 5)
    _assert_same_tree(gpu, ora, H)
    r, o = gpu.root_raw(), ora.root()
    assert (r.value, r.com, r.hash) == (o["v"], o["comc"], o["hash"])
    assert r.value == int(vals.sum()) % (1 << 64)  # liability-sum homomorphism (src/dapol/tests.rs:24)
    gpu.close()


@pytest.mark.parametrize("W", [4, 8, 12, 15, 16, 20, 22, 24, 26])
def test_every_comb_window_gives_the_same_tree(cref, W):
    """The comb window of the fixed-base tables is a speed knob only: every instantiated width gives the oracle's tree."""
    from dapol_b200 import Context, Dapol
    c = Context(0, W)
    H, n = 12, 200
    idx, vals, bl = _inputs(H, n, 4242)
    vals[:3] = [0, 2 ** 64 - 1, 1]
    gpu = Dapol.new_blank(c, 0, H, H).build(idx, vals, bl, PAD_SEED, 1)
    _assert_same_tree(gpu, cref.Tree(0, H, idx, vals, bl, PAD_SEED, 1), H)
    out = c.commit_batch([0, 1, 2 ** 64 - 1], np.frombuffer(bytes(32) + (1).to_bytes(32, "little") + ((1 << 255) - 19).to_bytes(32, "little"), np.uint8))
    assert [x.tobytes() for x in out] == [cref.commit(0, bytes(32)), cref.commit(1, (1).to_bytes(32, "little")),
                                          cref.commit(2 ** 64 - 1, ((1 << 255) - 19).to_bytes(32, "little"))]
    gpu.close()
    c.close()


@pytest.mark.parametrize("hash_id,H,n", [(0, 12, 300), (1, 6, 9), (0, 40, 50), (0, 64, 7), (0, 1, 1)])
def test_positional_padding_mode(cref, hash_id, H, n):
    """padding blindings keyed by (level, index) -- Paddable::padding(idx, secret) as a function of its
    arguments, which the reference leaves as a TODO (src/dapol/node.rs:85-88).  Every node equals the oracle's tree built in the
    same mode; the creation-order mode on the same inputs gives a different root, and pad_base no longer matters."""
    from dapol_b200 import Context, Dapol
    c = Context(0, 15)
    idx, vals, bl = _inputs(H, n, 777 + H)
    stream_root = Dapol.new_blank(c, hash_id, H, min(H, 4)).build(idx, vals, bl, PAD_SEED, 3).root_raw().com
    c.set_padding_mode(True)
    gpu = Dapol.new_blank(c, hash_id, H, min(H, 4)).build(idx, vals, bl, PAD_SEED, 3)
    _assert_same_tree(gpu, cref.Tree(hash_id, H, idx, vals, bl, PAD_SEED, positional=True), H)
    again = Dapol.new_blank(c, hash_id, H, min(H, 4)).build(idx, vals, bl, PAD_SEED, 99)
    assert again.root_raw().com == gpu.root_raw().com
    if gpu.num_padding:
        assert gpu.root_raw().com != stream_root
    c.set_padding_mode(False)
    back = Dapol.new_blank(c, hash_id, H, min(H, 4)).build(idx, vals, bl, PAD_SEED, 3)
    assert back.root_raw().com == stream_root
    for t in (gpu, again, back):
        t.close()
    c.close()


def test_identity_commitments_in_batch(ctx, cref):
    """Identity commitments (v = 0, r = 0 mod l) zero the batched inversion's input; the kernels mask them out."""
    from dapol_b200 import Dapol
    L = 2 ** 252 + 27742317777372353535851937790883648493
    H, n = 5, 7
    idx = np.array([1, 2, 3, 9, 17, 18, 30], np.uint64)
    vals = np.array([0, 0, 5, 0, 7, 0, 0], np.uint64)
    bl = np.zeros((n, 32), np.uint8)
    bl[1] = np.frombuffer(L.to_bytes(32, "little"), np.uint8)
    bl[2] = 9; bl[4] = 1
    bl[6] = np.frombuffer((2 * L).to_bytes(32, "little"), np.uint8)
    gpu = Dapol.new_blank(ctx, 0, H, H).build(idx, vals, bl, PAD_SEED)
    ora = cref.Tree(0, H, idx, vals, bl, PAD_SEED)
    _assert_same_tree(gpu, ora, H)
    lv = gpu.level(H)
    assert lv["comc"][list(lv["idx"]).index(1)].tobytes() == bytes(32)


def test_reference_kat_tree(ctx, cref):
    """src/dapol/tests.rs:17-25: leaves at 7,12,2,4 (Blake2s, H=4), root value 26."""
    from dapol_b200 import Dapol
    ib, io = cref.pack_ids([b"a", b"b", b"c", b"d"]); eb, eo = cref.pack_ids([b"w", b"x", b"y", b"z"])
    rc, idx, bl, _ = cref.derive_leaves(1, ib, io, eb, eo, b"test", 4)
    order = np.argsort(idx)
    gpu = Dapol.new_blank(ctx, 1, 4, 2).build(idx[order], np.array([3, 5, 7, 11], np.uint64)[order], bl[order], PAD_SEED)
    assert gpu.root_raw().value == 26
    lv = gpu.level(4)
    k = list(lv["idx"]).index(7)
    assert lv["comc"][k].tobytes().hex() == "bc9f755eff46224952e6c00e1ad0cef3d8e3393254af927c593ea1702e64e93a"
    assert lv["hash"][k].tobytes().hex() == "b1a1242d5ad2d09503b05d312f13b5c88440fb30a5db0b6aa62881bae94a44fb"


def test_paths_vs_oracle(ctx, cref):
    from dapol_b200 import Dapol
    H, n = 12, 300
    idx, vals, bl = _inputs(H, n, 77)
    gpu = Dapol.new_blank(ctx, 0, H, H).build(idx, vals, bl, PAD_SEED)
    ora = cref.Tree(0, H, idx, vals, bl, PAD_SEED)
    pick = idx[[0, 5, 17, 150, 299]]
    p = gpu.paths(pick)
    for q, leaf in enumerate(pick):
        o = ora.path(int(leaf))
        assert (p["v"][q] == o["v"]).all() and (p["comc"][q] == o["comc"]).all() and (p["hash"][q] == o["hash"]).all()
        nd = ora.get_node(H, int(leaf))
        assert p["leaf_comc"][q].tobytes() == nd["comc"] and p["leaf_hash"][q].tobytes() == nd["hash"]
    absent = next(x for x in range(1 << H) if x not in set(idx.tolist()))
    assert gpu.paths([absent]) is None  # reference: None for an unknown index (mod.rs:173)


def test_bad_inputs(ctx):
    from dapol_b200 import Dapol, DapolError
    idx, vals, bl = _inputs(8, 10, 3)
    with pytest.raises(DapolError) as e:
        Dapol.new_blank(ctx, 0, 8, 8).build(idx[::-1].copy(), vals, bl, PAD_SEED)  # unsorted (smtree rejects)
    assert e.value.code == 16
    with pytest.raises(DapolError) as e:
        Dapol.new_blank(ctx, 0, 65, 8).build(idx, vals, bl, PAD_SEED)
    assert e.value.code == 1
    with pytest.raises(DapolError) as e:
        Dapol.new_blank(ctx, 0, 3, 3).build(idx, vals, bl, PAD_SEED)  # index outside the tree
    assert e.value.code == 16


def test_large_tree_properties(ctx, cref):
    """2^16 leaves at H=32: full parity with the oracle at a size it finishes in seconds (8 threads),
    plus size-independent properties: root value = sum, padding count formula, sibling adjacency."""
    from dapol_b200 import Dapol
    H, n = 32, 1 << 14
    idx, vals, bl = _inputs(H, n, 5)
    gpu = Dapol.new_blank(ctx, 0, H, H).build(idx, vals, bl, PAD_SEED)
    ora = cref.Tree(0, H, idx, vals, bl, PAD_SEED)
    r, o = gpu.root_raw(), ora.root()
    assert (r.value, r.com, r.hash) == (o["v"], o["comc"], o["hash"])
    assert gpu.num_padding == ora.num_pads
    for h in (H, H - 1, 20, 8, 1):
        g, oo = gpu.level(h), ora.level(h)
        assert (g["comc"] == oo["comc"]).all() and (g["hash"] == oo["hash"]).all()
        assert ((g["idx"][0::2] ^ 1) == g["idx"][1::2]).all()


def test_full_size_tree_properties(cref):
    """BASELINE config 2 at its full size (2^20 users, height 32, from liabilities, the bench's synthetic set): properties that do
    not need the oracle to rebuild the tree -- the root value is the sum of the liabilities; node and padding counts follow from
    the leaf indexes; children of every parent are adjacent; the id -> index map is a permutation-free injection into the leaf
    level; and inclusion proofs of sampled users verify under the ORACLE's verifier against the GPU root (the Merkle fold hashes
    every level of the path, the range proofs cover every sibling commitment), and a proof against a wrong leaf does not."""
    import hashlib
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import AUDIT_SEED, synth_liabilities
    from dapol_b200 import Context, Dapol
    c = Context(0)
    c.set_rangeproof_window(8)  # 16 sampled proofs: no point in building wide tables
    n, H, agg = 1 << 20, 32, 4
    liab = synth_liabilities(n)
    t = Dapol.new(c, 0, liab, AUDIT_SEED, H, agg, PAD_SEED)
    root = t.root_raw()
    assert root.value == int(liab[4].sum())  # src/dapol/tests.rs:24 at scale
    leaf = t.level(H)
    real = leaf["is_pad"] == 0
    lidx = leaf["idx"][real]
    assert len(lidx) == n and (np.diff(lidx.astype(np.int64)) > 0).all()  # sorted, distinct: no index collision survived
    # node / padding counts from the leaf indexes alone (level h has one real node per distinct prefix)
    nodes, pads, cur = 1, 0, lidx
    for h in range(H, 0, -1):
        par = np.unique(cur >> np.uint64(1))
        nodes += 2 * len(par); pads += 2 * len(par) - len(cur)
        cur = par
    assert (t.num_nodes, t.num_padding) == (nodes, pads)
    for h in (H, 24, 12):
        g = t.level(h)
        assert ((g["idx"][0::2] ^ 1) == g["idx"][1::2]).all() and (g["v"][g["is_pad"] == 1] == 0).all()
    users = [0, 1, n // 3, n - 1] + [int(x) for x in np.random.default_rng(7).integers(0, n, 12)]
    picks = [t.leaf_index_of(u) for u in users]
    assert all(p is not None and lidx[np.searchsorted(lidx, np.uint64(p))] == p for p in picks)  # every mapped index is a real leaf
    proofs = t.generate_proofs(picks, hashlib.sha256(b"full-size").digest())
    paths = t.paths(picks)
    for k, pr in enumerate(proofs):
        lc, lh = paths["leaf_comc"][k].tobytes(), paths["leaf_hash"][k].tobytes()
        assert cref.verify_inclusion(0, 0, pr.serialize(), root.com, root.hash, lc, lh)
        other = paths["leaf_comc"][(k + 1) % len(picks)].tobytes()
        assert not cref.verify_inclusion(0, 0, pr.serialize(), root.com, root.hash, other, lh)
    t.close(); c.close()


def _liabs(n, seed, dense=False):
    rnd = random.Random(seed)
    ids = [rnd.randbytes(rnd.randrange(0, 40)) + i.to_bytes(3, "little") for i in range(n)]
    eids = [rnd.randbytes(rnd.randrange(0, 70)) for _ in range(n)]
    vals = [rnd.randrange(1 << 32) for _ in range(n)]
    return ids, eids, vals


@pytest.mark.parametrize("hash_id,H,n", [(1, 4, 4), (0, 7, 64), (0, 10, 300), (1, 9, 256), (0, 16, 5000), (0, 32, 3000), (0, 64, 50)])
def test_from_liabilities_vs_oracle(ctx, cref, hash_id, H, n):
    """Dapol::new end to end (leaf derivation incl. collisions, sort, build) == oracle derive + oracle build."""
    from dapol_b200 import Dapol
    ids, eids, vals = _liabs(n, 31 + n)
    seed = b"audit-seed"
    if H == 4:
        ids, eids, vals, seed = [b"a", b"b", b"c", b"d"], [b"w", b"x", b"y", b"z"], [3, 5, 7, 11], b"test"
    gpu = Dapol.new(ctx, hash_id, list(zip(ids, eids, vals)), seed, H, H, PAD_SEED)
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc, idx, bl, _ = cref.derive_leaves(hash_id, ib, io, eb, eo, seed, H)
    assert rc == 0
    for i in (0, 1, n // 2, n - 1):
        assert gpu.leaf_index_of(i) == int(idx[i])
    if H == 4:
        assert [gpu.leaf_index_of(i) for i in range(4)] == [7, 12, 2, 4] and gpu.root_raw().value == 26
    order = np.argsort(idx)
    ora = cref.Tree(hash_id, H, idx[order], np.array(vals, np.uint64)[order], bl[order], PAD_SEED)
    _assert_same_tree(gpu, ora, H)
    # leaf blindings are kept as Scalar::from_bits gives them (unreduced), like the reference node
    assert (gpu.level(H)["r"][gpu.level(H)["is_pad"] == 0] == bl[order]).all()


def test_from_liabilities_errors(ctx, cref):
    from dapol_b200 import Dapol, DapolError
    ids, eids, vals = _liabs(100, 5)
    ids[57] = ids[13]
    with pytest.raises(DapolError) as e:
        Dapol.new(ctx, 0, list(zip(ids, eids, vals)), b"s", 10, 10, PAD_SEED)
    assert e.value.code == 4 and e.value.detail == 57          # DuplicatedInternalId (mod.rs:341-343)
    ids, eids, vals = _liabs(100, 6)
    with pytest.raises(DapolError) as e:
        Dapol.new(ctx, 0, list(zip(ids, eids, vals)), b"s", 7, 7, PAD_SEED)
    assert e.value.code == 2                                   # SparsityTooSmall: 2^7 < 2*100 (mod.rs:110-116)
    with pytest.raises(DapolError) as e:
        Dapol.new(ctx, 0, list(zip(ids, eids, vals)), b"s", 65, 7, PAD_SEED)
    assert e.value.code == 1                                   # TreeHeightTooBig (mod.rs:104-109)
    # densest legal tree (N = 2^H / 2): many collisions, result must still match the serial rule
    ids, eids, vals = _liabs(128, 8)
    gpu = Dapol.new(ctx, 0, list(zip(ids, eids, vals)), b"s", 8, 8, PAD_SEED)
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc, idx, _, _ = cref.derive_leaves(0, ib, io, eb, eo, b"s", 8)
    assert rc == 0 and [gpu.leaf_index_of(i) for i in range(128)] == idx.tolist()
