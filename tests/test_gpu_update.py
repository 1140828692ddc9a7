"""Dapol::update (src/dapol/mod.rs:210-213) on the GPU through the C ABI.  Shape of the reference's own test
(src/tests.rs:36-47,78-95): a tree built in one go and a tree grown leaf by leaf with `update` have equal roots (DapolNode
equality is equality of the values, src/dapol/node.rs:115-120) and both give proofs that verify.  Stronger here: after every
batch of updates the tree equals, node for node, the oracle's build over the merged leaves -- and with position-keyed padding
(the deterministic mode) update and build agree bit for bit."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200 update").digest()


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(8)
    yield c
    c.close()


def _items(n, H, seed):
    rnd = random.Random(seed)
    idx = np.array(sorted(rnd.sample(range(1 << H), n)), np.uint64)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32)
    bl[:, 31] &= 0x7F
    return idx, vals, bl


def _same_levels(gpu, ora, H):
    for h in range(H + 1):
        g, o = gpu.level(h), ora.level(h)
        for key in ("idx", "v", "comc", "hash", "is_pad"):
            assert g[key].shape == o[key].shape and (g[key] == o[key]).all(), (h, key)


@pytest.mark.parametrize("hash_id,policy", [(0, 1), (0, 0), (2, 0)])
def test_update_leaf_by_leaf_equals_build(ctx, cref, hash_id, policy):
    """src/tests.rs:36-47: update_dapol.update(&item.0, item.1, &secret) for every item; roots equal; proofs of both verify."""
    from dapol_b200 import Dapol, DapolProof, DapolProofNode
    H, n, agg = 10, 100, 3
    idx, vals, bl = _items(n, H, 5 + hash_id)
    build_dapol = Dapol.new_blank(ctx, hash_id, H, agg, policy).build(idx, vals, bl, PAD_SEED)
    update_dapol = Dapol.new_blank(ctx, hash_id, H, agg, policy)
    order = list(range(n))
    random.Random(9).shuffle(order)  # any insertion order
    for step, i in enumerate(order):
        update_dapol.update(idx[i], vals[i], bl[i], PAD_SEED, 0)
        if step in (0, 1, 17, n - 1):
            have = sorted(order[:step + 1])
            ora = cref.Tree(hash_id, H, idx[have], vals[have], bl[have], PAD_SEED, 0)
            _same_levels(update_dapol, ora, H)
    a, b = build_dapol.root_raw(), update_dapol.root_raw()
    assert a.value == b.value == int(vals.sum())                       # assert_eq!(build_dapol.root_raw(), update_dapol.root_raw())
    assert (a.com, a.hash, a.blinding) == (b.com, b.hash, b.blinding)  # same pad stream and base: the very same tree
    picks = [int(x) for x in idx[::13]]
    for tree in (build_dapol, update_dapol):
        proofs = tree.generate_proofs(picks, PROVE_SEED)
        paths = tree.paths(picks)
        leaves = [DapolProofNode(paths["leaf_comc"][q].tobytes(), paths["leaf_hash"][q].tobytes()) for q in range(len(picks))]
        assert DapolProof.verify_many(ctx, tree.root(), leaves, proofs).all()
    build_dapol.close(); update_dapol.close()


@pytest.mark.parametrize("positional", [False, True])
def test_update_batches_replace_and_insert(ctx, cref, positional):
    from dapol_b200 import Dapol
    H = 14
    idx, vals, bl = _items(600, H, 21)
    ctx.set_padding_mode(positional)
    try:
        tree = Dapol.new_blank(ctx, 0, H, 1).build(idx[:400], vals[:400], bl[:400], PAD_SEED, 7)
        cur = {int(i): (int(v), b.tobytes()) for i, v, b in zip(idx[:400], vals[:400], bl[:400])}
        rnd = random.Random(3)
        for rnd_no, (lo, hi, nrep) in enumerate([(400, 401, 0), (401, 500, 25), (500, 600, 100), (0, 0, 400)]):
            new = {int(idx[j]): (int(vals[j]), bl[j].tobytes()) for j in range(lo, hi)}
            for x in rnd.sample(sorted(cur), nrep):                     # replaced leaves: new value and blinding at an existing index
                new[x] = (rnd.randrange(1 << 32), rnd.randbytes(31) + b"\x01")
            ks = sorted(new)
            nb = np.frombuffer(b"".join(new[x][1] for x in ks), np.uint8).reshape(-1, 32)
            pad_base = 1000 * (rnd_no + 1)
            tree.update(np.array(ks, np.uint64), np.array([new[x][0] for x in ks], np.uint64), nb, PAD_SEED, pad_base)
            cur.update(new)
            keys = sorted(cur)
            ora = cref.Tree(0, H, np.array(keys, np.uint64), np.array([cur[x][0] for x in keys], np.uint64),
                            np.frombuffer(b"".join(cur[x][1] for x in keys), np.uint8).reshape(-1, 32), PAD_SEED, pad_base, positional=positional)
            _same_levels(tree, ora, H)
            assert tree.root_raw().value == sum(v for v, _ in cur.values())
        tree.close()
    finally:
        ctx.set_padding_mode(False)


def test_update_errors(ctx):
    from dapol_b200 import Dapol, DapolError
    idx, vals, bl = _items(20, 8, 4)
    tree = Dapol.new_blank(ctx, 0, 8, 1).build(idx, vals, bl, PAD_SEED)
    for bad_idx in ([5, 5], [9, 3], [1 << 8]):                         # duplicate, unsorted, outside the tree
        with pytest.raises(DapolError) as e:
            tree.update(np.array(bad_idx, np.uint64), vals[:len(bad_idx)], bl[:len(bad_idx)], PAD_SEED)
        assert e.value.code == 16
    assert tree.root_raw().value == int(vals.sum())                    # a failed update leaves the tree as it was
    tree.close()
