"""Lookup by internal id on the GPU path: Dapol::generate_proof_for_id / generate_proof_batch_for_ids (src/dapol/mod.rs:148-165),
the reference's own KAT (src/dapol/tests.rs:30-108: ids a, b, c, d map to leaves 7, 12, 2, 4; a batch by ids equals the batch
by indexes), and ids longer than one BLAKE3 chunk (the reference hashes ids of any length, mod.rs:347-353)."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(8)
    yield c
    c.close()


def test_reference_kat_by_id(ctx, cref):
    """src/dapol/tests.rs:30-108: Blake2s, audit seed "test", height 4."""
    from dapol_b200 import Dapol, DapolProofNode
    liab = [(b"a", b"w", 3), (b"b", b"x", 5), (b"c", b"y", 7), (b"d", b"z", 11)]
    t = Dapol.new(ctx, 1, liab, b"test", 4, 1, PAD_SEED, policy=1)
    assert t.index_of_ids([b"a", b"b", b"c", b"d"]) == [7, 12, 2, 4]          # the reference's KAT
    assert t.root_raw().value == 26                                              # src/dapol/tests.rs:24
    assert t.index_of_ids([b"a", b"e"]) is None and t.generate_proof_for_id(b"zz", PROVE_SEED) is None   # reference: None
    p = t.generate_proof_for_id(b"c", PROVE_SEED)
    assert p.serialize() == t.generate_proof(2, PROVE_SEED).serialize()
    # generate_proof_batch_for_ids == generate_proof_batch on the mapped indexes (tests.rs:87-108); ids given in index order
    by_ids = t.generate_proof_batch_for_ids([b"c", b"d", b"a"], PROVE_SEED)
    by_idx = t.generate_proof_batch([2, 4, 7], PROVE_SEED)
    assert by_ids.serialize() == by_idx.serialize()
    paths = t.paths([2, 4, 7])
    leaves = [DapolProofNode(paths["leaf_comc"][k].tobytes(), paths["leaf_hash"][k].tobytes()) for k in range(3)]
    assert by_ids.verify_batch(ctx, t.root(), leaves)
    t.close()


@pytest.mark.parametrize("hash_id", [0, 1])
def test_long_ids_and_lookup_against_oracle(ctx, cref, hash_id):
    """ids of 0 .. 5000 bytes (BLAKE3 tree mode beyond 1024): leaf indexes, root and lookups equal the oracle's."""
    from dapol_b200 import Dapol
    rnd = random.Random(77 + hash_id)
    lens = [0, 1, 8, 63, 64, 65, 1000, 1023, 1024, 1025, 2048, 2049, 3000, 4096, 5000] + [rnd.randrange(1, 200) for _ in range(150)]
    ids = [rnd.randbytes(n) for n in lens]
    eids = [rnd.randbytes(rnd.choice([0, 5, 40, 1500])) for _ in ids]
    vals = [rnd.randrange(1 << 32) for _ in ids]
    H, seed = 12, b"long-id-seed"
    t = Dapol.new(ctx, hash_id, list(zip(ids, eids, vals)), seed, H, 2, PAD_SEED)
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc, idx, bl, _ = cref.derive_leaves(hash_id, ib, io, eb, eo, seed, H)
    assert rc == 0
    order = np.argsort(idx)
    ora = cref.Tree(hash_id, H, idx[order], np.array(vals, np.uint64)[order], bl[order], PAD_SEED)
    r, o = t.root_raw(), ora.root()
    assert (r.value, r.com, r.hash) == (o["v"], o["comc"], o["hash"])
    assert t.index_of_ids(ids) == [int(x) for x in idx]
    pick = [3, 9, 14, 40]
    assert t.index_of_ids([ids[i] for i in pick]) == [int(idx[i]) for i in pick]
    assert t.index_of_ids([ids[9] + b"x"]) is None
    assert t.generate_proof_for_id(ids[14], PROVE_SEED).serialize() == ora.prove_inclusion(int(idx[14]), 2, 0, PROVE_SEED)
    t.close()


def test_tree_from_nodes_has_no_id_map(ctx):
    from dapol_b200 import Dapol
    t = Dapol.new_blank(ctx, 0, 6, 1).build([1, 5, 9], [1, 2, 3], np.zeros((3, 32), np.uint8), PAD_SEED)
    assert t.index_of_ids([b"a"]) is None      # new_blank: empty id_to_idx_map (mod.rs:196-204)
    t.close()


@pytest.mark.parametrize("hash_id", [0, 1])
def test_id_salt_leaf_hash_mode(cref, hash_id):
    """Opt-in leaf hash of the DAPOL+ paper: every level of the GPU tree equals the oracle's tree built
    with the same leaf hashes; inclusion proofs verify (the leaf's proof node carries the id / salt hash); the default mode
    on the same liabilities gives a different root hash but the same commitments."""
    from dapol_b200 import Context, Dapol, DapolProof, DapolProofNode
    rnd = random.Random(5 + hash_id)
    n, H, seed = 120, 10, b"id-salt"
    ids = [b"user-%d" % i for i in range(n)]
    eids = [rnd.randbytes(rnd.choice([0, 4, 33, 1200])) for _ in range(n)]
    vals = [rnd.randrange(1 << 32) for _ in range(n)]
    c = Context(0, 15)
    c.set_rangeproof_window(8)
    plain = Dapol.new(c, hash_id, list(zip(ids, eids, vals)), seed, H, 2, PAD_SEED)
    c.set_leaf_hash_mode(True)
    t = Dapol.new(c, hash_id, list(zip(ids, eids, vals)), seed, H, 2, PAD_SEED)
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc, idx, bl, _ = cref.derive_leaves(hash_id, ib, io, eb, eo, seed, H)
    assert rc == 0
    lh = cref.leaf_id_hashes(hash_id, ib, io, eb, eo, seed)
    order = np.argsort(idx)
    ora = cref.tree_with_leaf_hashes(hash_id, H, idx[order], np.array(vals, np.uint64)[order], bl[order], lh[order], PAD_SEED)
    for h in range(H + 1):
        g, o = t.level(h), ora.level(h)
        for key in ("idx", "v", "comc", "hash", "is_pad"):
            assert (g[key] == o[key]).all(), (h, key)
    assert t.root_raw().com == plain.root_raw().com and t.root_raw().hash != plain.root_raw().hash
    picks = [int(idx[i]) for i in (0, 17, n - 1)]
    proofs = t.generate_proofs(picks, PROVE_SEED)
    paths = t.paths(picks)
    leaves = [DapolProofNode(paths["leaf_comc"][k].tobytes(), paths["leaf_hash"][k].tobytes()) for k in range(3)]
    assert leaves[1].hash == lh[17].tobytes()
    assert DapolProof.verify_many(c, t.root(), leaves, proofs).all()
    r = ora.root()
    for k in range(3):
        assert cref.verify_inclusion(hash_id, 0, proofs[k].serialize(), r["comc"], r["hash"], leaves[k].com, leaves[k].hash)
    t.close(); plain.close(); c.close()
