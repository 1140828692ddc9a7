"""D = blake2::Blake2b (64-byte digests) on the GPU -- the other half of the reference's integration matrix
(src/tests.rs:100-106: TesterDapol::<blake2::Blake2b, RangeProofPadding / RangeProofSplitting>) through the
C ABI against the CPU oracle.  Shape of TesterDapol::test (src/tests.rs:26-96): height 10, 100 leaves, aggregation factors
1..10, new_blank + build, batches of 10 leaves -> generate_proof_batch -> serialize -> deserialize -> verify_batch, then every
leaf -> generate_proof -> serialize -> deserialize -> verify.  Dapol::new stays 32-byte only (mod.rs:101-103)."""
import hashlib
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200 blake2b").digest()
B2B = 2


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(8)
    yield c
    c.close()


def _trees(ctx, cref, n, H, seed, agg=1, policy=0):
    from dapol_b200 import Dapol
    rnd = random.Random(seed)
    pick = set()
    while len(pick) < n:
        pick.add(rnd.randrange(1 << H) if H else 0)
    idx = np.array(sorted(pick), np.uint64)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32)
    bl[:, 31] &= 0x7F
    gpu = Dapol.new_blank(ctx, B2B, H, agg, policy).build(idx, vals, bl, PAD_SEED, 3)
    ora = cref.Tree(B2B, H, idx, vals, bl, PAD_SEED, 3)
    return gpu, ora, idx


@pytest.mark.parametrize("n,H", [(100, 10), (1, 0), (1, 1), (2, 1), (9, 6), (300, 33), (5, 64), (4096, 20)])
def test_every_node_vs_oracle(ctx, cref, n, H):
    gpu, ora, idx = _trees(ctx, cref, n, H, 11 * n + H)
    assert gpu.num_padding == ora.num_pads
    for h in range(H + 1):
        g, o = gpu.level(h), ora.level(h)
        assert g["hash"].shape == o["hash"].shape and g["hash"].shape[1] == 64
        for key in ("idx", "v", "comc", "hash", "is_pad"):
            assert (g[key] == o[key]).all(), (h, key)
    root, oroot = gpu.root_raw(), ora.root()
    assert (root.value, root.com, root.hash) == (oroot["v"], oroot["comc"], oroot["hash"]) and len(root.hash) == 64
    lf = gpu.level(H)
    k = int(np.flatnonzero(lf["is_pad"] == 0)[0])
    assert lf["hash"][k].tobytes() == hashlib.blake2b(lf["comc"][k].tobytes()).digest()  # DapolNode::new: D(compress(com)), node.rs:33-36
    gpu.close()


@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("agg", [1, 4, 7, 10])
def test_tester_dapol_blake2b(ctx, cref, agg, policy):
    """TesterDapol::<Blake2b, R>::test (src/tests.rs:26-96) for one aggregation factor: batch proofs of 10 leaves and single
    proofs, byte-identical with the oracle's prover, verified on the GPU and by the oracle."""
    from dapol_b200 import DapolProof, DapolProofNode
    H, n = 10, 100
    gpu, ora, idx = _trees(ctx, cref, n, H, 1000 + agg, agg, policy)
    root, oroot = gpu.root(), ora.root()
    assert (root.com, root.hash) == (oroot["comc"], oroot["hash"])
    all_idx = [int(x) for x in idx]
    for b in (0, 4, 9):  # batches of 10 consecutive leaves (src/tests.rs:56-76)
        picks = all_idx[10 * b:10 * b + 10]
        nodes = [ora.get_node(H, x) for x in picks]
        leaves = [DapolProofNode(nd["comc"], nd["hash"]) for nd in nodes]
        proof = gpu.generate_proof_batch(picks, PROVE_SEED)
        want = cref.prove_inclusion_batch(ora, picks, agg, policy, PROVE_SEED)
        assert want is not None and proof.serialize() == want
        back = DapolProof.deserialize(proof.serialize(), B2B, policy)
        assert back.verify_batch(ctx, root, leaves)
        assert cref.verify_inclusion_batch(B2B, policy, want, oroot["comc"], oroot["hash"], [l.com for l in leaves], [l.hash for l in leaves])
        assert not back.verify_batch(ctx, root, leaves[::-1])
        for at in (40, len(want) - 5, len(want) - 40):
            bad = bytearray(want); bad[at] ^= 1
            assert not DapolProof(bytes(bad), B2B, policy).verify_batch(ctx, root, leaves)
    picks = all_idx[::9]
    proofs = gpu.generate_proofs(picks, PROVE_SEED)
    paths = gpu.paths(picks)
    assert paths["hash"].shape == (len(picks), H, 64) and paths["leaf_hash"].shape == (len(picks), 64)
    leaves = []
    for q, (x, pf) in enumerate(zip(picks, proofs)):
        want = ora.prove_inclusion(x, agg, policy, PROVE_SEED)
        assert pf.serialize() == want, (x, len(pf.serialize()), len(want))
        nd = ora.get_node(H, x)
        assert (paths["leaf_comc"][q].tobytes(), paths["leaf_hash"][q].tobytes()) == (nd["comc"], nd["hash"])
        op = ora.path(x)
        assert (paths["hash"][q] == op["hash"]).all() and (paths["comc"][q] == op["comc"]).all()
        assert cref.verify_inclusion(B2B, policy, want, oroot["comc"], oroot["hash"], nd["comc"], nd["hash"])
        leaves.append(DapolProofNode(nd["comc"], nd["hash"]))
    assert DapolProof.verify_many(ctx, root, leaves, proofs).all()
    # reject parity: a wrong leaf, tampered bytes in the range part / in the lower and upper halves of a sibling hash, truncation
    bad_proofs, bad_leaves = [], []
    data = proofs[1].serialize()
    for at in (40, len(data) - 5, len(data) - 40, len(data) - 70, len(data) - 100):
        bb = bytearray(data); bb[at] ^= 1
        bad_proofs.append(DapolProof(bytes(bb), B2B, policy)); bad_leaves.append(leaves[1])
    bad_proofs.append(DapolProof(data[:-1], B2B, policy)); bad_leaves.append(leaves[1])
    bad_proofs.append(proofs[1]); bad_leaves.append(leaves[2])
    bad_proofs.append(proofs[2]); bad_leaves.append(leaves[2])  # a good one in the same batch
    got = DapolProof.verify_many(ctx, root, bad_leaves, bad_proofs).tolist()
    want_v = [bool(cref.verify_inclusion(B2B, policy, p.serialize(), oroot["comc"], oroot["hash"], l.com, l.hash)) for p, l in zip(bad_proofs, bad_leaves)]
    assert got == want_v == [False] * 7 + [True]
    gpu.close()


def test_blake2b_boundary(ctx, cref, tmp_path):
    """Dapol::new insists on 32-byte digests (mod.rs:101-103): InvalidDigestSize; a Blake2b tree saved and loaded gives the same
    levels and proofs; sizes follow the digest length."""
    from dapol_b200 import Dapol, DapolError, _ffi
    L = _ffi.lib()
    assert (L.dapol_digest_len(0), L.dapol_digest_len(1), L.dapol_digest_len(2), L.dapol_digest_len(9)) == (32, 32, 64, 0)
    assert L.dapol_inclusion_proof_size_d(10, 3, 0, 2) == L.dapol_inclusion_proof_size(10, 3, 0) + 32 * 10
    with pytest.raises(DapolError) as e:
        Dapol.new(ctx, B2B, [(b"a", b"b", 1)], b"seed", 8, 1, PAD_SEED)
    assert e.value.code == 3
    gpu, ora, idx = _trees(ctx, cref, 50, 9, 77, 2, 1)
    path = os.path.join(tmp_path, "b2b.tree")
    gpu.save(path)
    back = Dapol.load(ctx, path, 2, 1)
    assert back.hash_id == B2B
    for h in (0, 4, 9):
        a, b = gpu.level(h), back.level(h)
        assert all((a[k] == b[k]).all() for k in a)
    x = int(idx[7])
    assert back.generate_proof(x, PROVE_SEED).serialize() == gpu.generate_proof(x, PROVE_SEED).serialize() == ora.prove_inclusion(x, 2, 1, PROVE_SEED)
    back.close(); gpu.close()
