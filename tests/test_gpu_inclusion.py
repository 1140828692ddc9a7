"""GPU parity tests of inclusion proofs (Dapol::generate_proof + DapolProof::{serialize, verify}) through the C ABI against
the CPU oracle: byte-identical serialised proofs for both aggregation policies, verification both ways, reject parity.
Shapes follow the reference's tests: src/tests.rs:16-127 (H = 10 / 5, agg = 1..10), src/proof/tests.rs:6-35 (H = 8, agg = 1),
benches/dapol.rs:149-158 (agg = H)."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200 prover").digest()


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(8)
    yield c
    c.close()


def _tree(ctx, cref, hash_id, H, n, agg, policy, seed):
    from dapol_b200 import Dapol
    rnd = random.Random(seed)
    idx = np.array(sorted(rnd.sample(range(1 << H), n)), np.uint64)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32); bl[:, 31] &= 0x7F
    gpu = Dapol.new_blank(ctx, hash_id, H, agg, policy).build(idx, vals, bl, PAD_SEED)
    ora = cref.Tree(hash_id, H, idx, vals, bl, PAD_SEED)
    return gpu, ora, idx


@pytest.mark.parametrize("hash_id,H,n,agg,policy", [
    (0, 5, 10, 1, 1), (0, 5, 10, 1, 0), (0, 10, 100, 10, 0), (0, 10, 100, 10, 1), (0, 10, 100, 7, 1), (0, 10, 100, 3, 0),
    (1, 8, 20, 1, 1), (0, 16, 64, 16, 0), (0, 16, 64, 16, 1), (0, 6, 8, 6, 1), (0, 6, 8, 5, 0), (0, 12, 30, 0, 0), (0, 12, 30, 0, 1)])
def test_inclusion_proofs_vs_oracle(ctx, cref, hash_id, H, n, agg, policy):
    from dapol_b200 import DapolProof
    gpu, ora, idx = _tree(ctx, cref, hash_id, H, n, agg, policy, 7 * H + n + agg)
    pick = [int(x) for x in idx[[0, 1, n // 2, n - 1]]]
    proofs = gpu.generate_proofs(pick, PROVE_SEED)
    root = gpu.root()
    leaves = []
    for leaf, pf in zip(pick, proofs):
        want = ora.prove_inclusion(leaf, agg, policy, PROVE_SEED)
        assert pf.serialize() == want, (leaf, len(pf.serialize()), len(want))
        nd = ora.get_node(H, leaf)
        assert cref.verify_inclusion(hash_id, policy, pf.serialize(), root.com, root.hash, nd["comc"], nd["hash"])
        from dapol_b200 import DapolProofNode
        leaves.append(DapolProofNode(nd["comc"], nd["hash"]))
    assert DapolProof.verify_many(ctx, root, leaves, proofs).all()
    # serialize -> deserialize -> verify (src/proof/tests.rs:6-35)
    again = DapolProof.deserialize(proofs[0].serialize(), hash_id, policy)
    assert again.verify(ctx, root, leaves[0])
    # rejects: wrong leaf, wrong root, tampered bytes everywhere, truncation -- same verdicts as the oracle
    bad, bad_leaves = [], []
    data = proofs[1].serialize()
    rnd = random.Random(1)
    for pos in sorted(rnd.sample(range(len(data)), 12)) + [0, len(data) - 1]:
        b = bytearray(data); b[pos] ^= 0x10
        bad.append(DapolProof(bytes(b), hash_id, policy)); bad_leaves.append(leaves[1])
    bad.append(DapolProof(data[:-1], hash_id, policy)); bad_leaves.append(leaves[1])
    bad.append(DapolProof(data + b"\0", hash_id, policy)); bad_leaves.append(leaves[1])   # trailing bytes are ignored by the decoder
    bad.append(proofs[1]); bad_leaves.append(leaves[2])
    got = DapolProof.verify_many(ctx, root, bad_leaves, bad)
    want = [cref.verify_inclusion(hash_id, policy, p.serialize(), root.com, root.hash, l.com, l.hash) for p, l in zip(bad, bad_leaves)]
    assert got.tolist() == want
    from dapol_b200 import DapolProofNode
    assert not proofs[0].verify(ctx, DapolProofNode(root.hash, root.com), leaves[0])
    gpu.close()


def test_unknown_leaf_and_bad_aggregation(ctx, cref):
    from dapol_b200 import DapolError
    gpu, ora, idx = _tree(ctx, cref, 0, 8, 20, 4, 0, 3)
    absent = next(x for x in range(256) if x not in set(idx.tolist()))
    assert gpu.generate_proof(absent, PROVE_SEED) is None          # reference: None (mod.rs:173)
    gpu.aggregation_factor = 9                                      # > height: the reference panics on the slice
    with pytest.raises(DapolError):
        gpu.generate_proof(int(idx[0]), PROVE_SEED)
    gpu.close()


def test_all_leaves_prove_and_verify(ctx, cref):
    """src/tests.rs:108-127 prove_n_verify shape at a larger size: every leaf's proof verifies; the batch is one GPU pass."""
    from dapol_b200 import DapolProof, DapolProofNode
    H, n = 12, 256
    gpu, ora, idx = _tree(ctx, cref, 0, H, n, H, 1, 11)
    proofs = gpu.generate_proofs(idx, PROVE_SEED)
    lv = gpu.level(H)
    real = lv["is_pad"] == 0
    leaves = [DapolProofNode(c.tobytes(), h.tobytes()) for c, h in zip(lv["comc"][real], lv["hash"][real])]
    assert (lv["idx"][real] == idx).all()
    assert DapolProof.verify_many(ctx, gpu.root(), leaves, proofs).all()
    assert proofs[17].serialize() == ora.prove_inclusion(int(idx[17]), H, 1, PROVE_SEED)
    gpu.close()
