"""Batch proofs on the GPU: ONE DapolProof for several leaves -- Dapol::generate_proof_batch
(src/dapol/mod.rs:172-190) and DapolProof::verify_batch (src/proof/mod.rs:49-54) through the C ABI, against the CPU oracle.
Shapes follow the reference's own test (src/proof/tests.rs:6-35: height 8, 20 leaves, batch of 10, blake3 + Splitting,
aggregation 1: serialize -> deserialize -> verify_batch) and src/tests.rs:56-76 (batches of 10 out of 100 leaves, height 10,
aggregation factors 1..10, both policies)."""
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


def _trees(ctx, cref, n, H, hash_id, seed):
    from dapol_b200 import Dapol
    rnd = random.Random(seed)
    idx = np.array(sorted(rnd.sample(range(1 << H), n)), np.uint64)
    vals = np.array([rnd.randrange(1 << 32) for _ in range(n)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32)
    bl[:, 31] &= 0x0F
    gpu = Dapol.new_blank(ctx, hash_id, H, 1).build(idx, vals, bl, PAD_SEED)
    ora = cref.Tree(hash_id, H, idx, vals, bl, PAD_SEED)
    return gpu, ora, idx


@pytest.mark.parametrize("n,H,batch,agg,policy,hash_id", [(20, 8, 10, 1, 1, 0), (100, 10, 10, 1, 0, 0), (100, 10, 10, 7, 1, 0), (100, 10, 10, 10, 0, 1),
                                                          (100, 10, 10, 4, 0, 0), (9, 5, 9, 0, 1, 0), (64, 7, 64, 2, 0, 0), (300, 12, 2, 12, 1, 0)])
def test_batch_proof_bytes_and_verify_batch(ctx, cref, n, H, batch, agg, policy, hash_id):
    from dapol_b200 import DapolProof, DapolProofNode
    gpu, ora, idx = _trees(ctx, cref, n, H, hash_id, n * 131 + H + batch)
    gpu.aggregation_factor, gpu.policy = agg, policy
    rnd = random.Random(batch)
    picks = sorted(rnd.sample([int(x) for x in idx], batch))
    proof = gpu.generate_proof_batch(picks, PROVE_SEED)
    want = cref.prove_inclusion_batch(ora, picks, agg, policy, PROVE_SEED)
    assert want is not None and proof.serialize() == want, "GPU batch proof bytes != oracle"
    root = ora.root()
    nodes = [ora.get_node(H, x) for x in picks]
    leaves = [DapolProofNode(nd["comc"], nd["hash"]) for nd in nodes]
    # serialize -> deserialize -> verify_batch (src/proof/tests.rs:28-34), on the GPU and under the oracle's verifier
    back = DapolProof.deserialize(proof.serialize(), hash_id, policy)
    assert back.verify_batch(ctx, gpu.root(), leaves)
    assert cref.verify_inclusion_batch(hash_id, policy, proof.serialize(), root["comc"], root["hash"], [l.com for l in leaves], [l.hash for l in leaves])
    # rejects: a leaf swapped for another, a missing leaf, a flipped byte in the range part / in a sibling / in an index, truncation
    if batch > 1:
        swapped = [leaves[1], leaves[0]] + leaves[2:]
        assert not back.verify_batch(ctx, gpu.root(), swapped)
        assert not back.verify_batch(ctx, gpu.root(), leaves[:-1])
    raw = proof.serialize()
    for at in (40, len(raw) - 5, len(raw) - 70):
        bad = bytearray(raw); bad[at] ^= 1
        got = DapolProof(bytes(bad), hash_id, policy).verify_batch(ctx, gpu.root(), leaves)
        assert got == cref.verify_inclusion_batch(hash_id, policy, bytes(bad), root["comc"], root["hash"], [l.com for l in leaves], [l.hash for l in leaves])
        assert not got
    assert not DapolProof(raw[:-1], hash_id, policy).verify_batch(ctx, gpu.root(), leaves)
    gpu.close()


def test_batch_of_one_is_the_single_proof(ctx, cref):
    """generate_proof(idx) = generate_proof_batch(&[idx]) (mod.rs:167-169): identical bytes."""
    gpu, ora, idx = _trees(ctx, cref, 30, 9, 0, 5)
    gpu.aggregation_factor, gpu.policy = 3, 0
    x = int(idx[11])
    assert gpu.generate_proof_batch([x], PROVE_SEED).serialize() == gpu.generate_proof(x, PROVE_SEED).serialize() == ora.prove_inclusion(x, 3, 0, PROVE_SEED)
    gpu.close()


def test_batch_errors(ctx, cref):
    from dapol_b200 import DapolError
    gpu, ora, idx = _trees(ctx, cref, 30, 9, 0, 6)
    real = [int(x) for x in idx]
    absent = next(x for x in range(1 << 9) if x not in set(real))
    gpu.aggregation_factor, gpu.policy = 1, 1
    assert gpu.generate_proof_batch(sorted([real[0], real[3], absent]), PROVE_SEED) is None          # reference: None
    with pytest.raises(DapolError):
        gpu.generate_proof_batch([real[3], real[0]], PROVE_SEED)                                      # not increasing
    gpu.aggregation_factor = 200
    with pytest.raises(DapolError):
        gpu.generate_proof_batch(real[:4], PROVE_SEED)                                                # reference: slice out of bounds
    gpu.close()
