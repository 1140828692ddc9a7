"""GPU parity tests of the batched Bulletproofs prover / verifier (through the C ABI) against the CPU oracle:
byte-identical proofs under the seeded-RNG contract, accept / reject parity, cross-verification both ways."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = hashlib.sha256(b"dapol-b200").digest()
L = 2 ** 252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module", params=[8, 12])
def ctx(request):
    from dapol_b200 import Context
    c = Context(0)
    c.set_rangeproof_window(request.param)
    yield c
    c.close()


def _batch(rnd, nbits, m, k):
    vals = np.array([[rnd.randrange(1 << nbits) for _ in range(m)] for _ in range(k)], np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * m * k), np.uint8).copy().reshape(k, m, 32)
    bl[:, :, 31] &= 0x7F  # Scalar::from_bits domain: possibly unreduced
    streams = np.array([rnd.randrange(1 << 64) for _ in range(k)], np.uint64)
    bases = np.array([rnd.randrange(1 << 40) << 8 for _ in range(k)], np.uint64)
    return vals, bl, streams, bases


def _coms(cref, vals, bl):
    return np.array([[np.frombuffer(cref.commit(int(v), (int.from_bytes(b.tobytes(), "little") % L).to_bytes(32, "little")), np.uint8)
                      for v, b in zip(vr, br)] for vr, br in zip(vals, bl)])


@pytest.mark.parametrize("nbits,m,k", [(64, 1, 37), (64, 2, 5), (64, 16, 3), (32, 1, 9), (8, 4, 4), (64, 32, 2), (16, 1, 3), (64, 64, 1)])
def test_prove_bytes_and_verify(ctx, cref, nbits, m, k):
    rnd = random.Random(nbits + m * 1000 + k)
    vals, bl, streams, bases = _batch(rnd, nbits, m, k)
    vals[0, 0] = (1 << nbits) - 1
    vals[-1, -1] = 0
    proofs = ctx.rangeproof_prove_batch(nbits, vals, bl, SEED, streams, bases)
    assert proofs.shape[1] == 32 * (9 + 2 * ((nbits * m).bit_length() - 1))
    if (nbits, m) == (64, 1):
        assert proofs.shape[1] == 672  # SINGLE_PROOF_BYTE_NUM, src/range/mod.rs:18
    coms = _coms(cref, vals, bl)
    check = range(k) if nbits * m <= 1024 else range(min(k, 1))  # the oracle prover is slow for big aggregates
    for i in check:
        want = cref.rp_prove([int(x) for x in vals[i]], [b.tobytes() for b in bl[i]], SEED, int(streams[i]), int(bases[i]), nbits)
        assert proofs[i].tobytes() == want, i
    for i in range(k):  # GPU proofs verify under the oracle verifier
        assert cref.rp_verify(proofs[i].tobytes(), [c.tobytes() for c in coms[i]], nbits)
    assert ctx.rangeproof_verify_batch(nbits, m, proofs, coms).all()


def test_verify_reject_parity(ctx, cref):
    """1 % deliberately invalid and every malformed-input class: per-proof results equal the oracle's."""
    rnd = random.Random(99)
    nbits, m, k = 64, 1, 200
    vals, bl, streams, bases = _batch(rnd, nbits, m, k)
    proofs = ctx.rangeproof_prove_batch(nbits, vals, bl, SEED, streams, bases)
    coms = _coms(cref, vals, bl)
    bad = set(rnd.sample(range(k), 24))
    for n_, i in enumerate(sorted(bad)):
        kind = n_ % 6
        if kind == 0:
            proofs[i, rnd.randrange(672)] ^= 1 << rnd.randrange(8)           # flipped bit
        elif kind == 1:
            coms[i, 0] = coms[(i + 1) % k, 0]                                   # someone else's commitment
        elif kind == 2:
            proofs[i, 128:160] = np.frombuffer((L + 3).to_bytes(32, "little"), np.uint8)   # non-canonical t_x
        elif kind == 3:
            proofs[i, 32:64] = 0                                                # identity S
        elif kind == 4:
            proofs[i, 224:256] = np.frombuffer((2 ** 255 - 19 + 2).to_bytes(32, "little"), np.uint8)  # non-canonical point
        else:
            coms[i, 0] = np.frombuffer(cref.commit(int(vals[i, 0]) ^ 1, (int.from_bytes(bl[i, 0].tobytes(), "little") % L).to_bytes(32, "little")), np.uint8)
    got = ctx.rangeproof_verify_batch(nbits, m, proofs, coms)
    want = np.array([cref.rp_verify(proofs[i].tobytes(), [coms[i, 0].tobytes()], nbits) for i in range(k)])
    assert (got == want).all()
    assert not got[sorted(bad)].any() and got.sum() == k - len(bad)
    # a proof of the wrong length for (n, m) is a reject, not an error
    assert not ctx.rangeproof_verify_batch(nbits, m, proofs[:, :640].copy(), coms).any()


def test_oracle_proofs_verify_on_gpu(ctx, cref):
    rnd = random.Random(4)
    for nbits, m in [(64, 1), (64, 4), (32, 2)]:
        k = 3
        vals, bl, streams, bases = _batch(rnd, nbits, m, k)
        proofs = np.array([np.frombuffer(cref.rp_prove([int(x) for x in vals[i]], [b.tobytes() for b in bl[i]], SEED, int(streams[i]), int(bases[i]), nbits), np.uint8)
                           for i in range(k)])
        assert ctx.rangeproof_verify_batch(nbits, m, proofs, _coms(cref, vals, bl)).all()


def test_out_of_range_value(ctx, cref):
    bl = np.zeros((2, 1, 32), np.uint8); bl[:, 0, 0] = 7
    vals = np.array([[255], [300]], np.uint64)
    proofs = ctx.rangeproof_prove_batch(8, vals, bl, SEED, [0, 1], [0, 0])
    coms = _coms(cref, vals, bl)
    assert ctx.rangeproof_verify_batch(8, 1, proofs, coms).tolist() == [True, False]


def test_bad_shapes(ctx):
    from dapol_b200 import DapolError
    with pytest.raises(DapolError):
        ctx.rangeproof_prove_batch(64, np.zeros((1, 3), np.uint64), np.zeros((1, 3, 32), np.uint8), SEED, [0], [0])  # m not a power of two
    with pytest.raises(DapolError):
        ctx.rangeproof_prove_batch(24, np.zeros((1, 1), np.uint64), np.zeros((1, 1, 32), np.uint8), SEED, [0], [0])  # bulletproofs: n in {8,16,32,64}


@pytest.mark.parametrize("nbits,m,k,G,cbits", [(64, 1, 300, 64, 0), (64, 1, 257, 1000, 0), (64, 1, 64, 7, 4), (32, 2, 40, 16, 0), (64, 16, 12, 4, 6), (8, 1, 33, 33, 9)])
def test_batched_verifier_same_verdicts(ctx, cref, nbits, m, k, G, cbits):
    """dapol_ctx_set_verify_mode(G): groups of G proofs checked by one random linear combination with the bucket method (Pippenger).
    All-valid batches pass without a single per-proof re-verification; with bad proofs (every malformed-input class) the verdicts
    are those of the per-proof verifier and of the oracle, and only the groups that hold a bad proof are re-verified."""
    rnd = random.Random(nbits + 31 * m + k)
    vals, bl, streams, bases = _batch(rnd, nbits, m, k)
    proofs = ctx.rangeproof_prove_batch(nbits, vals, bl, SEED, streams, bases)
    coms = _coms(cref, vals, bl)
    try:
        ctx.set_verify_mode(G, cbits)
        f0 = ctx.verify_fallbacks
        assert ctx.rangeproof_verify_batch(nbits, m, proofs, coms).all()
        assert ctx.verify_fallbacks == f0
        ctx.set_verify_mode(G, cbits, bytes(range(32)))  # fixed weights: the same verdicts
        assert ctx.rangeproof_verify_batch(nbits, m, proofs, coms).all()
        bad = sorted(rnd.sample(range(k), max(2, k // 30)))
        plen = proofs.shape[1]
        for n_, i in enumerate(bad):
            kind = n_ % 5
            if kind == 0:
                proofs[i, rnd.randrange(plen)] ^= 1 << rnd.randrange(8)
            elif kind == 1:
                coms[i, 0] = coms[(i + 1) % k, 0]
            elif kind == 2:
                proofs[i, 128:160] = np.frombuffer((L + 3).to_bytes(32, "little"), np.uint8)
            elif kind == 3:
                proofs[i, 32:64] = 0
            else:
                proofs[i, 224:256] = np.frombuffer((2 ** 255 - 19 + 2).to_bytes(32, "little"), np.uint8)
        ctx.set_verify_mode(G, cbits)
        f0 = ctx.verify_fallbacks
        got = ctx.rangeproof_verify_batch(nbits, m, proofs, coms)
        redone = ctx.verify_fallbacks - f0
        ctx.set_verify_mode(0)
        per_proof = ctx.rangeproof_verify_batch(nbits, m, proofs, coms)
        assert (got == per_proof).all()
        assert not got[bad].any() and got.sum() == k - len(bad)
        groups = {i // min(G, k) for i in bad}
        assert redone == sum(min(min(G, k), k - g * min(G, k)) for g in groups)
        for i in bad[:3] + [j for j in range(k) if j not in bad][:2]:
            assert bool(got[i]) == cref.rp_verify(proofs[i].tobytes(), [c.tobytes() for c in coms[i]], nbits)
    finally:
        ctx.set_verify_mode(0)


@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_packed_small_shape_kernels_same_bytes(cref, lanes, monkeypatch):
    """Small shapes in large batches run 4 / 8 / 16 lanes per MSM (k_rp_p3g / k_rp_p10g) instead of a warp per MSM: forced here on small
    batches (DAPOL_RP_PACK_MIN_K = 1), the proofs are byte for byte the oracle's and verify."""
    from dapol_b200 import Context
    monkeypatch.setenv("DAPOL_RP_PACK_MIN_K", "1")
    monkeypatch.setenv("DAPOL_RP_PACK_MAX_N", "128")
    monkeypatch.setenv("DAPOL_RP_PACK_LANES", str(lanes))
    c = Context(0, 15)
    c.set_rangeproof_window(12)
    try:
        for nbits, m, k in [(64, 1, 70), (64, 2, 9), (32, 4, 5), (8, 1, 33), (16, 8, 3)]:
            rnd = random.Random(nbits * 3 + m + lanes)
            vals, bl, streams, bases = _batch(rnd, nbits, m, k)
            proofs = c.rangeproof_prove_batch(nbits, vals, bl, SEED, streams, bases)
            for i in range(min(k, 12)):
                want = cref.rp_prove([int(x) for x in vals[i]], [b.tobytes() for b in bl[i]], SEED, int(streams[i]), int(bases[i]), nbits)
                assert proofs[i].tobytes() == want, (nbits, m, i)
            assert c.rangeproof_verify_batch(nbits, m, proofs, _coms(cref, vals, bl)).all()
    finally:
        c.close()
