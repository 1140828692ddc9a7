"""CPU-side check of the product's range-proof device code: dapol_b200/csrc/rp_kernels.cuh compiled for the host
(tests/host_emu/emu_rp.cpp) with every pass of the batched prover / verifier driven in serial loops in the order the
CUDA orchestration uses, compared with the oracle: generators, window tables, merlin, byte-identical proofs under
the seeded-RNG contract, and accept / reject parity.  The GPU tests run the real kernels through the C ABI."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_emu", "libdapol_emu_rp.so")
SRC = os.path.join(HERE, "host_emu", "emu_rp.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "dapol_b200", "csrc")
SEED = hashlib.sha256(b"dapol-b200").digest()
L = 2 ** 252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def E():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SO, SRC], check=True)
    return C.CDLL(SO)


def B(x):
    return (C.c_uint8 * max(len(x), 1)).from_buffer_copy(x or b"\0")


def emu_prove(E, nbits, values, blindings, streams, bases, T=3):
    K, m = len(values), len(values[0])
    vals = np.array(values, np.uint64).reshape(K, m)
    bl = np.frombuffer(b"".join(b"".join(x) for x in blindings), np.uint8).copy()
    st = np.array(streams, np.uint64); bs = np.array(bases, np.uint64)
    lg = (nbits * m).bit_length() - 1
    plen = 32 * (9 + 2 * lg)
    out = np.zeros(K * plen, np.uint8)
    res = None
    # all rounds over the original generators, then (N >= 32) the hybrid rounds over folded generators: identical bytes
    for hybrid_min_n in (1 << 30, 32):
        E.emu_rp_set_hybrid_min_n(hybrid_min_n)
        rc = E.emu_rp_prove(nbits, m, C.c_uint64(K), vals.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), B(SEED),
                            st.ctypes.data_as(C.c_void_p), bs.ctypes.data_as(C.c_void_p), T, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        got = [out[i * plen:(i + 1) * plen].tobytes() for i in range(K)]
        assert res is None or res == got, "hybrid inner-product rounds changed the proof bytes"
        res = got
    E.emu_rp_set_hybrid_min_n(1024)
    return res


def emu_verify(E, nbits, m, proofs, coms, T=3):
    K = len(proofs)
    pr = np.frombuffer(b"".join(proofs), np.uint8).copy()
    cm = np.frombuffer(b"".join(b"".join(c) for c in coms), np.uint8).copy()
    res = None
    for groups in (8, 1, 4):  # threads per proof of V1 (Straus over every groups-th variable point): same verdicts
        ok = np.zeros(K, np.uint8)
        E.emu_rp_set_vgroups(groups)
        E.emu_rp_verify(nbits, m, C.c_uint64(K), pr.ctypes.data_as(C.c_void_p), cm.ctypes.data_as(C.c_void_p), T, ok.ctypes.data_as(C.c_void_p))
        assert res is None or res == [bool(x) for x in ok]
        res = [bool(x) for x in ok]
    return res


def emu_verify_batched(E, nbits, m, proofs, coms, G, cbits, T=3, wseed=SEED):
    """Group verdicts of the batched verifier (one random linear combination per G proofs, bucket method with cbits-bit windows)."""
    K = len(proofs)
    pr = np.frombuffer(b"".join(proofs), np.uint8).copy()
    cm = np.frombuffer(b"".join(b"".join(c) for c in coms), np.uint8).copy()
    gok = np.zeros((K + G - 1) // G, np.int32)
    E.emu_rp_verify_batched(nbits, m, C.c_uint64(K), pr.ctypes.data_as(C.c_void_p), cm.ctypes.data_as(C.c_void_p), T, G, cbits, B(wseed),
                            gok.ctypes.data_as(C.c_void_p))
    return [bool(x) for x in gok]


def test_generators_and_tables(E, cref):
    out = (C.c_uint8 * 32)()
    for is_h, party, i in [(0, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), (1, 1, 63), (0, 3, 17), (1, 2, 40)]:
        E.emu_rp_gen(is_h, party, i, out)
        assert bytes(out) == cref.bp_gen(is_h, party, i)
    E.emu_rp_gen(0, 0, 0, out)
    assert bytes(out).hex() == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"
    rnd = random.Random(3)
    for g in [0, 1, 63, 64, 127, 128, 129]:  # G_0[..], H_0[..], B, B_blinding
        for s in [1, 2, 7, 8, 9, L - 1, rnd.randrange(L), rnd.randrange(2 ** 256)]:
            assert E.emu_rp_table_check(g, B(s.to_bytes(32, "little")))


def test_merlin(E, cref):
    out = (C.c_uint8 * 32)()
    E.emu_merlin_test(B(b"test protocol"), 13, B(b"some label"), 10, B(b"some data"), 9, B(b"challenge"), 9, out)
    wide = cref.merlin_test(b"test protocol", b"some label", b"some data", b"challenge", 64)
    assert int.from_bytes(bytes(out), "little") == int.from_bytes(wide, "little") % L
    msg = bytes(range(200)) * 3  # crosses several STROBE blocks
    E.emu_merlin_test(None, 0, B(b"x"), 1, B(msg), len(msg), B(b"y"), 1, out)
    assert int.from_bytes(bytes(out), "little") == int.from_bytes(cref.merlin_test(b"", b"x", msg, b"y", 64), "little") % L


def _case(rnd, nbits, m):
    vals = [rnd.randrange(1 << nbits) for _ in range(m)]
    bls = [rnd.randrange(1 << 255).to_bytes(32, "little") for _ in range(m)]  # possibly unreduced (Scalar::from_bits)
    return vals, bls


@pytest.mark.parametrize("nbits,m,T", [(64, 1, 3), (64, 2, 5), (8, 2, 1), (32, 1, 4), (64, 4, 7), (16, 4, 2)])
def test_prover_bytes_and_verifier(E, cref, nbits, m, T):
    """SINGLE_PROOF_BYTE_NUM = 672 (src/range/mod.rs:18) for (64, 1); proofs byte-identical to the oracle's under the same
    seeded RNG; both verifiers accept; tampering and a wrong commitment are rejected by both."""
    rnd = random.Random(nbits * 10 + m)
    K = 3
    cases = [_case(rnd, nbits, m) for _ in range(K)]
    cases[0][0][0] = (1 << nbits) - 1
    cases[1][0][0] = 0
    streams, bases = [7, 8, 2 ** 40 + 1], [0, 5 << 32, 3]
    proofs = emu_prove(E, nbits, [c[0] for c in cases], [c[1] for c in cases], streams, bases, T)
    coms = []
    for (vals, bls), st, bs, pf in zip(cases, streams, bases, proofs):
        assert pf == cref.rp_prove(vals, bls, SEED, st, bs, nbits)
        if (nbits, m) == (64, 1):
            assert len(pf) == 672
        cm = [cref.commit(v, (int.from_bytes(b, "little") % L).to_bytes(32, "little")) for v, b in zip(vals, bls)]
        assert cref.rp_verify(pf, cm, nbits)
        coms.append(cm)
    assert emu_verify(E, nbits, m, proofs, coms, T) == [True] * K
    # rejections: flipped bit anywhere in the proof, wrong commitment, non-canonical scalar, identity point
    bad, badc = [], []
    for pos in [0, 40, 70, 100, 130, 170, 200, 230, len(proofs[0]) - 40, len(proofs[0]) - 1]:
        b = bytearray(proofs[0]); b[pos] ^= 4
        bad.append(bytes(b)); badc.append(coms[0])
    bad.append(proofs[1]); badc.append(coms[2])                                    # someone else's commitments
    b = bytearray(proofs[2]); b[128:160] = (L + 5).to_bytes(32, "little"); bad.append(bytes(b)); badc.append(coms[2])
    b = bytearray(proofs[2]); b[32:64] = bytes(32); bad.append(bytes(b)); badc.append(coms[2])
    got = emu_verify(E, nbits, m, bad, badc, T)
    want = [cref.rp_verify(p, c, nbits) for p, c in zip(bad, badc)]
    assert got == want == [False] * len(bad)


def test_out_of_range_value_rejected(E, cref):
    """A commitment to 2^n with a proof made for 2^n - 1 (and a proof for v mod 2^n of a larger v) does not verify."""
    rnd = random.Random(5)
    bl = rnd.randrange(L).to_bytes(32, "little")
    pf = emu_prove(E, 8, [[255]], [[bl]], [0], [0])[0]
    assert emu_verify(E, 8, 1, [pf], [[cref.commit(255, bl)]]) == [True]
    assert emu_verify(E, 8, 1, [pf], [[cref.commit(256, bl)]]) == [False]
    pf = emu_prove(E, 8, [[300]], [[bl]], [0], [0])[0]  # bits of 300 & 0xff are proven, the commitment holds 300
    assert emu_verify(E, 8, 1, [pf], [[cref.commit(300, bl)]]) == [cref.rp_verify(pf, [cref.commit(300, bl)], 8)] == [False]


@pytest.mark.parametrize("nbits,m,G,cbits", [(8, 1, 3, 3), (8, 2, 4, 5), (16, 1, 2, 7), (64, 1, 5, 4)])
def test_batched_verifier_group_verdicts(E, cref, nbits, m, G, cbits):
    """Batched verification by the bucket method (Pippenger; dapol_ctx_set_verify_mode): a group's random linear combination is the
    identity iff every proof of the group verifies on its own (the per-proof verifier and the oracle), for any grouping / window /
    weights; one bad proof fails exactly its group."""
    rnd = random.Random(nbits + 7 * m + G)
    K = 2 * G + 1  # two full groups and a partial one
    cases = [_case(rnd, nbits, m) for _ in range(K)]
    proofs = emu_prove(E, nbits, [c[0] for c in cases], [c[1] for c in cases], list(range(K)), [0] * K, 2)
    coms = [[cref.commit(v, (int.from_bytes(b, "little") % L).to_bytes(32, "little")) for v, b in zip(vals, bls)] for vals, bls in cases]
    assert emu_verify_batched(E, nbits, m, proofs, coms, G, cbits) == [True, True, True]
    assert emu_verify_batched(E, nbits, m, proofs, coms, K, cbits + 1, wseed=bytes(range(32))) == [True]          # one group of everything
    for victim, how in ((1, "bit"), (G, "com"), (2 * G, "scalar"), (G + 1, "point")):
        bad, badc = list(proofs), [list(c) for c in coms]
        if how == "bit":
            bb = bytearray(bad[victim]); bb[70] ^= 4; bad[victim] = bytes(bb)                                      # T_1 changed
        elif how == "com":
            badc[victim][0] = cref.commit(cases[victim][0][0] + 1, (int.from_bytes(cases[victim][1][0], "little") % L).to_bytes(32, "little"))
        elif how == "scalar":
            bb = bytearray(bad[victim]); bb[128:160] = (L + 5).to_bytes(32, "little"); bad[victim] = bytes(bb)    # non-canonical t_x
        else:
            bb = bytearray(bad[victim]); bb[0:32] = (2).to_bytes(32, "little"); bad[victim] = bytes(bb)           # A does not decompress
        per_proof = emu_verify(E, nbits, m, bad, badc, 2)
        assert per_proof == [i != victim for i in range(K)] == [cref.rp_verify(p, c, nbits) for p, c in zip(bad, badc)]
        want = [all(per_proof[g * G:(g + 1) * G]) for g in range((K + G - 1) // G)]
        assert emu_verify_batched(E, nbits, m, bad, badc, G, cbits) == want
