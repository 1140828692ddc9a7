"""CPU-side check of the product's device headers: dapol_b200/csrc/*.cuh compiled for the host with the PTX
carry-chain instructions emulated (tests/host_emu/emu.cpp), per-thread kernel bodies driven in serial loops,
compared with the oracle.  This validates field/scalar/group arithmetic, hashes, the comb, the structure pass,
leaf derivation + collision fix-point and the merge logic without a GPU.  The GPU tests run the real kernels."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_emu", "libdapol_emu.so")
SRC = os.path.join(HERE, "host_emu", "emu.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "dapol_b200", "csrc")


@pytest.fixture(scope="module")
def E():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.emu_tree_build.restype = C.c_void_p
    lib.emu_tree_level_size.restype = C.c_uint64
    lib.emu_tree_num_pads.restype = C.c_uint64
    lib.emu_tree_level_size.argtypes = [C.c_void_p, C.c_int]
    lib.emu_tree_num_pads.argtypes = [C.c_void_p]
    lib.emu_tree_free.argtypes = [C.c_void_p]
    return lib


def B(x):
    return (C.c_uint8 * len(x)).from_buffer_copy(x) if len(x) else None


P = 2 ** 255 - 19
L = 2 ** 252 + 27742317777372353535851937790883648493
PAD_SEED = hashlib.sha256(b"dapol-b200").digest()


def test_field_ops(E):
    rnd = random.Random(7)

    def fe_op(op, a, b=0):
        out = (C.c_uint8 * 32)()
        E.emu_fe_op(op, B(a.to_bytes(32, "little")), B(b.to_bytes(32, "little")), out)
        return int.from_bytes(bytes(out), "little")
    edge = [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 37, 2 ** 256 - 1, 2 ** 256 - 38, 2 ** 255, 2 ** 255 - 1,
            2 ** 255 + 18, 2 ** 256 - 39, 0xFFFFFFFF, 2 ** 32, (2 ** 256 - 1) ^ (2 ** 32 - 1)]
    # the product is one level of subtractive Karatsuba over the 128-bit halves: equal halves (zero difference), either sign of
    # a0 - a1 and b1 - b0, extreme halves
    M128 = 2 ** 128 - 1
    halves = [0, 1, 2, M128, M128 - 1, 2 ** 127, 2 ** 64, 2 ** 64 - 1, 2 ** 96 + 5, rnd.randrange(2 ** 128), rnd.randrange(2 ** 128)]
    edge += [lo | (hi << 128) for lo in halves for hi in halves]
    vals = edge + [rnd.randrange(2 ** 256) for _ in range(300)]
    out = (C.c_uint8 * 64)()
    for a in edge[-121:]:
        for b in edge[-121::7] + [rnd.randrange(2 ** 256)]:
            E.emu_mul_wide(B(a.to_bytes(32, "little")), B(b.to_bytes(32, "little")), out, 0)
            assert int.from_bytes(bytes(out), "little") == a * b, (hex(a), hex(b))
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        E.emu_mul_wide(B(a.to_bytes(32, "little")), B(b.to_bytes(32, "little")), out, 0)
        assert int.from_bytes(bytes(out), "little") == a * b
        E.emu_mul_wide(B(a.to_bytes(32, "little")), B(b.to_bytes(32, "little")), out, 1)
        assert int.from_bytes(bytes(out), "little") == a * a
        assert fe_op(0, a, b) == a * b % P and fe_op(1, a) == a * a % P
        assert fe_op(2, a, b) == (a + b) % P and fe_op(3, a, b) == (a - b) % P and fe_op(5, a) == (-a) % P
    for a in vals[:40]:
        if a % P:
            assert fe_op(4, a) == pow(a, P - 2, P)
        assert fe_op(6, a) == pow(a, (P - 5) // 8, P)


def test_scalar_ops(E):
    rnd = random.Random(8)

    def sc_op(op, a, b=0):
        out = (C.c_uint8 * 32)()
        E.emu_sc_op(op, B(a.to_bytes(64, "little")), B(b.to_bytes(32, "little")), out)
        return int.from_bytes(bytes(out), "little")
    sv = [0, 1, L - 1, L, L + 1, 2 * L, 2 ** 252, 2 ** 253 - 1, 2 ** 255 - 1, 2 ** 256 - 1] + [rnd.randrange(2 ** 256) for _ in range(100)]
    for i, a in enumerate(sv):
        b = sv[(i * 5 + 1) % len(sv)] % L
        assert sc_op(0, a, b) == a * b % L and sc_op(5, a) == a % L
        assert sc_op(1, a % L, b) == (a + b) % L and sc_op(2, a % L, b) == (a - b) % L and sc_op(6, a % L) == (-a) % L
        w = rnd.randrange(2 ** 512)
        assert sc_op(4, w) == w % L
        half = pow(2, -1, L)
        assert sc_op(7, w) == w % L and sc_op(8, w) == w * half % L and sc_op(9, a) == a * half % L  # padding scalars: r and r / 2
        assert E.emu_sc_is_canonical(B(a.to_bytes(32, "little"))) == (1 if a < L else 0)
    assert sc_op(4, 2 ** 512 - 1) == (2 ** 512 - 1) % L
    for w in (0, 1, L, L - 1, 2 ** 256, 2 ** 256 - 1, 2 ** 512 - 1, L << 256):
        assert sc_op(7, w) == w % L and sc_op(8, w) == w * pow(2, -1, L) % L
    for a in sv[10:16]:
        assert sc_op(3, a % L) == pow(a % L, L - 2, L)


def test_group_ops(E, pyref):
    o = pyref
    rnd = random.Random(9)
    out = (C.c_uint8 * 32)()
    for k in [1, 2, 3, 12345, L - 1, rnd.randrange(L)]:
        E.emu_scalarmult_base(B(k.to_bytes(32, "little")), out, 0)
        assert bytes(out) == o.compress(o.pt_mul(k, o.BASEPOINT))
        E.emu_scalarmult_base(B(k.to_bytes(32, "little")), out, 1)
        assert bytes(out) == o.compress(o.pt_mul(k, o.B_BLINDING))
    for i in range(40):
        s = rnd.randbytes(32) if i > 5 else o.compress(o.pt_mul(i + 1, o.BASEPOINT))
        ok = E.emu_decompress_recompress(B(s), out)
        assert bool(ok) == (o.decompress(s) is not None)
        if ok:
            assert bytes(out) == s
    for i in range(10):
        u = rnd.randbytes(64)
        E.emu_from_uniform(B(u), out)
        assert bytes(out) == o.compress(o.from_uniform_bytes(u))
    for w in (4, 5, 8):
        for v, r in [(0, 1), (5, 7), (2 ** 64 - 1, 2 ** 255 - 1), (rnd.randrange(2 ** 64), rnd.randrange(2 ** 255))]:
            E.emu_commit(w, C.c_uint64(v), B(r.to_bytes(32, "little")), out)
            assert bytes(out) == o.compress(o.pedersen_commit(v, r)), (w, v, r)


def test_hashes(E, pyref):
    import blake3
    rnd = random.Random(10)
    out = (C.c_uint8 * 32)()
    for hid, fn in ((0, lambda d: blake3.blake3(d).digest()), (1, lambda d: hashlib.blake2s(d).digest())):
        d = rnd.randbytes(32); E.emu_hash32(hid, B(d), out); assert bytes(out) == fn(d)
        d = rnd.randbytes(128); E.emu_hash128(hid, B(d), out); assert bytes(out) == fn(d)
        # ids of any length hash as the reference's D does (mod.rs:347-353): BLAKE3 beyond one 1024-byte chunk = tree mode
        for n in [0, 1, 3, 4, 31, 32, 63, 64, 65, 100, 127, 128, 129, 500, 1023, 1024, 1025, 1088, 2047, 2048, 2049, 3072, 3073, 4096, 5000,
                  7 * 1024, 8 * 1024 + 1, 31 * 1024 + 7, 65536, 100000]:
            d = rnd.randbytes(n)
            assert E.emu_hash_bytes(hid, B(d), n, out) == 0 and bytes(out) == fn(d), (hid, n)
    key = bytes(range(32))
    ks = (C.c_uint8 * 64)()
    E.emu_chacha(B(key), C.c_uint64(1 | (0x09000000 << 32)), C.c_uint64(0x4A000000), ks)
    assert bytes(ks) == pyref.chacha20_block(key, 1 | (0x09000000 << 32), 0x4A000000)
    st = bytearray(rnd.randbytes(200)); st2 = bytearray(st)
    pyref.keccak_f(st2)
    buf = (C.c_uint8 * 200).from_buffer_copy(bytes(st))
    E.emu_keccak(buf)
    assert bytes(buf) == bytes(st2)


@pytest.mark.parametrize("positional", [False, True])
@pytest.mark.parametrize("hid,H,n", [(0, 6, 9), (1, 4, 4), (0, 10, 200), (0, 16, 64), (0, 3, 4), (0, 1, 1), (0, 1, 2), (0, 5, 1), (0, 40, 20), (0, 64, 6),
                                     (2, 6, 9), (2, 10, 100), (2, 1, 1), (2, 0, 1), (2, 33, 12)])  # hid 2: Blake2b, 64-byte digests (N2)
def test_tree_bodies_vs_oracle(E, cref, hid, H, n, positional):
    """positional: padding blindings keyed by (level, index) instead of the creation-order stream."""
    rnd = random.Random(100 + H + n)
    idx = np.array(sorted({rnd.randrange(2 ** H) for _ in range(n)} if H > 20 else rnd.sample(range(2 ** H), n)), dtype=np.uint64)
    n = len(idx)
    vals_ = np.array([rnd.randrange(2 ** 32) for _ in range(n)], dtype=np.uint64)
    bl = np.frombuffer(rnd.randbytes(32 * n), dtype=np.uint8).copy().reshape(n, 32); bl[:, 31] &= 0x7F
    T = cref.Tree(hid, H, idx, vals_, bl, PAD_SEED, 5, positional=positional)
    E.emu_set_padding_mode(int(positional))
    t = E.emu_tree_build(hid, H, C.c_uint64(n), idx.ctypes.data_as(C.c_void_p), vals_.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), B(PAD_SEED), C.c_uint64(5))
    E.emu_set_padding_mode(0)
    assert t
    assert E.emu_tree_num_pads(t) == T.num_pads
    for h in range(H + 1):
        Lc = T.level(h); m = E.emu_tree_level_size(t, h)
        assert m == len(Lc["idx"])
        i2 = np.zeros(m, np.uint64); v2 = np.zeros(m, np.uint64); r2 = np.zeros((m, 32), np.uint8); c2 = np.zeros((m, 32), np.uint8)
        h2 = np.zeros((m, 64 if hid == 2 else 32), np.uint8); p2 = np.zeros(m, np.uint8)
        E.emu_tree_level_copy(C.c_void_p(t), h, *[a.ctypes.data_as(C.c_void_p) for a in (i2, v2, r2, c2, h2, p2)])
        assert (i2 == Lc["idx"]).all() and (v2 == Lc["v"]).all() and (c2 == Lc["comc"]).all() and (h2 == Lc["hash"]).all() and (p2 == Lc["is_pad"]).all()
        assert [int.from_bytes(x.tobytes(), "little") % L for x in r2] == [int.from_bytes(x.tobytes(), "little") % L for x in Lc["r"]]
    E.emu_tree_free(C.c_void_p(t))


def _emu_tree_levels(E, t, H):
    out = []
    for h in range(H + 1):
        m = E.emu_tree_level_size(t, h)
        i2 = np.zeros(m, np.uint64); v2 = np.zeros(m, np.uint64); r2 = np.zeros((m, 32), np.uint8); c2 = np.zeros((m, 32), np.uint8)
        h2 = np.zeros((m, 32), np.uint8); p2 = np.zeros(m, np.uint8)
        E.emu_tree_level_copy(C.c_void_p(t), h, *[a.ctypes.data_as(C.c_void_p) for a in (i2, v2, r2, c2, h2, p2)])
        out.append(dict(idx=i2, v=v2, r=r2, comc=c2, hash=h2, is_pad=p2))
    return out


def test_identity_commitments_in_batch(E, cref):
    """Leaves whose commitment is the identity (v = 0 with r = 0 or r = l) make the batched inversion's input zero:
    the batch must still give every node of the tree the oracle's bytes (the zero is masked out of the product)."""
    H, n = 5, 7
    idx = np.array([1, 2, 3, 9, 17, 18, 30], np.uint64)
    vals_ = np.array([0, 0, 5, 0, 7, 0, 0], np.uint64)
    bl = np.zeros((n, 32), np.uint8)
    bl[1] = np.frombuffer(L.to_bytes(32, "little"), np.uint8)   # unreduced l == 0
    bl[2] = 9; bl[4] = 1
    bl[6] = np.frombuffer((2 * L).to_bytes(32, "little"), np.uint8)
    T = cref.Tree(0, H, idx, vals_, bl, PAD_SEED, 0)
    t = E.emu_tree_build(0, H, C.c_uint64(n), idx.ctypes.data_as(C.c_void_p), vals_.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), B(PAD_SEED), C.c_uint64(0))
    assert t
    lv = _emu_tree_levels(E, t, H)
    assert bytes(lv[H]["comc"][list(lv[H]["idx"]).index(1)]) == bytes(32)
    for h in range(H + 1):
        Lc = T.level(h)
        assert (lv[h]["comc"] == Lc["comc"]).all() and (lv[h]["hash"] == Lc["hash"]).all(), h
    E.emu_tree_free(C.c_void_p(t))


def test_unsorted_leaves_rejected(E):
    idx = np.array([5, 3], np.uint64); v = np.zeros(2, np.uint64); bl = np.zeros((2, 32), np.uint8)
    assert not E.emu_tree_build(0, 4, C.c_uint64(2), idx.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), B(PAD_SEED), C.c_uint64(0))


@pytest.mark.parametrize("hid,H,n,dup", [(1, 4, 4, False), (0, 7, 64, False), (0, 10, 300, False), (1, 8, 100, True), (0, 6, 32, False)])
def test_leaf_derivation_bodies_vs_oracle(E, cref, hid, H, n, dup):
    rnd = random.Random(200 + H + n)
    ids = [rnd.randbytes(rnd.randrange(0, 40)) + i.to_bytes(2, "little") for i in range(n)]
    eids = [rnd.randbytes(rnd.randrange(0, 70)) for i in range(n)]
    seed = rnd.randbytes(20)
    if H == 4:  # the reference KAT, src/dapol/tests.rs:38-84
        ids, eids, seed = [b"a", b"b", b"c", b"d"], [b"w", b"x", b"y", b"z"], b"test"
    if dup:
        ids[57] = ids[13]
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc0, idx0, bl0, err0 = cref.derive_leaves(hid, ib, io, eb, eo, seed, H)
    idx1 = np.zeros(n, np.uint64); bl1 = np.zeros((n, 32), np.uint8); err1 = C.c_uint64(0)
    E.emu_derive_leaves.restype = C.c_int
    rc1 = E.emu_derive_leaves(hid, C.c_uint64(n), ib.ctypes.data_as(C.c_void_p), io.ctypes.data_as(C.c_void_p), eb.ctypes.data_as(C.c_void_p),
                              eo.ctypes.data_as(C.c_void_p), B(seed), len(seed), H, idx1.ctypes.data_as(C.c_void_p), bl1.ctypes.data_as(C.c_void_p), C.byref(err1))
    assert rc0 == rc1
    if rc0 == 0:
        assert (idx0 == idx1).all() and (bl0 == bl1).all()
        if H == 4:
            assert list(idx1) == [7, 12, 2, 4]
    else:
        assert err0 == err1.value == 57


def test_id_salt_leaf_hash_body(E, cref):
    """Opt-in id / salt leaf hash (DAPOL_LEAF_HASH_ID_SALT): the kernel body against the C oracle (itself pinned to hashlib / blake3 in
    tests/test_oracle_pins.py), ids and external ids of every length class incl. empty and longer than one BLAKE3 chunk."""
    rnd = random.Random(31)
    ids = [b"user-%d" % i for i in range(40)]
    eids = [rnd.randbytes(rnd.choice([0, 1, 4, 33, 64, 119, 1024, 1200, 2500])) for _ in range(40)]
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    for hid in (0, 1):
        want = cref.leaf_id_hashes(hid, ib, io, eb, eo, b"id-salt")
        out = np.zeros((40, 32), np.uint8)
        rc = E.emu_leaf_id_hashes(hid, C.c_uint64(40), ib.ctypes.data_as(C.c_void_p), io.ctypes.data_as(C.c_void_p), eb.ctypes.data_as(C.c_void_p),
                                  eo.ctypes.data_as(C.c_void_p), B(b"id-salt"), 7, 10, out.ctypes.data_as(C.c_void_p))
        assert rc == 0 and (out == want).all()
