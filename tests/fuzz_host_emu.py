"""Randomised cross-check of the product's kernel bodies (host emulation, tests/host_emu) against the CPU oracle -- run by hand
(not collected by pytest):  python tests/fuzz_host_emu.py [seconds=120]
Trees: random shapes up to height 64, both digests, both padding modes, edge values and blindings; commitments with edge
scalars; range proofs: random shapes with N <= 128, both inner-product round styles, three verifier group counts, tampering.
Last run of round 1: 3508 trees, 70160 commitments, 811 proofs, no mismatch."""
import ctypes as C
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np

from oracle import cref
import test_host_emu_rp as T

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
L = 2**252 + 27742317777372353535851937790883648493
B = lambda b: (C.c_uint8 * len(b)).from_buffer_copy(b)


def fuzz_trees(seconds):
    E_ = C.CDLL(os.path.join(HERE, 'host_emu', 'libdapol_emu.so'))
    E_.emu_tree_build.restype = C.c_void_p
    E_.emu_tree_level_size.restype = C.c_uint64; E_.emu_tree_level_size.argtypes = [C.c_void_p, C.c_int]
    E_.emu_tree_free.argtypes = [C.c_void_p]
    rnd = random.Random(int(time.time()))
    t0 = time.time(); trees = commits = 0
    while time.time() - t0 < seconds:
        # commitments with edge scalars
        for _ in range(20):
            v = rnd.choice([0, 1, 2**64 - 1, rnd.randrange(2**64), rnd.randrange(2**32)])
            r = rnd.choice([0, 1, L - 1, L, L + 1, 2**255 - 1, 2**252, rnd.randrange(2**255), rnd.randrange(2**255)])
            w = rnd.choice([4, 5, 8])
            out = (C.c_uint8 * 32)()
            E_.emu_commit(w, C.c_uint64(v), B(r.to_bytes(32, 'little')), out)
            assert bytes(out) == cref.commit(v, r.to_bytes(32, 'little')), (w, v, r)
            commits += 1
        H = rnd.choice([1, 2, 3, 5, 8, 13, 21, 33, 47, 64])
        n = rnd.randrange(1, min(40, 2**H) + 1)
        idx = np.array(sorted(rnd.sample(range(2**H), n)) if H <= 20 else sorted({rnd.randrange(2**H) for _ in range(n)}), np.uint64)
        n = len(idx)
        vals = np.array([rnd.choice([0, 2**64 - 1, rnd.randrange(2**40)]) for _ in range(n)], np.uint64)
        bl = np.frombuffer(rnd.randbytes(32 * n), np.uint8).copy().reshape(n, 32); bl[:, 31] &= 0x7F
        if rnd.random() < 0.3: bl[0] = 0
        hid = rnd.choice([0, 1]); pos = rnd.choice([0, 1]); base = rnd.choice([0, 5, 2**40])
        seed = rnd.randbytes(32)
        T = cref.Tree(hid, H, idx, vals, bl, seed, base, positional=bool(pos))
        E_.emu_set_padding_mode(pos)
        t = E_.emu_tree_build(hid, H, C.c_uint64(n), idx.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), B(seed), C.c_uint64(base))
        E_.emu_set_padding_mode(0)
        assert t
        for h in range(H + 1):
            Lc = T.level(h); m = E_.emu_tree_level_size(t, h)
            assert m == len(Lc['idx'])
            i2 = np.zeros(m, np.uint64); v2 = np.zeros(m, np.uint64); r2 = np.zeros((m, 32), np.uint8); c2 = np.zeros((m, 32), np.uint8)
            h2 = np.zeros((m, 32), np.uint8); p2 = np.zeros(m, np.uint8)
            E_.emu_tree_level_copy(C.c_void_p(t), h, *[a.ctypes.data_as(C.c_void_p) for a in (i2, v2, r2, c2, h2, p2)])
            assert (i2 == Lc['idx']).all() and (v2 == Lc['v']).all() and (c2 == Lc['comc']).all() and (h2 == Lc['hash']).all() and (p2 == Lc['is_pad']).all(), (H, n, hid, pos, h)
        E_.emu_tree_free(C.c_void_p(t)); trees += 1
    print('fuzz ok:', trees, 'trees', commits, 'commitments')


def fuzz_rangeproofs(seconds):
    E_ = C.CDLL(os.path.join(HERE, 'host_emu', 'libdapol_emu_rp.so'))
    rnd = random.Random(int(time.time()))
    t0 = time.time(); n_ok = 0
    while time.time() - t0 < seconds:
        nbits, m = rnd.choice([(8, 1), (8, 2), (8, 4), (16, 2), (16, 4), (32, 1), (32, 2), (64, 1), (64, 2), (8, 8), (16, 8)])
        K = rnd.randrange(1, 3)
        vals = [[rnd.choice([0, 2**nbits - 1, rnd.randrange(2**nbits)]) for _ in range(m)] for _ in range(K)]
        bls = [[rnd.choice([0, 1, L - 1, rnd.randrange(2**255)]).to_bytes(32, 'little') for _ in range(m)] for _ in range(K)]
        streams = [rnd.randrange(2**40) for _ in range(K)]; bases = [rnd.choice([0, 1 << 32, 7]) for _ in range(K)]
        Tn = rnd.choice([1, 2, 3, 5, 8])
        proofs = T.emu_prove(E_, nbits, vals, bls, streams, bases, Tn)   # runs all-table and hybrid rounds, asserts identical bytes
        for p in range(K):
            assert proofs[p] == cref.rp_prove(vals[p], bls[p], T.SEED, streams[p], bases[p], nbits), (nbits, m, p)
        coms = [[cref.commit(v, b) for v, b in zip(vals[p], bls[p])] for p in range(K)]
        ok = T.emu_verify(E_, nbits, m, proofs, coms, Tn)                # runs 3 group counts, asserts same verdicts
        assert all(ok)
        bad = bytearray(proofs[0]); bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        okb = T.emu_verify(E_, nbits, m, [bytes(bad)], [coms[0]], Tn)
        assert okb == [cref.rp_verify(bytes(bad), coms[0], nbits)] == [False] or okb == [cref.rp_verify(bytes(bad), coms[0], nbits)]
        n_ok += K
    print('rp fuzz ok:', n_ok, 'proofs')


if __name__ == "__main__":
    fuzz_trees(SECONDS / 2)
    fuzz_rangeproofs(SECONDS / 2)
