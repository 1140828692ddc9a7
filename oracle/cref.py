"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the C oracle (oracle/c/dapol_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  `build()` compiles oracle/_build/libdapol_oracle.so with the committed Makefile.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdapol_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "c", f) for f in ("dapol_oracle.c", "ec.h", "hashes.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dor_tree_level_size.restype = C.c_uint64
        _lib.dor_tree_num_pads.restype = C.c_uint64
        _lib.dor_inclusion_proof_size.restype = C.c_uint64
        _lib.dor_tree_level_size.argtypes = [C.c_void_p, C.c_int]
        _lib.dor_tree_num_pads.argtypes = [C.c_void_p]
        _lib.dor_tree_free.argtypes = [C.c_void_p]
        _lib.dor_init()
    return _lib


def _b(x: bytes):
    return (C.c_uint8 * len(x)).from_buffer_copy(x)


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


def dlen(hash_id: int) -> int:
    """Digest length of D: 32 (blake3, Blake2s) or 64 (Blake2b, src/tests.rs:104-105)."""
    return 64 if hash_id == 2 else 32


def hash(hash_id: int, data: bytes) -> bytes:
    out = (C.c_uint8 * dlen(hash_id))()
    rc = lib().dor_hash(hash_id, _b(data) if data else None, C.c_size_t(len(data)), out)
    assert rc == 0
    return bytes(out)


def commit(v: int, r: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_commit(C.c_uint64(v), _b(r), out)
    return bytes(out)


def scalarmult_base(s: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_scalarmult_base(_b(s), out)
    return bytes(out)


def decompress_recompress(s: bytes):
    out = (C.c_uint8 * 32)()
    return bytes(out) if lib().dor_decompress_recompress(_b(s), out) else None


def from_uniform(b: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_from_uniform(_b(b), out)
    return bytes(out)


def point_add(a: bytes, b: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_point_add(_b(a), _b(b), out)
    return bytes(out)


def get_constant(which: int) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_get_constant(which, out)
    return bytes(out)


def rng_scalar(seed: bytes, k: int, stream: int = 0) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_rng_scalar(_b(seed), C.c_uint64(k), C.c_uint64(stream), out)
    return bytes(out)


def sc_mul(a: bytes, b: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_sc_mul(_b(a), _b(b), out)
    return bytes(out)


def sc_invert(a: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_sc_invert(_b(a), out)
    return bytes(out)


def merlin_test(label: bytes, mlabel: bytes, msg: bytes, clabel: bytes, n: int) -> bytes:
    out = (C.c_uint8 * n)()
    lib().dor_merlin_test(_b(label) if label else None, C.c_uint32(len(label)), mlabel, _b(msg), C.c_uint32(len(msg)), clabel, out, C.c_uint32(n))
    return bytes(out)


def bp_gen(is_h: int, party: int, i: int) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().dor_bp_gen(is_h, C.c_uint32(party), C.c_uint32(i), out)
    return bytes(out)


def pack_ids(ids):
    off = np.zeros(len(ids) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in ids], dtype=np.uint64)
    blob = np.frombuffer(b"".join(ids) or b"\0", dtype=np.uint8).copy()
    return blob, off


def derive_leaves(hash_id, iid_blob, iid_off, eid_blob, eid_off, audit_seed: bytes, height: int):
    """-> (rc, idx[n] u64 in INPUT order, blind[n,32] u8, err_pos)."""
    n = len(iid_off) - 1
    idx = np.zeros(max(n, 1), dtype=np.uint64)
    blind = np.zeros((max(n, 1), 32), dtype=np.uint8)
    err = C.c_uint64(0)
    ib, ibp = _np(iid_blob, np.uint8)
    io, iop = _np(iid_off, np.uint64)
    eb, ebp = _np(eid_blob, np.uint8)
    eo, eop = _np(eid_off, np.uint64)
    rc = lib().dor_derive_leaves(hash_id, C.c_uint64(n), ibp, iop, ebp, eop, _b(audit_seed) if audit_seed else None,
                                 C.c_uint64(len(audit_seed)), height, idx.ctypes.data_as(C.c_void_p),
                                 blind.ctypes.data_as(C.c_void_p), C.byref(err))
    return rc, idx[:n], blind[:n], err.value


class Tree:
    def __init__(self, hash_id, height, idx_sorted, values, blindings, pad_seed: bytes, pad_base=0, nthreads=0, positional=False, level_base=None):
        """positional=True: padding blindings keyed by (level, index) instead of the creation-order stream.
        level_base (u64[height + 1]): one shard of a prefix-split tree -- the r-th padding node of level h draws block
        level_base[h] + r (the blocks the single-tree creation order gives this subtree)."""
        idx, ip = _np(idx_sorted, np.uint64)
        val, vp = _np(values, np.uint64)
        bl, bp = _np(blindings, np.uint8)
        assert bl.size == 32 * idx.size
        h = C.c_void_p()
        if level_base is not None:
            lb, lbp = _np(level_base, np.uint64)
            assert lb.size == height + 1
            rc = lib().dor_tree_build_shard(hash_id, height, C.c_uint64(idx.size), ip, vp, bp, _b(pad_seed), lbp, nthreads, C.byref(h))
        elif positional:
            rc = lib().dor_tree_build_positional(hash_id, height, C.c_uint64(idx.size), ip, vp, bp, _b(pad_seed), nthreads, C.byref(h))
        else:
            rc = lib().dor_tree_build(hash_id, height, C.c_uint64(idx.size), ip, vp, bp, _b(pad_seed), C.c_uint64(pad_base), nthreads, C.byref(h))
        if rc:
            raise ValueError(f"dor_tree_build rc={rc}")
        self.h, self.hash_id, self.height = h, hash_id, height

    def __del__(self):
        if getattr(self, "h", None):
            lib().dor_tree_free(self.h)
            self.h = None

    def level(self, h):
        n = lib().dor_tree_level_size(self.h, h)
        idx = np.zeros(n, np.uint64); v = np.zeros(n, np.uint64)
        r = np.zeros((n, 32), np.uint8); comc = np.zeros((n, 32), np.uint8); hs = np.zeros((n, dlen(self.hash_id)), np.uint8)
        pad = np.zeros(n, np.uint8)
        lib().dor_tree_level_copy(self.h, h, *[a.ctypes.data_as(C.c_void_p) for a in (idx, v, r, comc, hs, pad)])
        return dict(idx=idx, v=v, r=r, comc=comc, hash=hs, is_pad=pad)

    @property
    def num_pads(self):
        return lib().dor_tree_num_pads(self.h)

    def root(self):
        l = self.level(0)
        return dict(v=int(l["v"][0]), r=l["r"][0].tobytes(), comc=l["comc"][0].tobytes(), hash=l["hash"][0].tobytes())

    def path(self, leaf_idx):
        H = self.height
        v = np.zeros(H, np.uint64); r = np.zeros((H, 32), np.uint8); c = np.zeros((H, 32), np.uint8); hs = np.zeros((H, dlen(self.hash_id)), np.uint8)
        rc = lib().dor_tree_path(self.h, C.c_uint64(leaf_idx), *[a.ctypes.data_as(C.c_void_p) for a in (v, r, c, hs)])
        if rc:
            return None
        return dict(v=v, r=r, comc=c, hash=hs)

    def get_node(self, h, idx):
        v = C.c_uint64(); pad = C.c_uint8()
        r = (C.c_uint8 * 32)(); c = (C.c_uint8 * 32)(); hs = (C.c_uint8 * dlen(self.hash_id))()
        rc = lib().dor_tree_get_node(self.h, h, C.c_uint64(idx), C.byref(v), r, c, hs, C.byref(pad))
        if rc:
            return None
        return dict(v=v.value, r=bytes(r), comc=bytes(c), hash=bytes(hs), is_pad=pad.value)

    def prove_inclusion(self, leaf_idx, agg, policy, seed: bytes):
        lib().dor_inclusion_proof_size_d.restype = C.c_uint64
        cap = lib().dor_inclusion_proof_size_d(self.height, C.c_uint64(agg), policy, dlen(self.hash_id))
        if cap == 0:
            return None
        out = (C.c_uint8 * cap)()
        n = C.c_uint64()
        rc = lib().dor_prove_inclusion(self.h, C.c_uint64(leaf_idx), C.c_uint64(agg), policy, _b(seed), out, C.c_uint64(cap), C.byref(n))
        if rc:
            return None
        return bytes(out[: n.value])


def leaf_id_hashes(hash_id, iid_blob, iid_off, eid_blob, eid_off, audit_seed: bytes) -> np.ndarray:
    """Opt-in leaf hash of the DAPOL+ paper (include/dapol_b200.h, DAPOL_LEAF_HASH_ID_SALT), per liability in input order."""
    n = len(iid_off) - 1
    out = np.zeros((max(n, 1), 32), np.uint8)
    ib, ibp = _np(iid_blob, np.uint8); io, iop = _np(iid_off, np.uint64)
    eb, ebp = _np(eid_blob, np.uint8); eo, eop = _np(eid_off, np.uint64)
    lib().dor_leaf_id_hashes(hash_id, C.c_uint64(n), ibp, iop, ebp, eop, _b(audit_seed) if audit_seed else None, C.c_uint64(len(audit_seed)),
                             out.ctypes.data_as(C.c_void_p))
    return out[:n]


def tree_with_leaf_hashes(hash_id, height, idx_sorted, values, blindings, leaf_hashes, pad_seed: bytes, pad_base=0) -> "Tree":
    """dor_tree_build with the leaves' hashes given (sorted like the leaves) instead of D(compress(com))."""
    idx, ip = _np(idx_sorted, np.uint64); val, vp = _np(values, np.uint64); bl, bp = _np(blindings, np.uint8)
    lh, lp = _np(leaf_hashes, np.uint8)
    assert bl.size == 32 * idx.size == lh.size
    h = C.c_void_p()
    rc = lib().dor_tree_build_leaf_hashes(hash_id, height, C.c_uint64(idx.size), ip, vp, bp, lp, _b(pad_seed), C.c_uint64(pad_base), 0, C.byref(h))
    if rc:
        raise ValueError(f"dor_tree_build_leaf_hashes rc={rc}")
    t = Tree.__new__(Tree)
    t.h, t.hash_id, t.height = h, hash_id, height
    return t


def prove_inclusion_batch(tree: "Tree", leaf_idxs, agg, policy, seed: bytes):
    """Dapol::generate_proof_batch (mod.rs:172-190): ONE DapolProof for several leaves (strictly increasing indexes)."""
    L = lib()
    L.dor_batch_proof_size_d.restype = C.c_uint64
    ia, iap = _np(leaf_idxs, np.uint64)
    cap = L.dor_batch_proof_size_d(tree.height, C.c_uint64(ia.size), iap, C.c_uint64(agg), policy, dlen(tree.hash_id))
    if cap == 0:
        return None
    out = (C.c_uint8 * cap)()
    n = C.c_uint64()
    rc = L.dor_prove_inclusion_batch(tree.h, C.c_uint64(ia.size), iap, C.c_uint64(agg), policy, _b(seed), out, C.c_uint64(cap), C.byref(n))
    return None if rc else bytes(out[: n.value])


def verify_inclusion_batch(hash_id, policy, proof: bytes, root_com, root_hash, leaf_coms, leaf_hashes) -> bool:
    """DapolProof::deserialize + verify_batch(root, leaves) (proof/mod.rs:49-54); leaves in index order."""
    return bool(lib().dor_verify_inclusion_batch(hash_id, policy, _b(proof) if proof else None, C.c_uint64(len(proof)), _b(root_com), _b(root_hash),
                                                 C.c_uint64(len(leaf_coms)), _b(b"".join(leaf_coms)), _b(b"".join(leaf_hashes))))


def rp_prove(values, blindings, seed: bytes, stream=0, base=0, nbits=64) -> bytes:
    m = len(values)
    vals, vp = _np(values, np.uint64)
    bl = b"".join(blindings)
    out = (C.c_uint8 * (32 * 64))()
    n = C.c_uint64()
    rc = lib().dor_rp_prove(nbits, m, vp, _b(bl), _b(seed), C.c_uint64(stream), C.c_uint64(base), out, C.byref(n))
    assert rc == 0, rc
    return bytes(out[: n.value])


def rp_verify(proof: bytes, commitments, nbits=64) -> bool:
    coms = b"".join(commitments)
    return bool(lib().dor_rp_verify(nbits, len(commitments), _b(proof) if proof else None, C.c_uint64(len(proof)), _b(coms)))


def verify_inclusion(hash_id, policy, proof: bytes, root_com, root_hash, leaf_com, leaf_hash) -> bool:
    return bool(lib().dor_verify_inclusion(hash_id, policy, _b(proof) if proof else None, C.c_uint64(len(proof)),
                                           _b(root_com), _b(root_hash), _b(leaf_com), _b(leaf_hash)))
