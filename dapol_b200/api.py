"""Host-side mirror of the reference crate's public surface for the hot path (src/lib.rs:1-14),
over the C ABI.  Names, argument meaning and error behaviour follow the reference:

  Dapol::new / new_blank + build / root / root_raw / generate_proof*   src/dapol/mod.rs:100-208
  DapolNode::{new, get_value, get_blinding}                              src/dapol/node.rs:29-56
  DapolProofNode::{new, get_com, get_hash}                               src/proof/node.rs:30-49
  DapolError                                                             src/errors.rs:5-17
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass

import numpy as np

from . import _ffi

HASH_BLAKE3, HASH_BLAKE2S, HASH_BLAKE2B = 0, 1, 2  # Blake2b: 64-byte digests (src/tests.rs:104-105), new_blank + build trees only


def digest_len(hash_id: int) -> int:
    return 64 if hash_id == HASH_BLAKE2B else 32
POLICY_PADDING, POLICY_SPLITTING = 0, 1
MAX_TREE_HEIGHT = 64

_ERR_NAMES = {1: "TreeHeightTooBig", 2: "SparsityTooSmall", 3: "InvalidDigestSize", 4: "DuplicatedInternalId",
              5: "FailedToMapIndex", 16: "BadArgument", 17: "NotFound", 18: "BufferTooSmall", 19: "Cuda", 20: "Decode", 21: "Io"}


class DapolError(Exception):
    """src/errors.rs DapolError (codes 1-5) plus boundary errors (>= 16)."""

    def __init__(self, code: int, detail=None):
        L = _ffi.lib()
        msg = L.dapol_strerror(code).decode()
        if code == 19:
            msg += ": " + L.dapol_last_cuda_error().decode()
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}" + (f" (at input {detail})" if detail is not None else ""))
        self.code, self.detail = code, detail


def _check(rc, detail=None):
    if rc != 0:
        raise DapolError(rc, detail)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One CUDA device + precomputed generator tables (PedersenGens::default(), BulletproofGens)."""

    def __init__(self, device: int = 0, comb_window: int = 0):
        self._h = C.c_void_p()
        self._trees = weakref.WeakSet()  # trees built on this context: they must be destroyed before it (C-ABI contract)
        _check(_ffi.lib().dapol_ctx_create(device, comb_window, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            for t in list(getattr(self, "_trees", ())):
                t._free()
            _ffi.lib().dapol_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def set_stream(self, cuda_stream: int):
        """Run on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        _check(_ffi.lib().dapol_ctx_set_stream(self._h, cuda_stream))

    @property
    def kernel_launches(self) -> int:
        return _ffi.lib().dapol_kernel_launches(self._h)

    def params(self):
        """Tuning parameters in effect: comb window of the tree tables, nodes per shared inversion, range-proof window."""
        w, b, r = C.c_int(), C.c_int(), C.c_int()
        _check(_ffi.lib().dapol_ctx_params(self._h, C.byref(w), C.byref(b), C.byref(r)))
        return dict(comb_window=w.value, node_batch=b.value, rangeproof_window=r.value)

    def last_build_times(self):
        ms = np.zeros(5, np.float32)
        _check(_ffi.lib().dapol_last_build_times(self._h, _p(ms)))
        return dict(zip(("structure", "leaves", "padding", "merges", "total"), ms.tolist()))

    def commit_batch(self, values, blindings) -> np.ndarray:
        """PedersenGens::default().commit(v, r).compress() for a batch (node.rs:31)."""
        v = np.ascontiguousarray(values, np.uint64)
        b = np.ascontiguousarray(blindings, np.uint8).reshape(-1, 32)
        out = np.zeros((len(v), 32), np.uint8)
        _check(_ffi.lib().dapol_commit_batch(self._h, len(v), _p(v), _p(b), _p(out)))
        return out

    # -- range proofs (src/range/mod.rs:48-119), batched ----------------------------------------
    def set_padding_mode(self, positional: bool):
        """False (default): padding blindings from the creation-order stream (the reference's RNG, seeded).  True: keyed by
        (level, index) -- Paddable::padding(idx, secret) as a function of its arguments (node.rs:85-88 TODO); opt-in."""
        _check(_ffi.lib().dapol_ctx_set_padding_mode(self._h, 1 if positional else 0))

    def set_leaf_hash_mode(self, id_salt: bool):
        """False (default): leaf hash = D(compress(com)), the reference (node.rs:33-36).  True: the DAPOL+ paper's leaf hash
        D("leaf" || external_id || salt) with salt = D(audit_id || "salt_seed" || external_id) -- opt-in, not the reference's bytes."""
        _check(_ffi.lib().dapol_ctx_set_leaf_hash_mode(self._h, 1 if id_salt else 0))

    def set_rangeproof_table_budget(self, nbytes: int):
        """HBM budget of the generator tables under the automatic window (0 = default: 70 % of the free memory, at most 128 GB)."""
        _check(_ffi.lib().dapol_ctx_set_rangeproof_table_budget(self._h, nbytes))

    @property
    def rangeproof_table_bytes(self) -> int:
        return _ffi.lib().dapol_ctx_rangeproof_table_bytes(self._h)

    def set_verify_mode(self, group: int = 0, window_bits: int = 0, weight_seed: bytes | None = None):
        """group <= 1: every proof verified on its own (Straus).  group = G: G proofs per random-linear-combination check with the
        bucket method (Pippenger); failed groups are re-verified per proof, so the verdicts are the same (include/dapol_b200.h)."""
        sd = (C.c_uint8 * 32).from_buffer_copy(weight_seed) if weight_seed is not None else None
        _check(_ffi.lib().dapol_ctx_set_verify_mode(self._h, group, window_bits, sd))

    @property
    def verify_fallbacks(self) -> int:
        return _ffi.lib().dapol_ctx_verify_fallbacks(self._h)

    def set_rangeproof_window(self, window: int):
        _check(_ffi.lib().dapol_ctx_set_rangeproof_window(self._h, window))

    def rangeproof_last_times(self):
        ms = np.zeros(4, np.float32)
        _check(_ffi.lib().dapol_rangeproof_last_times(self._h, _p(ms)))
        return dict(zip(("total", "msm", "other", "table_build"), ms.tolist()))

    def rangeproof_last_kernel_times(self):
        """Device time of the last batch per kernel class (ms): the MSM passes split into k_rp_p10 / k_rp_p3 / hybrid rounds / verifier."""
        ms = np.zeros(8, np.float32)
        _check(_ffi.lib().dapol_rangeproof_last_kernel_times(self._h, _p(ms)))
        return dict(zip(("total", "msm", "other", "table_build", "p10", "p3", "hybrid", "verifier"), ms.tolist()))

    def rangeproof_prove_batch(self, nbits: int, values, blindings, seed: bytes, streams, base_blocks) -> np.ndarray:
        """k x generate_aggregated_range_proof (m = values.shape[1] parties; m = 1 is generate_single_range_proof).
        values [k][m] u64, blindings [k][m][32] Scalar bytes; returns [k][proof_len] bytes."""
        v = np.ascontiguousarray(values, np.uint64)
        k, m = v.shape
        b = np.ascontiguousarray(blindings, np.uint8).reshape(k, m, 32)
        st = np.ascontiguousarray(streams, np.uint64); bb = np.ascontiguousarray(base_blocks, np.uint64)
        size = _ffi.lib().dapol_rangeproof_size(nbits, m)
        if size == 0 or len(st) != k or len(bb) != k:
            raise DapolError(16)
        out = np.zeros((k, size), np.uint8)
        sd = (C.c_uint8 * 32).from_buffer_copy(seed)
        _check(_ffi.lib().dapol_rangeproof_prove_batch(self._h, nbits, m, k, _p(v), _p(b), sd, _p(st), _p(bb), _p(out)))
        return out

    def rangeproof_verify_batch(self, nbits: int, m: int, proofs, commitments) -> np.ndarray:
        """k x verify_aggregated_range_proof: proofs [k][proof_len], commitments [k][m][32]; returns bool[k]."""
        pr = np.ascontiguousarray(proofs, np.uint8)
        k = pr.shape[0]
        cm = np.ascontiguousarray(commitments, np.uint8).reshape(k, m, 32)
        ok = np.zeros(k, np.uint8)
        _check(_ffi.lib().dapol_rangeproof_verify_batch(self._h, nbits, m, k, _p(pr), pr.shape[1], _p(cm), _p(ok)))
        return ok.astype(bool)

    def imad_peak(self, variant: int) -> float:
        x = C.c_double()
        _check(_ffi.lib().dapol_imad_peak(self._h, variant, C.byref(x)))
        return x.value

    def fe_bench(self, op: int) -> float:
        x = C.c_double()
        _check(_ffi.lib().dapol_fe_bench(self._h, op, C.byref(x)))
        return x.value


@dataclass
class DapolNode:
    """src/dapol/node.rs DapolNode {v, v_blinding, com, hash}; com is the compressed commitment."""
    value: int
    blinding: bytes
    com: bytes
    hash: bytes

    def get_value(self):
        return self.value

    def get_blinding(self):
        return self.blinding

    def get_proof_node(self):
        return DapolProofNode(self.com, self.hash)


@dataclass
class DapolProofNode:
    """src/proof/node.rs DapolProofNode {com, hash}."""
    com: bytes
    hash: bytes

    def get_com(self):
        return self.com

    def get_hash(self):
        return self.hash

    def serialize(self) -> bytes:  # proof/node.rs:74-79
        return self.com + self.hash


class DapolProof:
    """src/proof/mod.rs DapolProof<D, R> in its serialised form (DapolProof::serialize, proof/mod.rs:68-73)."""

    def __init__(self, data: bytes, hash_id: int, policy: int):
        self.data, self.hash_id, self.policy = bytes(data), hash_id, policy

    def serialize(self) -> bytes:
        return self.data

    @classmethod
    def deserialize(cls, data: bytes, hash_id: int, policy: int):
        return cls(data, hash_id, policy)

    def verify(self, ctx: "Context", root: "DapolProofNode", leaf: "DapolProofNode") -> bool:
        """DapolProof::verify(&root, &leaf) (proof/mod.rs:41-47)."""
        return bool(self.verify_many(ctx, root, [leaf], [self])[0])

    def verify_batch(self, ctx: "Context", root: "DapolProofNode", leaves) -> bool:
        """DapolProof::verify_batch(&root, &leaves) (proof/mod.rs:49-54): this ONE proof covers all the leaves (index order)."""
        lc = np.frombuffer(b"".join(l.com for l in leaves), np.uint8).copy()
        lh = np.frombuffer(b"".join(l.hash for l in leaves), np.uint8).copy()
        blob = np.frombuffer(self.data or b"\0", np.uint8).copy()
        ok = np.zeros(1, np.uint8)
        _check(_ffi.lib().dapol_proof_verify_batch(ctx._h, self.hash_id, self.policy, len(leaves), _p(np.frombuffer(root.com, np.uint8).copy()),
                                                   _p(np.frombuffer(root.hash, np.uint8).copy()), _p(lc), _p(lh), _p(blob), len(self.data), _p(ok)))
        return bool(ok[0])

    @staticmethod
    def verify_many(ctx: "Context", root: "DapolProofNode", leaves, proofs) -> np.ndarray:
        """k independent DapolProof::verify calls against one root, as one GPU batch."""
        k = len(proofs)
        blob = np.frombuffer(b"".join(p.data for p in proofs) or b"\0", np.uint8).copy()
        off = np.zeros(k + 1, np.uint64)
        off[1:] = np.cumsum([len(p.data) for p in proofs], dtype=np.uint64)
        lc = np.frombuffer(b"".join(l.com for l in leaves), np.uint8).copy()
        lh = np.frombuffer(b"".join(l.hash for l in leaves), np.uint8).copy()
        ok = np.zeros(k, np.uint8)
        rc = _ffi.lib().dapol_verify_batch(ctx._h, proofs[0].hash_id, proofs[0].policy, k, _p(np.frombuffer(root.com, np.uint8).copy()),
                                           _p(np.frombuffer(root.hash, np.uint8).copy()), _p(lc), _p(lh), _p(blob), _p(off), _p(ok))
        _check(rc)
        return ok.astype(bool)


class Dapol:
    """src/dapol/mod.rs Dapol<D, R>: D = hash_id, R = policy."""

    def __init__(self, ctx: Context, hash_id: int, height: int, aggregation_factor: int, policy: int = POLICY_PADDING):
        self.ctx, self.hash_id, self.height = ctx, hash_id, height
        self.aggregation_factor, self.policy = aggregation_factor, policy
        self._t = None
        ctx._trees.add(self)

    # -- constructors -------------------------------------------------------------------------
    @classmethod
    def new_blank(cls, ctx, hash_id, height, aggregation_factor, policy=POLICY_PADDING):
        """Dapol::new_blank (mod.rs:196-204)."""
        return cls(ctx, hash_id, height, aggregation_factor, policy)

    def build(self, leaf_idx, values, blindings, pad_seed: bytes, pad_base: int = 0):
        """Dapol::build(&items, &secret) (mod.rs:206-208) with items = DapolNode::new(values[i], blindings[i])
        at TreeIndex::from_u64(height, leaf_idx[i]); padding randomness from the seeded stream."""
        idx = np.ascontiguousarray(leaf_idx, np.uint64)
        val = np.ascontiguousarray(values, np.uint64)
        bl = np.ascontiguousarray(blindings, np.uint8).reshape(-1, 32)
        if not (len(idx) == len(val) == len(bl)):
            raise DapolError(16)
        self._free()
        h = C.c_void_p()
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        _check(_ffi.lib().dapol_tree_build_from_nodes(self.ctx._h, self.hash_id, self.height, len(idx), _p(idx), _p(val), _p(bl),
                                                      seed, pad_base, C.byref(h)))
        self._t = h
        return self

    def update(self, leaf_idx, values, blindings, pad_seed: bytes, pad_base: int = 0):
        """Dapol::update(&idx, node, &secret) (mod.rs:210-213) for one leaf or a batch (strictly increasing indexes): the tree now
        holds the old leaves plus these, a leaf at an existing index being replaced.  On a blank Dapol this is build()."""
        idx = np.atleast_1d(np.ascontiguousarray(leaf_idx, np.uint64))
        val = np.atleast_1d(np.ascontiguousarray(values, np.uint64))
        bl = np.ascontiguousarray(blindings, np.uint8).reshape(-1, 32)
        if not (len(idx) == len(val) == len(bl)):
            raise DapolError(16)
        if self._t is None:
            return self.build(idx, val, bl, pad_seed, pad_base)
        h = C.c_void_p()
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        _check(_ffi.lib().dapol_tree_update(self._t, len(idx), _p(idx), _p(val), _p(bl), seed, pad_base, C.byref(h)))
        self._free()
        self._t = h
        return self

    def build_dev(self, n, d_leaf_idx: int, d_values: int, d_blindings: int, pad_seed: bytes, pad_base: int = 0):
        """Same as build() with the inputs already in device memory (raw device pointers)."""
        self._free()
        h = C.c_void_p()
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        _check(_ffi.lib().dapol_tree_build_from_nodes_dev(self.ctx._h, self.hash_id, self.height, n, d_leaf_idx, d_values,
                                                          d_blindings, seed, pad_base, C.byref(h)))
        self._t = h
        return self

    @staticmethod
    def pack_ids(ids):
        """list of bytes -> (blob u8[], offsets u64[n+1]) as the C ABI takes them."""
        off = np.zeros(len(ids) + 1, dtype=np.uint64)
        if len(ids):
            off[1:] = np.cumsum([len(x) for x in ids], dtype=np.uint64)
        blob = np.frombuffer(b"".join(ids) or b"\0", dtype=np.uint8).copy()
        return blob, off

    @classmethod
    def new(cls, ctx, hash_id, liabilities, audit_seed: bytes, tree_height: int, aggregation_factor: int, pad_seed: bytes,
            policy=POLICY_PADDING, pad_base: int = 0):
        """Dapol::new(liabilities, options) (mod.rs:100-128).  liabilities = [(internal_id, external_id, value)] or the
        packed form (iid_blob, iid_off, eid_blob, eid_off, values).  `secret` of DapolOptions is ignored by the
        reference's padding (node.rs:86); pad_seed seeds the padding RNG stream instead."""
        self = cls(ctx, hash_id, tree_height, aggregation_factor, policy)
        if isinstance(liabilities, tuple) and len(liabilities) == 5 and isinstance(liabilities[0], np.ndarray):
            ib, io, eb, eo, vals = liabilities
        else:
            ib, io = cls.pack_ids([x[0] for x in liabilities])
            eb, eo = cls.pack_ids([x[1] for x in liabilities])
            vals = np.array([x[2] for x in liabilities], np.uint64)
        ib = np.ascontiguousarray(ib, np.uint8); eb = np.ascontiguousarray(eb, np.uint8)
        io = np.ascontiguousarray(io, np.uint64); eo = np.ascontiguousarray(eo, np.uint64)
        vals = np.ascontiguousarray(vals, np.uint64)
        n = len(io) - 1
        h = C.c_void_p(); err = C.c_uint64(0)
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        aseed = (C.c_uint8 * max(len(audit_seed), 1)).from_buffer_copy(audit_seed or b"\0")
        rc = _ffi.lib().dapol_tree_build_from_liabilities(ctx._h, hash_id, tree_height, n, _p(ib), _p(io), _p(eb), _p(eo), _p(vals),
                                                          aseed, len(audit_seed), seed, pad_base, C.byref(h), C.byref(err))
        _check(rc, err.value if rc in (4, 5) else None)
        self._t = h
        self._n = n
        return self

    def leaf_index_of(self, input_pos: int):
        """id_to_idx_map lookup (mod.rs:148-151) by input position; None if the tree was not built from liabilities."""
        x = C.c_uint64()
        rc = _ffi.lib().dapol_tree_leaf_index_of(self._t, input_pos, C.byref(x))
        if rc == 17:
            return None
        _check(rc)
        return x.value

    # -- persistence (no reference counterpart: the crate keeps the tree in memory only) ------------
    def save(self, path: str):
        """Whole node store + slot maps + id -> leaf-index map to one file (holds the secret blindings)."""
        _check(_ffi.lib().dapol_tree_save(self._t, os.fsencode(path)))

    @classmethod
    def load(cls, ctx, path: str, aggregation_factor: int, policy: int = POLICY_PADDING):
        """The saved tree back in HBM on `ctx`: same levels, paths and inclusion proofs as the tree that was saved."""
        h = C.c_void_p()
        _check(_ffi.lib().dapol_tree_load(ctx._h, os.fsencode(path), C.byref(h)))
        L = _ffi.lib()
        self = cls(ctx, 0, L.dapol_tree_height(h), aggregation_factor, policy)
        self._t = h
        self.hash_id = int(L.dapol_tree_hash_id(h))
        return self

    def generate_proofs_to_file(self, leaf_idx, seed: bytes, path: str, chunk: int = 0) -> int:
        """Dapol::generate_proof for every index, streamed to `path` (mod.rs:250 TODO); returns the size of one proof."""
        li = np.ascontiguousarray(leaf_idx, np.uint64)
        size = C.c_uint64()
        sd = (C.c_uint8 * 32).from_buffer_copy(seed)
        _check(_ffi.lib().dapol_prove_to_file(self._t, len(li), _p(li), self.aggregation_factor, self.policy, sd, chunk, os.fsencode(path),
                                              C.byref(size)))
        return size.value

    def _free(self):
        if getattr(self, "_t", None):
            if getattr(self.ctx, "_h", None):  # a tree that outlived its context was already released with it
                _ffi.lib().dapol_tree_destroy(self._t)
            self._t = None

    close = _free
    __del__ = _free

    # -- accessors ----------------------------------------------------------------------------
    def root_raw(self) -> DapolNode:
        """Dapol::root_raw (mod.rs:134-136)."""
        com = np.zeros(32, np.uint8); hs = np.zeros(digest_len(self.hash_id), np.uint8); bl = np.zeros(32, np.uint8)
        v = C.c_uint64()
        _check(_ffi.lib().dapol_tree_root(self._t, _p(com), _p(hs), C.byref(v), _p(bl)))
        return DapolNode(v.value, bl.tobytes(), com.tobytes(), hs.tobytes())

    def root(self) -> DapolProofNode:
        """Dapol::root (mod.rs:139-141)."""
        return self.root_raw().get_proof_node()

    @property
    def num_nodes(self):
        return _ffi.lib().dapol_tree_num_nodes(self._t)

    @property
    def num_padding(self):
        return _ffi.lib().dapol_tree_num_padding(self._t)

    def level(self, h: int):
        n = _ffi.lib().dapol_tree_level_size(self._t, h)
        idx = np.zeros(n, np.uint64); v = np.zeros(n, np.uint64)
        r = np.zeros((n, 32), np.uint8); c = np.zeros((n, 32), np.uint8); hs = np.zeros((n, digest_len(self.hash_id)), np.uint8)
        pad = np.zeros(n, np.uint8)
        _check(_ffi.lib().dapol_tree_level_copy(self._t, h, _p(idx), _p(v), _p(r), _p(c), _p(hs), _p(pad)))
        return dict(idx=idx, v=v, r=r, comc=c, hash=hs, is_pad=pad)

    def paths(self, leaf_idx):
        """Siblings (leaf level first) of each requested leaf + the leaf proof nodes; None if any is absent."""
        li = np.ascontiguousarray(leaf_idx, np.uint64)
        k, H = len(li), max(self.height, 1)
        v = np.zeros((k, H), np.uint64)
        dl = digest_len(self.hash_id)
        r = np.zeros((k, H, 32), np.uint8); c = np.zeros((k, H, 32), np.uint8); hs = np.zeros((k, H, dl), np.uint8)
        lc = np.zeros((k, 32), np.uint8); lh = np.zeros((k, dl), np.uint8)
        rc = _ffi.lib().dapol_tree_paths(self._t, k, _p(li), _p(v), _p(r), _p(c), _p(hs), _p(lc), _p(lh))
        if rc == 17:
            return None
        _check(rc)
        return dict(v=v, r=r, comc=c, hash=hs, leaf_comc=lc, leaf_hash=lh)

    # -- inclusion proofs -----------------------------------------------------------------------
    def generate_proofs(self, leaf_idx, seed: bytes):
        """[Dapol::generate_proof(idx) for idx in leaf_idx] (mod.rs:167-190) as one GPU batch; None if any index is not a
        leaf (the reference returns None).  Prover randomness comes from the seeded stream (include/dapol_b200.h)."""
        li = np.ascontiguousarray(leaf_idx, np.uint64)
        k = len(li)
        L = _ffi.lib()
        size = L.dapol_inclusion_proof_size_d(self.height, self.aggregation_factor, self.policy, self.hash_id)
        if size == 0:
            raise DapolError(16)
        out = np.zeros(k * size, np.uint8)
        got = C.c_uint64()
        sd = (C.c_uint8 * 32).from_buffer_copy(seed)
        rc = L.dapol_prove_batch(self._t, k, _p(li), self.aggregation_factor, self.policy, sd, _p(out), out.nbytes, C.byref(got))
        if rc == 17:
            return None
        _check(rc)
        return [DapolProof(out[i * size:(i + 1) * size].tobytes(), self.hash_id, self.policy) for i in range(k)]

    def generate_proof(self, leaf_idx: int, seed: bytes):
        """Dapol::generate_proof (mod.rs:167-169)."""
        r = self.generate_proofs([leaf_idx], seed)
        return None if r is None else r[0]

    def index_of_ids(self, internal_ids):
        """id_to_idx_map lookups (mod.rs:148-165) by the internal ids themselves; None if any id is unknown."""
        blob, off = self.pack_ids(list(internal_ids))
        k = len(off) - 1
        idx = np.zeros(k, np.uint64); found = np.zeros(k, np.uint8)
        rc = _ffi.lib().dapol_tree_index_of_batch(self._t, k, _p(blob), _p(off), _p(idx), _p(found))
        if rc == 17:
            return None
        _check(rc)
        return [int(x) for x in idx]

    def generate_proof_for_id(self, internal_id: bytes, seed: bytes):
        """Dapol::generate_proof_for_id (mod.rs:148-151)."""
        idx = self.index_of_ids([internal_id])
        return None if idx is None else self.generate_proof(idx[0], seed)

    def generate_proof_batch_for_ids(self, internal_ids, seed: bytes):
        """Dapol::generate_proof_batch_for_ids (mod.rs:155-165): one batch proof for the leaves of the given ids, in the ids' order
        (the reference passes the indexes on as they come; smtree wants them increasing)."""
        idx = self.index_of_ids(internal_ids)
        return None if idx is None else self.generate_proof_batch(idx, seed)

    def generate_proof_batch(self, leaf_idx, seed: bytes):
        """Dapol::generate_proof_batch(&[TreeIndex]) (mod.rs:172-190): ONE DapolProof for all the given leaves (strictly
        increasing indexes); None if any index is not a leaf."""
        li = np.ascontiguousarray(leaf_idx, np.uint64)
        L = _ffi.lib()
        size = L.dapol_batch_proof_size_d(self.height, len(li), _p(li), self.aggregation_factor, self.policy, self.hash_id)
        if size == 0:
            raise DapolError(16)
        out = np.zeros(size, np.uint8)
        got = C.c_uint64()
        sd = (C.c_uint8 * 32).from_buffer_copy(seed)
        rc = L.dapol_generate_proof_batch(self._t, len(li), _p(li), self.aggregation_factor, self.policy, sd, _p(out), out.nbytes, C.byref(got))
        if rc == 17:
            return None
        _check(rc)
        return DapolProof(out.tobytes(), self.hash_id, self.policy)
