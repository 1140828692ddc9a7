"""CPU-side check: the in-tree CUDA library loads and exports every symbol include/dapol_b200.h declares.
No compute call is made (no GPU here)."""
import ctypes
import os

import pytest

from dapol_b200 import _ffi


def test_library_is_built_in_tree():
    assert os.path.exists(_ffi.LIB_PATH), "run ./build.sh"


def test_exports_match_header():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    syms = _ffi.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_strerror_and_no_device_fails_loudly():
    L = _ffi.lib()
    assert L.dapol_strerror(4).decode().startswith("liability set contains a duplicated")
    import torch
    if not torch.cuda.is_available():
        from dapol_b200 import Context, DapolError
        with pytest.raises(DapolError) as e:
            Context(0)
        assert e.value.code == 19  # DAPOL_ERR_CUDA: no silent CPU fallback


def test_proof_sizes_match_the_oracle_layout():
    """Host-only entry points (no device needed): sizes of a range proof (src/range/mod.rs:18: 672 bytes for n = 64, m = 1) and
    of a serialised DapolProof = R::serialize || MerkleProof::serialize (src/proof/mod.rs:68-73) for both policies, against
    the oracle's wire-format restatement (framing widths src/range/mod.rs:19-21)."""
    from oracle import pyref
    L = _ffi.lib()
    L.dapol_rangeproof_size.restype = ctypes.c_uint64
    L.dapol_inclusion_proof_size.restype = ctypes.c_uint64
    assert L.dapol_rangeproof_size(64, 1) == pyref.SINGLE_PROOF_BYTE_NUM == 672
    for n, m in [(8, 1), (16, 2), (32, 4), (64, 16), (64, 32), (64, 64)]:
        lg = (n * m).bit_length() - 1
        assert L.dapol_rangeproof_size(n, m) == 32 * (9 + 2 * lg)
    assert L.dapol_rangeproof_size(64, 3) == 0 and L.dapol_rangeproof_size(12, 1) == 0  # not a power of two / unsupported width
    for H in (1, 4, 9, 16, 32, 40, 64):
        for agg in sorted({0, 1, 2, 3, H // 2, H - 1, H}):
            if agg < 0 or agg > H:
                continue
            for policy in (0, 1):
                groups, singles = pyref.policy_plan(H, agg, policy)
                rng = (0 if policy == 0 else 2) + sum(8 + 32 * (9 + 2 * ((64 * m).bit_length() - 1)) for _, _, m in groups) + 8 + 672 * len(singles)
                merkle = len(pyref.merkle_serialize(H, 0, [(bytes(32), bytes(32))] * H))
                assert L.dapol_inclusion_proof_size(H, agg, policy) == rng + merkle, (H, agg, policy)
        assert L.dapol_inclusion_proof_size(H, H + 1, 0) == 0  # the reference panics (slice out of bounds)


def test_batch_and_digest_aware_sizes_match_the_oracle():
    """Host-only: the sibling plan of a batch proof (Dapol::generate_proof_batch, mod.rs:172-190: a sibling is carried only if it is not
    itself on the way up from the batch) and the digest-aware wire sizes (a sibling is com || hash of Dlen bytes, proof/node.rs:74-79)
    of the library's C++ against both oracle restatements, for random batches incl. adjacent leaves, one leaf and a whole level."""
    import random

    import numpy as np

    from oracle import cref, pyref
    L = _ffi.lib()
    O = cref.lib()
    O.dor_batch_proof_size_d.restype = ctypes.c_uint64
    O.dor_inclusion_proof_size_d.restype = ctypes.c_uint64
    assert [L.dapol_digest_len(h) for h in (0, 1, 2, 3, -1)] == [32, 32, 64, 0, 0]
    rnd = random.Random(11)
    for H in (1, 3, 8, 10, 20, 40, 64):
        for hash_id, dl in ((0, 32), (2, 64)):
            for agg, policy in ((0, 0), (1, 1), (H, 0), (H // 2, 1)):
                want = O.dor_inclusion_proof_size_d(H, ctypes.c_uint64(agg), policy, dl)
                assert L.dapol_inclusion_proof_size_d(H, agg, policy, hash_id) == want > 0
                assert want == L.dapol_inclusion_proof_size(H, agg, policy) + (dl - 32) * H
        assert L.dapol_inclusion_proof_size_d(H, 1, 0, 7) == 0  # unknown digest
        for k in (1, 2, 3, 10, min(64, 1 << H)):
            k = min(k, 1 << min(H, 20))
            picks = set()
            while len(picks) < k:
                picks.add(rnd.randrange(1 << H) if rnd.random() < 0.5 or not picks else min((1 << H) - 1, max(picks) ^ 1))
            idx = np.array(sorted(picks), np.uint64)
            plan = pyref.batch_sibling_plan(H, [int(x) for x in idx])
            for agg, policy in ((0, 0), (1, 1), (len(plan), 0), (len(plan) // 2, 1)):
                for hash_id, dl in ((0, 32), (2, 64)):
                    got = L.dapol_batch_proof_size_d(H, len(idx), idx.ctypes.data_as(ctypes.c_void_p), agg, policy, hash_id)
                    want = O.dor_batch_proof_size_d(H, ctypes.c_uint64(len(idx)), idx.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(agg), policy, dl)
                    assert got == want, (H, k, agg, policy, hash_id)
                    if len(idx) > 1:  # the Merkle part alone (framing + one (32 + Dlen)-byte entry per planned sibling) is smaller than the whole
                        nb = (H + 7) // 8
                        assert got == 0 or (got - (2 + 8 + len(idx) * nb + 8 + (32 + dl) * len(plan))) > 0
            assert L.dapol_batch_proof_size_d(H, len(idx), idx.ctypes.data_as(ctypes.c_void_p), len(plan) + 1, 0, 0) == 0 or len(idx) == 1  # slice out of bounds in the reference
        if H >= 3:
            bad = np.array([5, 2], np.uint64)  # not increasing: smtree rejects
            assert L.dapol_batch_proof_size_d(H, 2, bad.ctypes.data_as(ctypes.c_void_p), 0, 0, 0) == 0


def test_python_mirror_covers_the_reference_surface():
    """The host-side mirror exposes what the crate's root re-exports (src/lib.rs:1-14) and the methods the hot path uses:
    Dapol::{new, new_blank, build, update, root, root_raw, generate_proof[_for_id], generate_proof_batch[_for_ids]} (src/dapol/mod.rs:100-213),
    DapolNode::{get_value, get_blinding} (node.rs:48-56), DapolProofNode::{get_com, get_hash} (proof/node.rs:30-49),
    DapolProof::{serialize, deserialize, verify, verify_batch} (proof/mod.rs:41-84), both policies, the three digests."""
    import dapol_b200 as d
    for name in ("Dapol", "DapolNode", "DapolProof", "DapolProofNode", "DapolError", "Context", "ShardedDapol", "POLICY_PADDING", "POLICY_SPLITTING",
                 "HASH_BLAKE3", "HASH_BLAKE2S", "HASH_BLAKE2B"):
        assert hasattr(d, name), name
    for m in ("new", "new_blank", "build", "update", "root", "root_raw", "generate_proof", "generate_proof_for_id", "generate_proof_batch",
              "generate_proof_batch_for_ids", "leaf_index_of", "save", "load"):
        assert callable(getattr(d.Dapol, m)), m
    for m in ("serialize", "deserialize", "verify", "verify_batch", "verify_many"):
        assert callable(getattr(d.DapolProof, m)), m
    n = d.DapolNode(5, bytes(32), bytes(32), bytes(64))
    assert (n.get_value(), n.get_blinding(), n.get_proof_node().serialize()) == (5, bytes(32), bytes(96))  # com || hash (proof/node.rs:74-79), 64-byte digest
    assert (d.POLICY_PADDING, d.POLICY_SPLITTING, d.HASH_BLAKE3, d.HASH_BLAKE2S, d.HASH_BLAKE2B) == (0, 1, 0, 1, 2)
