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
