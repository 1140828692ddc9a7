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
