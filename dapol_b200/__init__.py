"""dapol_b200: B200-native (sm_100a CUDA) engine for the DAPOL+ hot path of MystenLabs/dapol.

The compute lives in dapol_b200/lib/libdapol_b200.so (C ABI: include/dapol_b200.h); this package is
the host-side mirror of the reference crate's public surface (src/lib.rs:1-14).  There is no CPU
fallback: importing works anywhere, but every compute call needs the CUDA library and a GPU.
"""
from .api import (Dapol, DapolError, DapolNode, DapolProof, DapolProofNode, HASH_BLAKE2B, HASH_BLAKE2S, HASH_BLAKE3, POLICY_PADDING, POLICY_SPLITTING,
                  Context)

from .sharded import Comm, CudaEngine, NativeComm, ShardedDapol, shard_pad_bases

__all__ = ["Comm", "CudaEngine", "NativeComm", "ShardedDapol", "shard_pad_bases", "Dapol", "DapolError", "DapolNode", "DapolProof", "DapolProofNode", "Context", "HASH_BLAKE3", "HASH_BLAKE2S", "HASH_BLAKE2B",
           "POLICY_PADDING", "POLICY_SPLITTING"]
