"""ctypes binding of include/dapol_b200.h.  Fails loudly if the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DAPOL_B200_LIB") or os.path.join(_HERE, "lib", "libdapol_b200.so")  # override: kernel-variant experiments
_lib = None

u64 = C.c_uint64
vp = C.c_void_p


class LibraryMissing(RuntimeError):
    pass


# dapol_comm_ops (include/dapol_b200.h): a host-provided transport for the sharded build's collectives
ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, vp, vp, vp, u64, vp)
ALL_TO_ALL_FN = C.CFUNCTYPE(C.c_int, vp, vp, C.POINTER(u64), C.POINTER(u64), vp, C.POINTER(u64), C.POINTER(u64), vp)


class CommOps(C.Structure):
    _fields_ = [("user", vp), ("all_gather", ALL_GATHER_FN), ("all_to_all", ALL_TO_ALL_FN)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} not built: run ./build.sh (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.dapol_strerror.restype = C.c_char_p
    L.dapol_last_cuda_error.restype = C.c_char_p
    L.dapol_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.dapol_ctx_destroy.argtypes = [vp]
    L.dapol_ctx_set_stream.argtypes = [vp, vp]
    L.dapol_tree_build_from_liabilities_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp, u64, vp, u64,
                                                        C.POINTER(vp), C.POINTER(u64)]
    L.dapol_tree_destroy.argtypes = [vp]
    L.dapol_tree_build_from_nodes.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, u64, C.POINTER(vp)]
    L.dapol_tree_build_from_nodes_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, u64, C.POINTER(vp)]
    L.dapol_tree_update.argtypes = [vp, u64, vp, vp, vp, vp, u64, C.POINTER(vp)]
    L.dapol_tree_build_from_liabilities.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp, u64, vp, u64,
                                                    C.POINTER(vp), C.POINTER(u64)]
    L.dapol_leaves_derive_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, u64, vp, vp, vp, vp]
    L.dapol_leaves_assign_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, C.c_int, u64, vp, vp, vp, u64,
                                          C.POINTER(u64), C.POINTER(u64)]
    L.dapol_tree_level_pad_counts_dev.argtypes = [vp, C.c_int, u64, vp, vp]
    L.dapol_tree_build_shard_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.dapol_tree_root_record.argtypes = [vp, vp]
    L.dapol_tree_build_from_records.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, u64, C.POINTER(vp)]
    L.dapol_tree_attach_top.argtypes = [vp, vp, u64]
    L.dapol_tree_root.argtypes = [vp, vp, vp, C.POINTER(u64), vp]
    L.dapol_ctx_set_padding_mode.argtypes = [vp, C.c_int]
    L.dapol_tree_height.argtypes = [vp]
    L.dapol_tree_hash_id.argtypes = [vp]
    L.dapol_tree_num_nodes.argtypes = [vp]
    L.dapol_tree_num_nodes.restype = u64
    L.dapol_tree_num_padding.argtypes = [vp]
    L.dapol_tree_num_padding.restype = u64
    L.dapol_tree_level_size.argtypes = [vp, C.c_int]
    L.dapol_tree_level_size.restype = u64
    L.dapol_tree_level_copy.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.dapol_tree_leaf_index_of.argtypes = [vp, u64, C.POINTER(u64)]
    L.dapol_tree_paths.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp]
    L.dapol_commit_batch.argtypes = [vp, u64, vp, vp, vp]
    L.dapol_imad_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.dapol_fe_bench.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.dapol_kernel_launches.argtypes = [vp]
    L.dapol_kernel_launches.restype = u64
    L.dapol_last_build_times.argtypes = [vp, vp]
    L.dapol_ctx_params.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dapol_inclusion_proof_size.argtypes = [C.c_int, u64, C.c_int]
    L.dapol_inclusion_proof_size.restype = u64
    L.dapol_inclusion_proof_size_d.argtypes = [C.c_int, u64, C.c_int, C.c_int]
    L.dapol_inclusion_proof_size_d.restype = u64
    L.dapol_batch_proof_size_d.argtypes = [C.c_int, u64, vp, u64, C.c_int, C.c_int]
    L.dapol_batch_proof_size_d.restype = u64
    L.dapol_digest_len.argtypes = [C.c_int]
    L.dapol_prove_batch.argtypes = [vp, u64, vp, u64, C.c_int, vp, vp, u64, C.POINTER(u64)]
    L.dapol_verify_batch.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp, vp]
    L.dapol_tree_save.argtypes = [vp, C.c_char_p]
    L.dapol_tree_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.dapol_prove_to_file.argtypes = [vp, u64, vp, u64, C.c_int, vp, u64, C.c_char_p, C.POINTER(u64)]
    L.dapol_rangeproof_size.argtypes = [C.c_int, C.c_int]
    L.dapol_rangeproof_size.restype = u64
    L.dapol_rangeproof_prove_batch.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp]
    L.dapol_rangeproof_prove_batch_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp]
    L.dapol_rangeproof_verify_batch.argtypes = [vp, C.c_int, C.c_int, u64, vp, u64, vp, vp]
    L.dapol_rangeproof_verify_batch_dev.argtypes = [vp, C.c_int, C.c_int, u64, vp, u64, vp, vp]
    L.dapol_ctx_set_verify_mode.argtypes = [vp, u64, C.c_int, vp]
    L.dapol_ctx_verify_fallbacks.argtypes = [vp]
    L.dapol_ctx_verify_fallbacks.restype = u64
    L.dapol_ctx_set_rangeproof_window.argtypes = [vp, C.c_int]
    L.dapol_rangeproof_last_times.argtypes = [vp, vp]
    L.dapol_rangeproof_last_kernel_times.argtypes = [vp, vp]
    L.dapol_ctx_set_leaf_hash_mode.argtypes = [vp, C.c_int]
    L.dapol_ctx_set_rangeproof_table_budget.argtypes = [vp, u64]
    L.dapol_ctx_rangeproof_table_bytes.argtypes = [vp]
    L.dapol_ctx_rangeproof_table_bytes.restype = u64
    L.dapol_tree_index_of.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.dapol_tree_index_of_batch.argtypes = [vp, u64, vp, vp, vp, vp]
    L.dapol_batch_proof_size.argtypes = [C.c_int, u64, vp, u64, C.c_int]
    L.dapol_batch_proof_size.restype = u64
    L.dapol_generate_proof_batch.argtypes = [vp, u64, vp, u64, C.c_int, vp, vp, u64, C.POINTER(u64)]
    L.dapol_proof_verify_batch.argtypes = [vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, u64, vp]
    L.dapol_comm_create.argtypes = [C.c_int, C.c_int, C.POINTER(CommOps), C.POINTER(vp)]
    L.dapol_comm_nccl_unique_id.argtypes = [vp]
    L.dapol_comm_nccl_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.dapol_comm_destroy.argtypes = [vp]
    L.dapol_comm_rank.argtypes = [vp]
    L.dapol_comm_world.argtypes = [vp]
    L.dapol_sharded_build.argtypes = [vp, vp, C.c_int, C.c_int, u64, vp, vp, vp, vp, vp, vp, u64, vp, u64, C.POINTER(vp), C.POINTER(vp),
                                      C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), vp]
    _lib = L
    return L


def header_symbols():
    """Function names declared in include/dapol_b200.h (used by the CPU-side export test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "dapol_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dapol_[a-z0-9_]+)\s*\(", txt)))
