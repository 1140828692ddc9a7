// Internals shared by the translation units of libdapol_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <string>
#include "../../include/dapol_b200.h"
#include <vector>
#include "tree_kernels.cuh"

std::string &dapol_cuda_err();  // thread-local last CUDA error text (dapol_last_cuda_error)
#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            dapol_cuda_err() = std::string(#expr) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__); \
            return DAPOL_ERR_CUDA;                                                                           \
        }                                                                                                    \
    } while (0)

// Device memory on the hot path comes from the stream-ordered pool (cudaMallocAsync) with the release
// threshold lifted, so repeated calls reuse the same HBM without paying cudaMalloc/cudaFree each time.
static inline cudaError_t dmalloc(void **p, size_t bytes, cudaStream_t st) { return cudaMallocAsync(p, bytes ? bytes : 1, st); }
template <typename T>
static inline cudaError_t dmalloc(T **p, size_t bytes, cudaStream_t st) { return dmalloc(reinterpret_cast<void **>(p), bytes, st); }
static inline void dfree(void *p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }
static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// comb windows of the tree tables that are instantiated (dapol_ctx_create's comb_window; 0 = default)
#ifndef DAPOL_DEFAULT_COMB_WINDOW
#define DAPOL_DEFAULT_COMB_WINDOW 15
#endif
// comb_window 0 picks the wide HBM-resident window when at least this much device memory is free, else the default
#ifndef DAPOL_WIDE_COMB_WINDOW
#define DAPOL_WIDE_COMB_WINDOW 24
#endif
#ifndef DAPOL_WIDE_COMB_MIN_FREE_GB
#define DAPOL_WIDE_COMB_MIN_FREE_GB 48
#endif
#ifndef DAPOL_W_CASES
#define DAPOL_W_CASES(X) X(4) X(8) X(12) X(15) X(16) X(20) X(22) X(24) X(26)
#endif

struct dapol_ctx {
    int device = 0;
    int W = 8;  // comb window of the tree tables
    int pad_mode = 0;  // DAPOL_PADDING_STREAM / DAPOL_PADDING_POSITIONAL (dapol_ctx_set_padding_mode)
    int leaf_hash_mode = 0;  // DAPOL_LEAF_HASH_COMMITMENT / DAPOL_LEAF_HASH_ID_SALT (dapol_ctx_set_leaf_hash_mode)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    ge_niels *tab_b = nullptr, *tab_bbl = nullptr;  // comb tables for B and B_blinding at window W (253/W+1 windows each)
    uint64_t launches = 0;
    float last_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[6] = {};
    unsigned long long *scratch = nullptr;  // 1 KB of device scratch (histograms, counters)
    // range proofs: window tables of the Bulletproof generators G_j[i], H_j[i] (j < rp_mcap, i < 64) and of B, B_blinding
    int rp_W = 12;
    bool rp_W_auto = true;  // pick the widest window whose tables fit the HBM budget when they are built
    int rp_mcap = 0;
    ge_niels *rp_tab = nullptr;  // [(128 mcap + 2)][NW][2^(W-1)]
    size_t rp_budget = 0;        // HBM budget of those tables in bytes (0 = 70 % of the free memory, at most 128 GB)
    size_t rp_table_bytes = 0;   // size of the tables in place
    // last range-proof batch: [0] total, [1] MSM passes, [2] other passes, [3] table build, and the MSM passes split per
    // kernel class: [4] k_rp_p10 (L / R of the table rounds), [5] k_rp_p3 (A, S), [6] hybrid rounds (pm, pv, pf), [7] verifier (v1, v2)
    float rp_last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // verifier mode (dapol_ctx_set_verify_mode): 0 / 1 = every proof on its own (Straus); G > 1 = groups of G proofs checked by one
    // random linear combination with the bucket method, failed groups re-verified per proof
    int rp_verify_group = 0, rp_verify_window = 0;
    bool rp_verify_seeded = false;
    uint32_t rp_verify_seed[8] = {};
    uint64_t rp_verify_redone = 0;
    int rp_pack_min_k = 8192;  // ... from this many proofs per batch on (DAPOL_RP_PACK_MIN_K)
    int rp_pack_max_n = 32;    // ... for shapes of at most this many generators per vector (DAPOL_RP_PACK_MAX_N)
    int rp_pack_lanes = 8;  // lanes per MSM of the packed prover kernels for small shapes in large batches (0: a warp per MSM); DAPOL_RP_PACK_LANES  // proofs re-verified one by one so far (their group's combination failed)
};

struct dapol_tree {
    dapol_ctx *ctx = nullptr;
    int hash_id = 0, height = 0;
    uint64_t n_leaves = 0, T = 0, n_pads = 0;
    std::vector<uint64_t> level_off, level_n, n_real;  // per level h = 0..H
    NodeStore ns = {};
    std::vector<uint32_t *> pos;  // pos[h]: slot of the k-th real node of level h (device), h = 1..H
    uint32_t *pos_all = nullptr;  // backing allocation of pos[]
    uint32_t **d_pos = nullptr;   // device copy of the pointer table
    uint64_t *d_level_off = nullptr;
    uint64_t *leaf_index_of = nullptr;  // device [n]: leaf idx of the i-th input liability (from_liabilities only)
    bool custom_leaf_hashes = false;  // leaf hashes are the id / salt ones (DAPOL_LEAF_HASH_ID_SALT), not D(compress(com))
    uint64_t index_map_first = 0, index_map_n = 0;  // sharded build: the map covers input positions [first, first + n) (0: all n_leaves)
    // lookup by internal id (Dapol::generate_proof_for_id, mod.rs:148-165): audit ids of the mapped liabilities, the audit seed, and
    // -- built on the first lookup -- the ids' 64-bit prefixes sorted with the input position of each
    uint32_t *audit_ids = nullptr;
    std::vector<uint8_t> audit_seed;
    mutable uint64_t *akey_sorted = nullptr;
    mutable uint32_t *akey_who = nullptr;
    mutable std::mutex index_mu;
    std::vector<uint64_t> npads;        // padding nodes per level
    uint32_t root_ext[32] = {};         // half point of the root commitment (kept for the shard root record)
    uint32_t root_comc[8] = {}, root_hash[16] = {};  // compressed commitment and hash (32 or 64 bytes) of the root (host copy: prover nonce key)
    // sharded trees: this tree is the subtree under node `prefix` of level top->height of `top`
    const dapol_tree *top = nullptr;
    uint64_t prefix = 0;
};
static inline int dapol_total_height(const dapol_tree *t) { return t->height + (t->top ? t->top->height : 0); }

// units per thread of the node passes = batch size of the shared inversion (ge_dc_batch)
#ifndef NODE_BATCH
#define NODE_BATCH 24
#endif
// minimum resident 128-thread CTAs per SM the node kernels are compiled for (register cap = 65536 / (128 * MINB))
#ifndef DAPOL_LEAF_MINB
#define DAPOL_LEAF_MINB 4
#endif
#ifndef DAPOL_PAD_MINB
#define DAPOL_PAD_MINB 4
#endif
#ifndef DAPOL_MERGE_MINB
#define DAPOL_MERGE_MINB 4
#endif

// Threads for n units at 1..NODE_BATCH units each.  A thread's cost is per * unit_cost + inv_cost (the shared field
// inversion of its batch; costs in thousands of MAC32); the grid runs in waves of 148 SMs x resident CTAs, so the batch
// size is chosen to minimise waves x thread cost -- small levels keep one unit per thread and spread over the SMs, and a
// level of a few waves does not end on a nearly empty one.
template <typename K>
static inline uint64_t batch_stride(uint64_t n, K kernel, double unit_cost, size_t smem = 0, double inv_cost = 12.0) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    int resident = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, 128, smem) != cudaSuccess || resident <= 0) resident = 3;
    const double wave = (double)sms * resident;
    double best = 1e300;
    uint64_t best_threads = n;
    for (int per = 1; per <= NODE_BATCH; per++) {
        uint64_t threads = (n + per - 1) / per, ctas = (threads + 127) / 128;
        double waves = ctas / wave;
        if (waves < 6.0) waves = ceil(waves);  // few waves: the last one costs a full wave
        double cost = waves * (per * unit_cost + inv_cost);
        if (cost < best) { best = cost; best_threads = threads; }
    }
    return best_threads;
}

// the three merge steps of a built structure (dapol_merge.cu: compiled with inlined field products)
void dapol_launch_merges(dapol_ctx *ctx, dapol_tree *t);

// bump allocator over one device allocation (256-byte aligned pieces)
struct Arena {
    uint8_t *base = nullptr;
    size_t size = 0, used = 0;
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
    static size_t need(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }
};

// device-resident batch range proofs (dapol_rp.cu), shared with the inclusion-proof path
int dapol_rp_prove_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint64_t *d_values, const uint8_t *d_blind, const uint8_t seed[32],
                       const uint64_t *d_stream, const uint64_t *d_base, uint8_t *d_proofs);
int dapol_rp_verify_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint8_t *d_proofs, const uint8_t *d_coms, uint8_t *d_ok);

// device-resident Merkle paths of k leaves (dapol_lib.cu): siblings leaf level first, [k][total height] each; leaf
// indexes are those of the whole tree (prefix included when the tree is a shard with its top tree attached)
int dapol_tree_paths_dev(const dapol_tree *t, uint64_t k, const uint64_t *d_leaf_idx, uint64_t *d_v, uint32_t *d_r, uint32_t *d_c, uint32_t *d_h,
                         uint32_t *d_lc, uint32_t *d_lh, int *d_not_found, uint32_t *d_hh = nullptr, uint32_t *d_lhh = nullptr);
static inline int dapol_dlen(int hash_id) { return hash_id == DAPOL_HASH_BLAKE2B ? 64 : 32; }
