// dapol_b200 C-ABI library: CUDA kernels (sm_100a) + host orchestration for the DAPOL+ hot path.
// Interface: include/dapol_b200.h.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "dapol_internal.h"
#include "tree_kernels.cuh"

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_cuda_err;
std::string &dapol_cuda_err() { return g_cuda_err; }

extern "C" const char *dapol_last_cuda_error(void) { return g_cuda_err.c_str(); }
extern "C" const char *dapol_strerror(int code) {
    switch (code) {
        case DAPOL_OK: return "ok";
        case DAPOL_ERR_TREE_HEIGHT_TOO_BIG: return "DAPOL tree height must not exceed 64";
        case DAPOL_ERR_SPARSITY_TOO_SMALL: return "tree height too small for the liability set (2^height < 2*N)";
        case DAPOL_ERR_INVALID_DIGEST_SIZE: return "expected digest size to be 32";
        case DAPOL_ERR_DUPLICATED_INTERNAL_ID: return "liability set contains a duplicated internal ID";
        case DAPOL_ERR_FAILED_TO_MAP_INDEX: return "failed to map audit ID to a tree index within 128 tries";
        case DAPOL_ERR_BAD_ARG: return "bad argument";
        case DAPOL_ERR_NOT_FOUND: return "not found";
        case DAPOL_ERR_BUFFER: return "buffer too small";
        case DAPOL_ERR_CUDA: return "CUDA error";
        case DAPOL_ERR_DECODE: return "decoding error";
        case DAPOL_ERR_IO: return "file cannot be opened, read or written, or is not a tree file";
    }
    return "unknown";
}


// ------------------------------------------------------------------------------------------------ kernels
template <int W>
__global__ void k_comb_table(ge_niels *table, int nw, int which, uint64_t total) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) comb_table_body<W>(t, table, nw, which);
}
template <int W>
__global__ void k_comb_bases(ge *bases, int nw, int which) {
    if (blockIdx.x == 0 && threadIdx.x == 0) comb_bases_body<W>(bases, nw, which);
}
#define COMB_RUN 32
template <int W>
__global__ void __launch_bounds__(64) k_comb_table_run(ge_niels *table, int nw, const ge *bases, uint64_t total_runs) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total_runs) comb_table_run_body<W, COMB_RUN>(t, table, nw, bases);
}
// adjacent-leaf msb histogram (hist[0..63]) + input validation (hist[64] = bad flag)
__global__ void k_leaf_msb_hist(const uint64_t *idx, uint64_t n, int height, unsigned long long *hist) {
    __shared__ unsigned int sh[65];
    for (int i = threadIdx.x; i < 65; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        int bad = 0;
        int m = leaf_pair_msb(k, idx, height, &bad);
        if (bad) atomicAdd(&sh[64], 1u);
        if (m >= 0) atomicAdd(&sh[m], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 65; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}
// exclusive prefix sum of u64, three passes over tiles of SCAN_TILE items
#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t x, uint64_t *total) {
    __shared__ uint64_t warp_sums[SCAN_BLOCK / 32];
    __shared__ uint64_t block_total;
    unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < SCAN_BLOCK / 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (unsigned)d) wi += y;
        }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_BLOCK / 32 - 1) block_total = wi;
    }
    __syncthreads();
    if (total) *total = block_total;
    uint64_t r = incl - x + warp_sums[wid];
    __syncthreads();
    return r;
}
// pass 1: flags of a level + per-tile sums
__global__ void k_struct_flags_tiles(const uint64_t *idx, uint64_t c, uint64_t *flags, uint64_t *tile_sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < c) { uint64_t f = struct_flags_body(base + i, idx, c); flags[base + i] = f; s += f; }
    uint64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void k_scan_tile_offsets(uint64_t *tile_sums, uint64_t ntiles) {  // one block; in place -> exclusive
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < ntiles; base += SCAN_BLOCK) {
        uint64_t i = base + threadIdx.x;
        uint64_t x = i < ntiles ? tile_sums[i] : 0, total;
        uint64_t e = block_exclusive_scan(x, &total);
        if (i < ntiles) tile_sums[i] = e + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}
// pass 3: exclusive scan inside the tile + slots / parents / padding destinations
__global__ void k_struct_apply(const uint64_t *idx, uint64_t c, const uint64_t *flags, const uint64_t *tile_offsets, uint32_t *pos,
                               uint64_t *parent_idx, uint64_t level_off, NodeStore ns, uint64_t *pad_dest, uint64_t pad_ord_base,
                               uint64_t *pad_rng, uint64_t pad_rng_base, int positional) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t x[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { x[i] = base + i < c ? flags[base + i] : 0; s += x[i]; }
    uint64_t e = block_exclusive_scan(s, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < c) struct_apply_body(base + i, idx, x[i], e, pos, parent_idx, level_off, ns, pad_dest, pad_ord_base, pad_rng, pad_rng_base, positional);
        e += x[i];
    }
}

// ---- leaf derivation (K1)
__global__ void k_derive(uint64_t n, int hash_id, const uint8_t *iid_blob, const uint64_t *iid_off, const uint8_t *eid_blob,
                         const uint64_t *eid_off, const uint8_t *audit_seed, uint32_t seed_len, int height, uint32_t *audit,
                         uint32_t *cur_seed, uint64_t *cand, uint32_t *blind, uint32_t *tries, uint64_t *audit_key, uint32_t *iota,
                         int *too_long) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (derive_body(i, hash_id, iid_blob, iid_off, eid_blob, eid_off, audit_seed, seed_len, height, audit, cur_seed, cand, blind)) *too_long = 1;
    if (tries) {
        tries[i] = 1;
        audit_key[i] = (uint64_t)audit[8 * i] | ((uint64_t)audit[8 * i + 1] << 32);
        iota[i] = (uint32_t)i;
    }
}
// stage-2 state for records that were derived elsewhere (sharded build)
__global__ void k_assign_init(uint64_t n, const uint32_t *audit, uint32_t *tries, uint64_t *audit_key, uint32_t *iota) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tries[i] = 1;
    audit_key[i] = (uint64_t)audit[8 * i] | ((uint64_t)audit[8 * i + 1] << 32);
    iota[i] = (uint32_t)i;
}
// [lo, hi) = positions of the sorted leaf indexes that start with `prefix` (top prefix_bits of height bits)
__global__ void k_prefix_bounds(uint64_t n, const uint64_t *sorted_idx, int shift, uint64_t prefix, uint64_t *out) {
    if (threadIdx.x || blockIdx.x) return;
    for (int side = 0; side < 2; side++) {
        uint64_t want = prefix + side, lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if ((shift >= 64 ? 0 : sorted_idx[mid] >> shift) < want) lo = mid + 1; else hi = mid;
        }
        out[side] = lo;
    }
}
// leaves of one shard out of the global sorted order: local index (prefix stripped), value, blinding
__global__ void k_gather_shard(uint64_t lo, uint64_t cnt, uint64_t mask, const uint64_t *sorted_idx, const uint32_t *who, const uint64_t *values,
                               const uint32_t *blind, uint64_t *o_idx, uint64_t *o_values, uint32_t *o_blind) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    uint64_t u = who[lo + j];
    o_idx[j] = sorted_idx[lo + j] & mask;
    o_values[j] = values[u];
    uint32_t w[8];
    load8(w, blind + 8 * u);
    store8(o_blind + 8 * j, w);
}
// sorted by 64-bit audit-id prefix (stable, so ties are in input order): exact duplicate detection inside runs
__global__ void k_find_dups(uint64_t n, const uint64_t *keys, const uint32_t *who, const uint32_t *audit, unsigned long long *first_dup) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n || keys[j] != keys[j - 1]) return;
    uint32_t a[8], b[8];
    load8(a, audit + 8 * (uint64_t)who[j]);
    for (uint64_t t = j; t-- > 0 && keys[t] == keys[j];) {
        load8(b, audit + 8 * (uint64_t)who[t]);
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
        if (d == 0) { atomicMin(first_dup, (unsigned long long)who[j]); return; }
    }
}
// sorted by candidate index (stable => inside a group the earliest input position is first and keeps the slot):
// every later member of a group lost and re-hashes.  counters[0] = losers this round, counters[1] = min failed position
__global__ void k_resolve_collisions(uint64_t n, const uint64_t *sorted_cand, const uint32_t *who, int hash_id, int height,
                                     uint32_t *cur_seed, uint64_t *cand, uint32_t *tries, unsigned long long *counters) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n || sorted_cand[j] != sorted_cand[j - 1]) return;
    uint64_t u = who[j];
    if (tries[u] > 128) return;  // already failed
    if (rehash_body(u, hash_id, height, cur_seed, cand, tries)) atomicAdd(&counters[0], 1ull);
    else { tries[u] = 129; atomicMin(&counters[1], (unsigned long long)u); }
}
// Opt-in leaf hash of the DAPOL+ paper (NOT the reference's bytes, which hash the commitment only, node.rs:33-36):
//   salt = D(audit_id || "salt_seed" || external_id),  leaf hash = D("leaf" || external_id || salt)
// so that a leaf's hash binds the user's id and a per-user salt instead of being computable from the commitment alone.
__global__ void k_leaf_id_hashes(uint64_t n, int hash_id, const uint32_t *audit, const uint8_t *eid_blob, const uint64_t *eid_off, uint32_t *out, int *too_long) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && leaf_id_hash_body(i, hash_id, audit, eid_blob, eid_off, out)) *too_long = 1;
}
__global__ void k_gather_hashes(uint64_t n, const uint32_t *who, const uint32_t *src, uint32_t *dst) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t w[8];
    load8(w, src + 8 * (uint64_t)who[j]);
    store8(dst + 8 * j, w);
}
__global__ void k_leafpad_hash_b2b(uint64_t T, NodeStore ns, uint64_t leaf_level_off) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < T) leafpad_hash_b2b_body(g, ns, leaf_level_off);
}
__global__ void k_set_leaf_hashes(uint64_t n, NodeStore ns, uint64_t level_off, const uint32_t *pos, const uint32_t *src) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t w[8];
    load8(w, src + 8 * j);
    store8(ns.hash + 8 * (level_off + pos[j]), w);
}
__global__ void k_gather_leaves(uint64_t n, const uint32_t *who, const uint64_t *values, const uint32_t *blind, uint64_t *values_sorted,
                                uint32_t *blind_sorted) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t u = who[j];
    values_sorted[j] = values[u];
    uint32_t w[8];
    load8(w, blind + 8 * u);
    store8(blind_sorted + 8 * j, w);
}

template <int W>
__global__ void __launch_bounds__(128, DAPOL_LEAF_MINB) k_leaf(uint64_t n, uint64_t stride, NodeStore ns, uint64_t level_off, const uint32_t *pos, int hash_id,
                                              const uint64_t *values, const uint32_t *blind, const ge_niels *tab_b,
                                              const ge_niels *tab_bbl) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < stride) leaf_batch_body<W, NODE_BATCH>(i, stride, n, ns, level_off, pos, hash_id, values, blind, tab_b, tab_bbl);
}
struct Seed8 {
    uint32_t w[8];
};
template <int W>
__global__ void __launch_bounds__(128, DAPOL_PAD_MINB) k_pad(uint64_t n, uint64_t stride, NodeStore ns, const uint64_t *pad_dest, int hash_id, Seed8 seed,
                                             const uint64_t *pad_rng, const ge_niels *tab_bbl, const __grid_constant__ PadStreams ps) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < stride) pad_batch_body<W, NODE_BATCH>(g, stride, n, ns, pad_dest, hash_id, seed.w, pad_rng, tab_bbl, ps);
}
// leaves of a top tree: subtree-root records gathered from the shards
__global__ void k_leaf_records(uint64_t n, NodeStore ns, uint64_t level_off, const uint32_t *pos, const uint32_t *recs) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) record_leaf_body(i, ns, level_off, pos, recs);
}
// leaves only (no tree): commitments for dapol_commit_batch
template <int W>
__global__ void __launch_bounds__(128) k_commit(uint64_t n, const uint64_t *values, const uint32_t *blind, uint32_t *out,
                                                const ge_niels *tab_b, const ge_niels *tab_bbl) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int WV = comb_value_window<W>::value;
    constexpr int NWR = 253 / W + 1, NWV = 64 / WV + 1;
    sc rs, rh;
    load8(rs.v, blind + 8 * i);
    sc_half256(rh, rs);
    uint64_t v = values[i];
    uint32_t vw[2] = {(uint32_t)v, (uint32_t)(v >> 32)};
    int32_t dr[NWR], dv[NWV];
    sc_signed_digits<W, NWR>(dr, rh.v, 8);
    sc_signed_digits<WV, NWV>(dv, vw, 2);
    ge acc;  // half point of the commitment: v * (B/2) + (r/2) * B_blinding
    ge_identity(acc);
    ge_comb_accumulate<WV, NWV>(acc, tab_b, dv);
    ge_comb_accumulate<W, NWR>(acc, tab_bbl, dr);
    ge_dc_batch<1> dc;
    dc.init();
    dc.push(acc);
    dc.solve();
    uint32_t cc[8];
    dc.get(0, cc);
    store8(out + 8 * i, cc);
}
// Siblings of leaf_idx[q] (or of the fixed leaf `fixed_idx` when leaf_idx == nullptr: the shard's own root inside the
// top tree), leaf level first, written at [q][lvl0 ..] of arrays with out_stride levels per proof.  Leaf indexes carry
// the shard prefix above `height` bits when prefix_check is set.
__global__ void k_paths(uint64_t k, const uint64_t *leaf_idx, uint64_t fixed_idx, int prefix_check, uint64_t prefix, NodeStore ns,
                        const uint64_t *level_off, uint64_t n_leaf_level, uint32_t *const *pos, int height, int out_stride, int lvl0,
                        uint64_t *o_v, uint32_t *o_r, uint32_t *o_c, uint32_t *o_h, uint32_t *o_lc, uint32_t *o_lh, int *not_found,
                        uint32_t *o_hh, uint32_t *o_lhh) {  // o_hh / o_lhh: upper halves of 64-byte hashes (Blake2b trees), else nullptr
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= k) return;
    uint64_t want = leaf_idx ? leaf_idx[q] : fixed_idx;
    if (prefix_check) {
        if (height < 64 && (want >> height) != prefix) { *not_found = 1; return; }
        if (height < 64) want &= (1ull << height) - 1;
    }
    uint64_t off = level_off[height];
    int64_t slot = find_leaf_slot(ns.idx + off, n_leaf_level, want);
    if (slot < 0 || ns.is_pad[off + slot]) { *not_found = 1; return; }
    uint32_t w[8];
    if (o_lc) { load8(w, ns.comc + 8 * (off + slot)); store8(o_lc + 8 * q, w); }
    if (o_lh) { load8(w, ns.hash + 8 * (off + slot)); store8(o_lh + 8 * q, w); }
    if (o_lhh && ns.hash_hi) { load8(w, ns.hash_hi + 8 * (off + slot)); store8(o_lhh + 8 * q, w); }
    uint64_t p = (uint64_t)slot;
    for (int h = height, lvl = lvl0; h >= 1; h--, lvl++) {
        uint64_t g = level_off[h] + (p ^ 1), o = q * (uint64_t)out_stride + lvl;
        o_v[o] = ns.v[g];
        load8(w, ns.r + 8 * g); store8(o_r + 8 * o, w);
        load8(w, ns.comc + 8 * g); store8(o_c + 8 * o, w);
        load8(w, ns.hash + 8 * g); store8(o_h + 8 * o, w);
        if (o_hh && ns.hash_hi) { load8(w, ns.hash_hi + 8 * g); store8(o_hh + 8 * o, w); }
        if (h > 1) p = pos[h - 1][p >> 1];
    }
}

// ---- microbenchmarks: integer-multiply pipe peak and field-op throughput
template <int VARIANT>
__global__ void k_imad_peak(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    uint32_t z0 = 0, z1 = 1, z2 = 2, z3 = 3, z4 = 4, z5 = 5, z6 = 6, z7 = 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (VARIANT == 0) {  // 8 independent mad.lo chains: 8 IMAD
                asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                             "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                             : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
            } else if (VARIANT == 1) {  // 8 IMAD.WIDE.U32 with a 64-bit addend each = 8 MAC32 (ptxas fuses each mad.lo.cc / madc.hi pair);
                                        // the multiplicand is the neighbour chain's low word, so nothing is loop-invariant
#define DAPOL_WIDE_MAC(lo, hi, m, c) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(m), "r"(c))
                DAPOL_WIDE_MAC(x0, z0, x1, b); DAPOL_WIDE_MAC(x1, z1, x2, b); DAPOL_WIDE_MAC(x2, z2, x3, a); DAPOL_WIDE_MAC(x3, z3, x4, a);
                DAPOL_WIDE_MAC(x4, z4, x5, b); DAPOL_WIDE_MAC(x5, z5, x6, b); DAPOL_WIDE_MAC(x6, z6, x7, a); DAPOL_WIDE_MAC(x7, z7, x0, a);
#undef DAPOL_WIDE_MAC
            } else {  // the carry-chain shape fe_mul uses: (mad.lo.cc, madc.hi.cc) x4 = 4 MAC32 in 8 instructions
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                             "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5; madc.lo.cc.u32 %6, %8, %8, %6; madc.hi.u32 %7, %8, %8, %7;"
                             : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b), "r"(x0 | 1u));
            }
        }
    }
    uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ z0 ^ z1 ^ z2 ^ z3 ^ z4 ^ z5 ^ z6 ^ z7;
    if (r == 0x12345u) out[0] = r;  // practically never; keeps the chains live
}
template <int OP>
__global__ void __launch_bounds__(256) k_fe_bench(uint32_t *out, int iters) {
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; b.v[i] = a.v[i] ^ 0x9e3779b9u; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { fe_mul(a, a, b); fe_mul(b, b, a); }
        else { fe_sq(a, a); fe_sq(b, b); }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a.v[i] ^ b.v[i];
    if (r == 0x12345u) out[0] = r;
}

// ------------------------------------------------------------------------------------------------ ctx
// table of `which` (0 = B/2, 1 = B_blinding) at window W with nw windows
template <int W>
static int build_one_table(dapol_ctx *ctx, ge_niels **tab, int nw, int which) {
    constexpr uint64_t half = 1ull << (W - 1);
    uint64_t total = (uint64_t)nw * half;
    CUDA_TRY(cudaMalloc(tab, total * sizeof(ge_niels)));
    if constexpr (W >= 10) {  // a run of consecutive multiples per thread (one addition per entry, one inversion per run)
        ge *bases = nullptr;
        CUDA_TRY(cudaMalloc(&bases, (size_t)nw * sizeof(ge)));
        k_comb_bases<W><<<1, 32, 0, ctx->stream>>>(bases, nw, which);
        uint64_t runs = total / COMB_RUN;
        k_comb_table_run<W><<<grid_for(runs, 64), 64, 0, ctx->stream>>>(*tab, nw, bases, runs);
        ctx->launches += 2;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(bases);
        CUDA_TRY(e);
    } else {
        k_comb_table<W><<<grid_for(total, 64), 64, 0, ctx->stream>>>(*tab, nw, which, total);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return DAPOL_OK;
}
template <int W>
static int build_tables(dapol_ctx *ctx) {
    // B/2 multiples for 64-bit values (window WV), B_blinding multiples for halved 253-bit scalars (window W)
    constexpr int WV = comb_value_window<W>::value;
    constexpr int NWR = 253 / W + 1, NWV = 64 / WV + 1;
    int rc = build_one_table<WV>(ctx, &ctx->tab_b, NWV, 0);
    if (rc) return rc;
    rc = build_one_table<W>(ctx, &ctx->tab_bbl, NWR, 1);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DAPOL_OK;
}

extern "C" int dapol_ctx_create(int device, int comb_window, dapol_ctx **out) {
    if (!out) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { g_cuda_err = "no such CUDA device"; return DAPOL_ERR_CUDA; }
    CUDA_TRY(cudaSetDevice(device));
    dapol_ctx *ctx = new dapol_ctx();
    ctx->device = device;
    ctx->W = comb_window;
    if (const char *e = getenv("DAPOL_RP_PACK_LANES")) {  // tuning experiments: 0 (a warp per MSM), 4, 8, 16
        int v = atoi(e);
        if (v == 0 || v == 4 || v == 8 || v == 16) ctx->rp_pack_lanes = v;
    }
    if (const char *e = getenv("DAPOL_RP_PACK_MIN_K")) { int v = atoi(e); if (v >= 1) ctx->rp_pack_min_k = v; }
    if (const char *e = getenv("DAPOL_RP_PACK_MAX_N")) { int v = atoi(e); if (v >= 8 && v <= 128) ctx->rp_pack_max_n = v; }
    if (comb_window == 0) {
        // the wide window keeps 9.5 GB of tables in HBM (11 additions per blinding instead of 17); a device that is short
        // of memory (or shared with other contexts) stays with the L2-resident 27 MB tables
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        ctx->W = free_b >= (size_t)DAPOL_WIDE_COMB_MIN_FREE_GB << 30 ? DAPOL_WIDE_COMB_WINDOW : DAPOL_DEFAULT_COMB_WINDOW;
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto &e : ctx->ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaMalloc(&ctx->scratch, 1024));
    {
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    auto build = [&](int w) {
        switch (w) {
#define W_CASE(w_) case w_: return build_tables<w_>(ctx);
            DAPOL_W_CASES(W_CASE)
#undef W_CASE
        }
        return (int)DAPOL_ERR_BAD_ARG;
    };
    int rc = build(ctx->W);
    if (rc == DAPOL_ERR_CUDA && comb_window == 0 && ctx->W != DAPOL_DEFAULT_COMB_WINDOW) {
        // the wide tables did not fit after all (fragmentation, another context grabbed the memory): the L2-resident ones do
        cudaGetLastError();
        cudaFree(ctx->tab_b); cudaFree(ctx->tab_bbl);
        ctx->tab_b = ctx->tab_bbl = nullptr;
        ctx->W = DAPOL_DEFAULT_COMB_WINDOW;
        rc = build(ctx->W);
    }
    if (rc) { dapol_ctx_destroy(ctx); return rc; }
    *out = ctx;
    return DAPOL_OK;
}
extern "C" void dapol_ctx_destroy(dapol_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->tab_b); cudaFree(ctx->tab_bbl); cudaFree(ctx->scratch); cudaFree(ctx->rp_tab);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" int dapol_ctx_set_stream(dapol_ctx *ctx, void *cuda_stream) {
    if (!ctx) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return DAPOL_OK;
}
extern "C" int dapol_ctx_set_leaf_hash_mode(dapol_ctx *ctx, int mode) {
    if (!ctx || (mode != DAPOL_LEAF_HASH_COMMITMENT && mode != DAPOL_LEAF_HASH_ID_SALT)) return DAPOL_ERR_BAD_ARG;
    ctx->leaf_hash_mode = mode;
    return DAPOL_OK;
}
extern "C" int dapol_ctx_set_padding_mode(dapol_ctx *ctx, int mode) {
    if (!ctx || (mode != DAPOL_PADDING_STREAM && mode != DAPOL_PADDING_POSITIONAL)) return DAPOL_ERR_BAD_ARG;
    ctx->pad_mode = mode;
    return DAPOL_OK;
}
extern "C" uint64_t dapol_kernel_launches(const dapol_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int dapol_ctx_params(const dapol_ctx *ctx, int *comb_window, int *node_batch, int *rangeproof_window) {
    if (!ctx) return DAPOL_ERR_BAD_ARG;
    if (comb_window) *comb_window = ctx->W;
    if (node_batch) *node_batch = NODE_BATCH;
    if (rangeproof_window) *rangeproof_window = ctx->rp_W;
    return DAPOL_OK;
}
extern "C" int dapol_last_build_times(const dapol_ctx *ctx, float ms[5]) {
    if (!ctx) return DAPOL_ERR_BAD_ARG;
    memcpy(ms, ctx->last_ms, sizeof(float) * 5);
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ tree build
extern "C" void dapol_tree_destroy(dapol_tree *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStream_t st = t->ctx->stream;
    dfree(t->ns.idx, st); dfree(t->ns.v, st); dfree(t->ns.r, st); dfree(t->ns.comc, st); dfree(t->ns.hash, st); dfree(t->ns.hash_hi, st);
    dfree(t->ns.ext, st); dfree(t->ns.is_pad, st);
    dfree(t->pos_all, st);
    dfree(t->d_pos, st); dfree(t->d_level_off, st); dfree(t->leaf_index_of, st);
    dfree(t->audit_ids, st); dfree(t->akey_sorted, st); dfree(t->akey_who, st);
    delete t;
}

template <int W>
static void launch_leaf_pad(dapol_ctx *ctx, dapol_tree *t, const uint64_t *d_values, const uint32_t *d_blind, const uint64_t *d_pad_dest,
                            const Seed8 &seed, const uint64_t *d_pad_rng, int phase, const PadStreams &ps) {
    int H = t->height;
    if (phase == 0) {
        uint64_t stride = batch_stride(t->n_leaves, k_leaf<W>, 24.0);
        k_leaf<W><<<grid_for(stride, 128), 128, 0, ctx->stream>>>(t->n_leaves, stride, t->ns, t->level_off[H], t->pos[H], t->hash_id, d_values,
                                                                  d_blind, ctx->tab_b, ctx->tab_bbl);
        ctx->launches++;
    } else if (t->n_pads) {
        uint64_t stride = batch_stride(t->n_pads, k_pad<W>, 13.0);
        k_pad<W><<<grid_for(stride, 128), 128, 0, ctx->stream>>>(t->n_pads, stride, t->ns, d_pad_dest, t->hash_id, seed, d_pad_rng, ctx->tab_bbl, ps);
        ctx->launches++;
    }
}

// d_records != nullptr: the leaves are subtree-root records (top tree of a sharded build) instead of (value, blinding).
// pad_level_base != nullptr (host, [height + 1]): RNG block of the first padding node of each level (a shard's levels
// interleave with the other shards' in the single-tree creation order); nullptr: pad_base + running creation ordinal.
static int tree_build_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *d_leaf_idx, const uint64_t *d_values,
                          const uint8_t *d_blindings, const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out,
                          const uint64_t *pad_level_base = nullptr, const uint32_t *d_records = nullptr, const uint32_t *d_leaf_hashes = nullptr) {
    if (!ctx || !out || !d_leaf_idx || !pad_seed) return DAPOL_ERR_BAD_ARG;
    if (!d_records && (!d_values || !d_blindings)) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    // D = Blake2b (64-byte digests): new_blank + build only, as in the reference (Dapol::new insists on 32 bytes, mod.rs:101-103)
    const bool b2b = hash_id == DAPOL_HASH_BLAKE2B;
    if (hash_id != DAPOL_HASH_BLAKE3 && hash_id != DAPOL_HASH_BLAKE2S && !b2b) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (b2b && (d_records || d_leaf_hashes || pad_level_base)) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_TREE_HEIGHT_TOO_BIG;
    if (height < 0 || n == 0 || (height == 0 && n != 1) || n >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int H = height;
    CUDA_TRY(cudaEventRecord(ctx->ev[0], st));

    // ---- sizes of every level from one pass over adjacent leaves (+ validation)
    unsigned long long h_hist[65];
    CUDA_TRY(cudaMemsetAsync(ctx->scratch, 0, 65 * 8, st));
    k_leaf_msb_hist<<<grid_for(n, 256), 256, 0, st>>>(d_leaf_idx, n, H, ctx->scratch);
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(h_hist, ctx->scratch, 65 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_hist[64]) return DAPOL_ERR_BAD_ARG;  // unsorted / duplicate / out-of-tree leaf index

    dapol_tree *t = new dapol_tree();
    t->ctx = ctx; t->hash_id = hash_id; t->height = H; t->n_leaves = n;
    t->level_off.assign(H + 1, 0); t->level_n.assign(H + 1, 0); t->n_real.assign(H + 1, 0);
    t->pos.assign(H + 1, nullptr);
    std::vector<uint64_t> pos_off(H + 2, 0);
    std::vector<uint64_t> &npads = t->npads;
    npads.assign(H + 1, 0);
    {   // real nodes at level h = 1 + #{adjacent pairs whose highest differing bit >= H - h}
        uint64_t acc = 0;
        for (int h = 0; h <= H; h++) {
            if (h >= 1) acc += h_hist[H - h];
            t->n_real[h] = 1 + acc;
        }
    }
    uint64_t T = 1, total_pads = 0;
    t->level_n[0] = 1;
    for (int h = 1; h <= H; h++) {
        t->level_n[h] = 2 * t->n_real[h - 1];
        npads[h] = t->level_n[h] - t->n_real[h];
        t->level_off[h] = T;
        T += t->level_n[h];
        total_pads += npads[h];
        pos_off[h + 1] = pos_off[h] + ((t->n_real[h] + 63) & ~63ull);
    }
    t->T = T; t->n_pads = total_pads;

    uint8_t *arena_mem = nullptr;
#define TRY_T(expr)                                                          \
    do {                                                                     \
        cudaError_t e_ = (expr);                                             \
        if (e_ != cudaSuccess) {                                             \
            g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(e_); \
            dfree(arena_mem, st); dapol_tree_destroy(t); return DAPOL_ERR_CUDA; \
        }                                                                    \
    } while (0)
    TRY_T(dmalloc(&t->ns.idx, T * 8, st));
    TRY_T(dmalloc(&t->ns.v, T * 8, st));
    TRY_T(dmalloc(&t->ns.r, T * 32, st));
    TRY_T(dmalloc(&t->ns.comc, T * 32, st));
    TRY_T(dmalloc(&t->ns.hash, T * 32, st));
    if (b2b) TRY_T(dmalloc(&t->ns.hash_hi, T * 32, st));
    TRY_T(dmalloc(&t->ns.ext, T * 128, st));
    TRY_T(dmalloc(&t->ns.is_pad, T, st));
    TRY_T(dmalloc(&t->pos_all, (pos_off[H + 1] + 64) * 4, st));
    for (int h = 1; h <= H; h++) t->pos[h] = t->pos_all + pos_off[h];
    uint64_t ntiles_max = (n + SCAN_TILE - 1) / SCAN_TILE;
    Arena ar;
    ar.size = 2 * Arena::need(n, 8) + Arena::need(n, 8) + Arena::need(ntiles_max, 8) + 2 * Arena::need(total_pads + 1, 8);
    TRY_T(dmalloc(&arena_mem, ar.size, st));
    ar.base = arena_mem;
    uint64_t *realA = ar.take<uint64_t>(n), *realB = ar.take<uint64_t>(n), *d_flags = ar.take<uint64_t>(n);
    uint64_t *d_tiles = ar.take<uint64_t>(ntiles_max), *d_pad_dest = ar.take<uint64_t>(total_pads + 1);
    uint64_t *d_pad_rng = ar.take<uint64_t>(total_pads + 1);
    if (H == 0) {
        TRY_T(cudaMemcpyAsync(t->ns.idx, d_leaf_idx, 8, cudaMemcpyDeviceToDevice, st));
        TRY_T(cudaMemsetAsync(t->ns.is_pad, 0, 1, st));
        TRY_T(cudaMemsetAsync(t->pos_all, 0, 4, st));
        t->pos[0] = t->pos_all;
    } else {
        TRY_T(cudaMemsetAsync(t->ns.idx, 0, 8, st));  // root: TreeIndex::zero(0)
        TRY_T(cudaMemsetAsync(t->ns.is_pad, 0, 1, st));
    }
    // ---- structure: per level flags -> scan -> slots, parents, padding destinations.  Padding RNG ordinal =
    // creation order of smtree's build: level H..1, left to right.
    // padding randomness: creation-order stream (the reference's RNG, seeded) or keyed by position (ctx->pad_mode)
    const int positional = ctx->pad_mode == DAPOL_PADDING_POSITIONAL;
    PadStreams ps;
    memset(&ps, 0, sizeof ps);
    ps.positional = positional; ps.levels = H; ps.level0 = positional && pad_level_base ? (int)pad_level_base[0] : 0;
    const uint64_t *cur = d_leaf_idx;
    uint64_t ord = 0;
    for (int h = H; h >= 1; h--) {
        ps.start[h] = ord;
        uint64_t c = t->n_real[h];
        uint64_t *next = (cur == realA) ? realB : realA;
        unsigned ntiles = (unsigned)((c + SCAN_TILE - 1) / SCAN_TILE);
        k_struct_flags_tiles<<<ntiles, SCAN_BLOCK, 0, st>>>(cur, c, d_flags, d_tiles);
        k_scan_tile_offsets<<<1, SCAN_BLOCK, 0, st>>>(d_tiles, ntiles);
        k_struct_apply<<<ntiles, SCAN_BLOCK, 0, st>>>(cur, c, d_flags, d_tiles, t->pos[h], next, t->level_off[h], t->ns, d_pad_dest, ord,
                                                      d_pad_rng, pad_level_base ? pad_level_base[h] : (positional ? 0 : pad_base + ord), positional);
        ctx->launches += 3;
        ord += npads[h];
        cur = next;
    }
    ps.start[0] = ord;
    TRY_T(cudaEventRecord(ctx->ev[1], st));
    // ---- leaves, padding nodes
    Seed8 seed;
    memcpy(seed.w, pad_seed, 32);
    const uint32_t *d_blind = reinterpret_cast<const uint32_t *>(d_blindings);
    for (int phase = 0; phase < 2; phase++) {
        if (phase == 0 && d_records) {
            k_leaf_records<<<grid_for(n, 128), 128, 0, st>>>(n, t->ns, t->level_off[H], t->pos[H], d_records);
            ctx->launches++;
        } else switch (ctx->W) {
#define W_CASE(w) case w: launch_leaf_pad<w>(ctx, t, d_values, d_blind, d_pad_dest, seed, d_pad_rng, phase, ps); break;
            DAPOL_W_CASES(W_CASE)
#undef W_CASE
        }
        if (phase == 0 && d_leaf_hashes) {  // id / salt leaf hashes (opt-in mode) replace D(compress(com)) before any parent is hashed
            k_set_leaf_hashes<<<grid_for(n, 256), 256, 0, st>>>(n, t->ns, t->level_off[H], t->pos[H], d_leaf_hashes);
            ctx->launches++;
        }
        if (phase == 1 && b2b) {  // 64-byte digests of the leaf-level and padding nodes (the node kernels left 32-byte placeholders)
            k_leafpad_hash_b2b<<<grid_for(T, 128), 128, 0, st>>>(T, t->ns, t->level_off[H]);
            ctx->launches++;
        }
        TRY_T(cudaEventRecord(ctx->ev[2 + phase], st));
    }
    // pointer tables for the compress pass and path extraction
    TRY_T(dmalloc(&t->d_pos, (H + 1) * sizeof(uint32_t *), st));
    TRY_T(cudaMemcpyAsync(t->d_pos, t->pos.data(), (H + 1) * sizeof(uint32_t *), cudaMemcpyHostToDevice, st));
    TRY_T(dmalloc(&t->d_level_off, (H + 1) * 8, st));
    TRY_T(cudaMemcpyAsync(t->d_level_off, t->level_off.data(), (H + 1) * 8, cudaMemcpyHostToDevice, st));
    // ---- merges: sums level by level, one compress pass over all internal nodes, hashes level by level (dapol_merge.cu)
    dapol_launch_merges(ctx, t);
    TRY_T(cudaEventRecord(ctx->ev[4], st));
    TRY_T(cudaMemcpyAsync(t->root_ext, t->ns.ext, 128, cudaMemcpyDeviceToHost, st));  // root = global node 0
    TRY_T(cudaMemcpyAsync(t->root_comc, t->ns.comc, 32, cudaMemcpyDeviceToHost, st));
    TRY_T(cudaMemcpyAsync(t->root_hash, t->ns.hash, 32, cudaMemcpyDeviceToHost, st));
    if (b2b) TRY_T(cudaMemcpyAsync(t->root_hash + 8, t->ns.hash_hi, 32, cudaMemcpyDeviceToHost, st));
    TRY_T(cudaGetLastError());
    TRY_T(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&ctx->last_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    cudaEventElapsedTime(&ctx->last_ms[4], ctx->ev[0], ctx->ev[4]);
    // the extended points are only needed while merging
    dfree(t->ns.ext, st);
    t->ns.ext = nullptr;
    dfree(arena_mem, st);
#undef TRY_T
    *out = t;
    return DAPOL_OK;
}

extern "C" int dapol_tree_build_from_nodes_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *d_leaf_idx,
                                               const uint64_t *d_values, const uint8_t *d_blindings, const uint8_t pad_seed[32],
                                               uint64_t pad_base, dapol_tree **out) {
    return tree_build_dev(ctx, hash_id, height, n, d_leaf_idx, d_values, d_blindings, pad_seed, pad_base, out);
}
extern "C" int dapol_tree_build_from_nodes(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *leaf_idx,
                                           const uint64_t *values, const uint8_t *blindings, const uint8_t pad_seed[32], uint64_t pad_base,
                                           dapol_tree **out) {
    if (!ctx || !leaf_idx || !values || !blindings || n == 0) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = 2 * Arena::need(n, 8) + Arena::need(n, 32);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_idx = ar.take<uint64_t>(n), *d_val = ar.take<uint64_t>(n);
    uint8_t *d_bl = ar.take<uint8_t>(n * 32);
    cudaMemcpyAsync(d_idx, leaf_idx, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_val, values, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_bl, blindings, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    int rc = tree_build_dev(ctx, hash_id, height, n, d_idx, d_val, d_bl, pad_seed, pad_base, out);
    dfree(mem, ctx->stream);
    return rc;
}

// ---- Dapol::update (src/dapol/mod.rs:210-213 -> smtree SparseMerkleTree::update), batched: the leaves of `t` with k
// leaves inserted (or replaced where the index is already a leaf).  Both lists are sorted, so the union is a merge by rank:
// binary searches + one scan of the "replaces an old leaf" flags; then the level-synchronous build runs over the merged leaves.
__global__ void k_update_find(uint64_t k, const uint64_t *new_idx, const uint64_t *old_lvl_idx, uint64_t n_lvl, const uint8_t *old_is_pad,
                              int height, uint64_t *found, int *bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const uint64_t x = new_idx[i];
    if ((height < 64 && (x >> height)) || (i && new_idx[i - 1] >= x)) *bad = 1;  // out of the tree / not strictly increasing
    int64_t slot = find_leaf_slot(old_lvl_idx, n_lvl, x);
    found[i] = (slot >= 0 && !old_is_pad[slot]) ? 1 : 0;
}
DAPOL_HD_INLINE uint64_t lower_bound_u64(const uint64_t *a, uint64_t n, uint64_t x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
// old real leaf j (slot pos[j] of level H) -> rank j + #{new < idx} - #{replaced old before it}; dropped if a new leaf has its index
__global__ void k_update_place_old(uint64_t n, NodeStore ns, uint64_t level_off, const uint32_t *pos, uint64_t k, const uint64_t *new_idx,
                                   const uint64_t *found_prefix, uint64_t *o_idx, uint64_t *o_v, uint32_t *o_r) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t g = level_off + pos[j], x = ns.idx[g];
    const uint64_t lb = lower_bound_u64(new_idx, k, x);
    if (lb < k && new_idx[lb] == x) return;  // replaced
    const uint64_t dst = j + lb - found_prefix[lb];
    uint32_t w[8];
    o_idx[dst] = x; o_v[dst] = ns.v[g];
    load8(w, ns.r + 8 * g); store8(o_r + 8 * dst, w);
}
// new leaf i -> rank i + #{old real leaves < idx} - #{replaced old before it}
__global__ void k_update_place_new(uint64_t k, const uint64_t *new_idx, const uint64_t *new_v, const uint32_t *new_r, const uint64_t *old_sorted_idx,
                                   uint64_t n, const uint64_t *found_prefix, uint64_t *o_idx, uint64_t *o_v, uint32_t *o_r) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const uint64_t x = new_idx[i];
    const uint64_t dst = i + lower_bound_u64(old_sorted_idx, n, x) - found_prefix[i];
    uint32_t w[8];
    o_idx[dst] = x; o_v[dst] = new_v[i];
    load8(w, new_r + 8 * i); store8(o_r + 8 * dst, w);
}
__global__ void k_gather_real_idx(uint64_t n, NodeStore ns, uint64_t level_off, const uint32_t *pos, uint64_t *out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = ns.idx[level_off + pos[j]];
}
extern "C" int dapol_tree_update(const dapol_tree *t, uint64_t k, const uint64_t *leaf_idx, const uint64_t *values, const uint8_t *blindings,
                                 const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out) {
    if (!t || !k || !leaf_idx || !values || !blindings || !pad_seed || !out) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    // a shard is updated through a rebuild of its slice; id / salt leaf hashes cannot be recomputed from (value, blinding)
    if (t->top || t->custom_leaf_hashes || t->height == 0 || t->n_leaves + k >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    dapol_ctx *ctx = t->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int H = t->height;
    const uint64_t n = t->n_leaves, cap = n + k;
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)(k + 1), st);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = 2 * Arena::need(k, 8) + Arena::need(k, 32) + 2 * Arena::need(k + 1, 8) + Arena::need(n, 8) + 2 * Arena::need(cap, 8) + Arena::need(cap, 32) +
              Arena::need(scan_bytes, 1) + 256;
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *n_idx = ar.take<uint64_t>(k), *n_v = ar.take<uint64_t>(k);
    uint32_t *n_r = ar.take<uint32_t>(8 * k);
    uint64_t *found = ar.take<uint64_t>(k + 1), *prefix = ar.take<uint64_t>(k + 1), *old_idx = ar.take<uint64_t>(n);
    uint64_t *o_idx = ar.take<uint64_t>(cap), *o_v = ar.take<uint64_t>(cap);
    uint32_t *o_r = ar.take<uint32_t>(8 * cap);
    uint8_t *scan_tmp = ar.take<uint8_t>(scan_bytes);
    int *d_bad = ar.take<int>(1), bad = 0;
    uint64_t replaced = 0;
    int rc = DAPOL_OK;
    cudaMemsetAsync(d_bad, 0, 4, st);
    cudaMemsetAsync(found, 0, (k + 1) * 8, st);
    cudaMemcpyAsync(n_idx, leaf_idx, k * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(n_v, values, k * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(n_r, blindings, k * 32, cudaMemcpyHostToDevice, st);
    const uint64_t loff = t->level_off[H];
    k_update_find<<<grid_for(k, 128), 128, 0, st>>>(k, n_idx, t->ns.idx + loff, t->level_n[H], t->ns.is_pad + loff, H, found, d_bad);
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, found, prefix, (int)(k + 1), st);
    k_gather_real_idx<<<grid_for(n, 256), 256, 0, st>>>(n, t->ns, loff, t->pos[H], old_idx);
    k_update_place_old<<<grid_for(n, 256), 256, 0, st>>>(n, t->ns, loff, t->pos[H], k, n_idx, prefix, o_idx, o_v, o_r);
    k_update_place_new<<<grid_for(k, 128), 128, 0, st>>>(k, n_idx, n_v, n_r, old_idx, n, prefix, o_idx, o_v, o_r);
    ctx->launches += 5;
    cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&replaced, prefix + k, 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { dfree(mem, st); CUDA_TRY(cudaGetLastError()); return DAPOL_ERR_CUDA; }
    if (bad) rc = DAPOL_ERR_BAD_ARG;
    if (rc == DAPOL_OK)
        rc = tree_build_dev(ctx, t->hash_id, H, cap - replaced, o_idx, o_v, reinterpret_cast<const uint8_t *>(o_r), pad_seed, pad_base, out, nullptr, nullptr, nullptr);
    dfree(mem, st);
    return rc;
}

// ---- build_leaf_nodes (mod.rs:323-399) in two device stages, so that a sharded build can hash its own slice of the
// liabilities and run the (global) duplicate / collision rules over the gathered per-user records.
// stage 1: per-user hashing of a slice (k_derive); tries / audit_key / iota are written only when given.
static int derive_stage(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob, const uint64_t *d_iid_off,
                        const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint8_t *d_seed, uint32_t seed_len, uint32_t *audit,
                        uint32_t *cur_seed, uint64_t *cand, uint32_t *blind, uint32_t *tries, uint64_t *akey, uint32_t *iota, int *d_too_long) {
    k_derive<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, hash_id, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, d_seed, seed_len, height, audit,
                                                        cur_seed, cand, blind, tries, akey, iota, d_too_long);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}
// scratch of stage 2 for n users
struct AssignScratch {
    uint64_t *akey, *keys_sorted;
    uint32_t *tries, *iota, *who;
    uint8_t *cub_tmp;
    size_t cub_bytes;
    unsigned long long *counters;  // [0] losers, [1] min failed pos, [2] first dup pos, [3] too-long flag (int)
    static size_t cub_need(uint64_t n, cudaStream_t st) {
        size_t b = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (int)n, 0, 64, st);
        return b;
    }
    static size_t need(uint64_t n, size_t cub_bytes) { return 2 * Arena::need(n, 8) + 3 * Arena::need(n, 4) + Arena::need(cub_bytes, 1) + 1024; }
    void take(Arena &ar, uint64_t n, size_t cb) {
        akey = ar.take<uint64_t>(n); keys_sorted = ar.take<uint64_t>(n);
        tries = ar.take<uint32_t>(n); iota = ar.take<uint32_t>(n); who = ar.take<uint32_t>(n);
        cub_tmp = ar.take<uint8_t>(cb); cub_bytes = cb;
        counters = ar.take<unsigned long long>(8);
    }
};
// stage 2 over ALL users in input order: duplicate internal ids (mod.rs:345-349 <=> identical audit ids), shuffle_index's
// first-come-first-served collision rule as a fix-point (mod.rs:408-441), final sort by index (mod.rs:396).
// On return sc.keys_sorted = sorted leaf indexes, sc.who = input position of each, cand[i] = index of user i.
// The caller initialised sc.counters = {0, ~0, ~0, too_long} and sc.tries / sc.akey / sc.iota (k_derive or k_assign_init).
static int assign_stage(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint32_t *audit, uint32_t *cur_seed, uint64_t *cand,
                        AssignScratch &sc, uint64_t *err_pos) {
    cudaStream_t st = ctx->stream;
    unsigned long long h_cnt[4] = {0, ~0ull, ~0ull, 0};
    cub::DeviceRadixSort::SortPairs(sc.cub_tmp, sc.cub_bytes, sc.akey, sc.keys_sorted, sc.iota, sc.who, (int)n, 0, 64, st);
    k_find_dups<<<grid_for(n, 256), 256, 0, st>>>(n, sc.keys_sorted, sc.who, audit, sc.counters + 2);
    ctx->launches += 2;
    int end_bit = height < 1 ? 1 : height;
    for (int round = 0;; round++) {
        cub::DeviceRadixSort::SortPairs(sc.cub_tmp, sc.cub_bytes, cand, sc.keys_sorted, sc.iota, sc.who, (int)n, 0, end_bit, st);
        CUDA_TRY(cudaMemsetAsync(sc.counters, 0, 8, st));
        k_resolve_collisions<<<grid_for(n, 256), 256, 0, st>>>(n, sc.keys_sorted, sc.who, hash_id, height, cur_seed, cand, sc.tries, sc.counters);
        ctx->launches += 2;
        CUDA_TRY(cudaMemcpyAsync(h_cnt, sc.counters, 32, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_cnt[0] == 0) break;
        if (round > 4096) return DAPOL_ERR_BAD_ARG;
    }
    if ((int)h_cnt[3]) return DAPOL_ERR_BAD_ARG;  // id too long for single-chunk hashing (> 1024 B)
    unsigned long long f = h_cnt[1], d = h_cnt[2];
    if (d != ~0ull && d <= f) { if (err_pos) *err_pos = d; return DAPOL_ERR_DUPLICATED_INTERNAL_ID; }
    if (f != ~0ull) { if (err_pos) *err_pos = f; return DAPOL_ERR_FAILED_TO_MAP_INDEX; }
    return DAPOL_OK;
}
static int liabilities_check(int hash_id, int height, uint64_t n, uint64_t seed_len) {
    if (hash_id != DAPOL_HASH_BLAKE3 && hash_id != DAPOL_HASH_BLAKE2S) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_TREE_HEIGHT_TOO_BIG;
    if (height < 0) return DAPOL_ERR_BAD_ARG;
    if (height < 64 && (1ull << height) < n * 2) return DAPOL_ERR_SPARSITY_TOO_SMALL;  // MIN_SPARSITY = 2 (mod.rs:27,110)
    if (n == 0 || height == 0 || n >= (1ull << 31) || seed_len > 512) return DAPOL_ERR_BAD_ARG;
    return DAPOL_OK;
}

// Dapol::new: checks (mod.rs:101-116), build_leaf_nodes on the device (mod.rs:323-399), sort, build.
static int liabilities_build_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob, const uint64_t *d_iid_off,
                                 const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint64_t *d_values, const uint8_t *audit_seed,
                                 uint64_t seed_len, const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out, uint64_t *err_pos) {
    if (!ctx || !out) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    int rc = liabilities_check(hash_id, height, n, seed_len);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t cub_bytes = AssignScratch::cub_need(n, st);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = 2 * Arena::need(n, 32) + Arena::need(n, 8) + Arena::need(n, 32) + AssignScratch::need(n, cub_bytes) + Arena::need(seed_len + 1, 1);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint32_t *audit = nullptr;  // survives in the tree: dapol_tree_index_of looks liabilities up by their audit id
    uint32_t *cur_seed = ar.take<uint32_t>(8 * n), *blind = ar.take<uint32_t>(8 * n);
    uint64_t *values_sorted = ar.take<uint64_t>(n);
    uint32_t *blind_sorted = ar.take<uint32_t>(8 * n);
    AssignScratch sc;
    sc.take(ar, n, cub_bytes);
    uint8_t *d_seed = ar.take<uint8_t>(seed_len + 1);
    uint64_t *cand = nullptr;
#define TRY_L(expr)                                                          \
    do {                                                                     \
        cudaError_t e_ = (expr);                                             \
        if (e_ != cudaSuccess) {                                             \
            g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(e_); \
            dfree(mem, st); dfree(cand, st); dfree(audit, st); return DAPOL_ERR_CUDA; \
        }                                                                    \
    } while (0)
    TRY_L(dmalloc(&cand, n * 8, st));  // survives as the id -> leaf index map of the tree
    TRY_L(dmalloc(&audit, n * 32, st));
    if (seed_len) TRY_L(cudaMemcpyAsync(d_seed, audit_seed, seed_len, cudaMemcpyHostToDevice, st));
    unsigned long long h_cnt[4] = {0, ~0ull, ~0ull, 0};
    TRY_L(cudaMemcpyAsync(sc.counters, h_cnt, 32, cudaMemcpyHostToDevice, st));
    rc = derive_stage(ctx, hash_id, height, n, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, d_seed, (uint32_t)seed_len, audit, cur_seed, cand, blind,
                      sc.tries, sc.akey, sc.iota, reinterpret_cast<int *>(sc.counters + 3));
    if (rc == DAPOL_OK) rc = assign_stage(ctx, hash_id, height, n, audit, cur_seed, cand, sc, err_pos);
    if (rc != DAPOL_OK) { dfree(mem, st); dfree(cand, st); dfree(audit, st); return rc; }
    // the last sort (no losers) is the final sorted order: result.sort_by_key(index) (mod.rs:396)
    k_gather_leaves<<<grid_for(n, 256), 256, 0, st>>>(n, sc.who, d_values, blind, values_sorted, blind_sorted);
    ctx->launches++;
    uint32_t *leaf_hashes = nullptr, *leaf_hashes_in = nullptr;
    if (ctx->leaf_hash_mode == DAPOL_LEAF_HASH_ID_SALT) {
        TRY_L(dmalloc(&leaf_hashes_in, n * 32, st));
        if (dmalloc(&leaf_hashes, n * 32, st) != cudaSuccess) { dfree(leaf_hashes_in, st); dfree(mem, st); dfree(cand, st); dfree(audit, st); return DAPOL_ERR_CUDA; }
        k_leaf_id_hashes<<<grid_for(n, 128), 128, 0, st>>>(n, hash_id, audit, d_eid_blob, d_eid_off, leaf_hashes_in, reinterpret_cast<int *>(sc.counters + 3));
        k_gather_hashes<<<grid_for(n, 256), 256, 0, st>>>(n, sc.who, leaf_hashes_in, leaf_hashes);
        ctx->launches += 2;
    }
    rc = tree_build_dev(ctx, hash_id, height, n, sc.keys_sorted, values_sorted, reinterpret_cast<const uint8_t *>(blind_sorted), pad_seed, pad_base, out,
                        nullptr, nullptr, leaf_hashes);
    dfree(leaf_hashes, st); dfree(leaf_hashes_in, st);
    dfree(mem, st);
    if (rc != DAPOL_OK) { dfree(cand, st); dfree(audit, st); return rc; }
    (*out)->leaf_index_of = cand;
    (*out)->custom_leaf_hashes = leaf_hashes != nullptr;
    (*out)->audit_ids = audit;
    (*out)->audit_seed.assign(audit_seed, audit_seed + seed_len);
#undef TRY_L
    return DAPOL_OK;
}
extern "C" int dapol_tree_build_from_liabilities_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob,
                                                     const uint64_t *d_iid_off, const uint8_t *d_eid_blob, const uint64_t *d_eid_off,
                                                     const uint64_t *d_values, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                                     const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out, uint64_t *err_pos) {
    return liabilities_build_dev(ctx, hash_id, height, n, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, d_values, audit_seed, audit_seed_len,
                                 pad_seed, pad_base, out, err_pos);
}
extern "C" int dapol_tree_build_from_liabilities(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *iid_blob,
                                                 const uint64_t *iid_off, const uint8_t *eid_blob, const uint64_t *eid_off,
                                                 const uint64_t *values, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                                 const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out, uint64_t *err_pos) {
    if (!ctx || !iid_off || !eid_off || !values || !out) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    if (height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_TREE_HEIGHT_TOO_BIG;
    if (height >= 0 && height < 64 && (1ull << height) < n * 2) return DAPOL_ERR_SPARSITY_TOO_SMALL;
    if (n == 0) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t ib = iid_off[n], eb = eid_off[n];
    cudaStream_t st = ctx->stream;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(ib + 1, 1) + Arena::need(eb + 1, 1) + 3 * Arena::need(n + 1, 8);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint8_t *d_ib = ar.take<uint8_t>(ib + 1), *d_eb = ar.take<uint8_t>(eb + 1);
    uint64_t *d_io = ar.take<uint64_t>(n + 1), *d_eo = ar.take<uint64_t>(n + 1), *d_v = ar.take<uint64_t>(n + 1);
    if (ib) cudaMemcpyAsync(d_ib, iid_blob, ib, cudaMemcpyHostToDevice, st);
    if (eb) cudaMemcpyAsync(d_eb, eid_blob, eb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_io, iid_off, (n + 1) * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_eo, eid_off, (n + 1) * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_v, values, n * 8, cudaMemcpyHostToDevice, st);
    int rc = liabilities_build_dev(ctx, hash_id, height, n, d_ib, d_io, d_eb, d_eo, d_v, audit_seed, audit_seed_len, pad_seed, pad_base, out, err_pos);
    dfree(mem, st);
    return rc;
}

// ------------------------------------------------------------------------------------------------ sharded build
extern "C" int dapol_leaves_derive_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob, const uint64_t *d_iid_off,
                                       const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                       uint8_t *d_audit, uint8_t *d_seed_state, uint64_t *d_cand, uint8_t *d_blind) {
    if (!ctx || !d_iid_off || !d_eid_off || !d_audit || !d_seed_state || !d_cand || !d_blind) return DAPOL_ERR_BAD_ARG;
    if (hash_id != DAPOL_HASH_BLAKE3 && hash_id != DAPOL_HASH_BLAKE2S) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_TREE_HEIGHT_TOO_BIG;
    if (height < 1 || n == 0 || n >= (1ull << 31) || audit_seed_len > 512) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint8_t *d_seed = nullptr;
    CUDA_TRY(dmalloc(&d_seed, audit_seed_len + 8, st));
    int *d_flag = reinterpret_cast<int *>(d_seed + ((audit_seed_len + 3) & ~3ull)), flag = 0;
    cudaMemsetAsync(d_flag, 0, 4, st);
    if (audit_seed_len) cudaMemcpyAsync(d_seed, audit_seed, audit_seed_len, cudaMemcpyHostToDevice, st);
    int rc = derive_stage(ctx, hash_id, height, n, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, d_seed, (uint32_t)audit_seed_len,
                          reinterpret_cast<uint32_t *>(d_audit), reinterpret_cast<uint32_t *>(d_seed_state), d_cand,
                          reinterpret_cast<uint32_t *>(d_blind), nullptr, nullptr, nullptr, d_flag);
    cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    dfree(d_seed, st);
    if (rc) return rc;
    CUDA_TRY(e);
    return flag ? DAPOL_ERR_BAD_ARG : DAPOL_OK;
}
extern "C" int dapol_leaves_assign_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n_total, const uint8_t *d_audit, uint8_t *d_seed_state,
                                       uint64_t *d_cand, const uint8_t *d_blind, const uint64_t *d_values, int prefix_bits, uint64_t prefix,
                                       uint64_t *d_out_idx, uint64_t *d_out_values, uint8_t *d_out_blind, uint64_t cap, uint64_t *n_out,
                                       uint64_t *err_pos) {
    if (!ctx || !d_audit || !d_seed_state || !d_cand || !d_blind || !d_values || !n_out) return DAPOL_ERR_BAD_ARG;
    int rc = liabilities_check(hash_id, height, n_total, 0);
    if (rc) return rc;
    if (prefix_bits < 0 || prefix_bits >= height || prefix_bits > 16 || (prefix >> prefix_bits)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = n_total;
    size_t cub_bytes = AssignScratch::cub_need(n, st);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = AssignScratch::need(n, cub_bytes) + 256;
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    AssignScratch sc;
    sc.take(ar, n, cub_bytes);
    unsigned long long h_cnt[4] = {0, ~0ull, ~0ull, 0};
    cudaMemcpyAsync(sc.counters, h_cnt, 32, cudaMemcpyHostToDevice, st);
    const uint32_t *audit = reinterpret_cast<const uint32_t *>(d_audit);
    k_assign_init<<<grid_for(n, 256), 256, 0, st>>>(n, audit, sc.tries, sc.akey, sc.iota);
    ctx->launches++;
    rc = assign_stage(ctx, hash_id, height, n, audit, reinterpret_cast<uint32_t *>(d_seed_state), d_cand, sc, err_pos);
    if (rc) { dfree(mem, st); return rc; }
    uint64_t bounds[2] = {0, n};
    const int shift = height - prefix_bits;
    if (prefix_bits) {
        k_prefix_bounds<<<1, 32, 0, st>>>(n, sc.keys_sorted, shift, prefix, reinterpret_cast<uint64_t *>(sc.counters + 4));
        ctx->launches++;
        cudaMemcpyAsync(bounds, sc.counters + 4, 16, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { dfree(mem, st); CUDA_TRY(cudaGetLastError()); return DAPOL_ERR_CUDA; }
    }
    const uint64_t cnt = bounds[1] - bounds[0];
    *n_out = cnt;
    if (cnt > cap || (cnt && (!d_out_idx || !d_out_values || !d_out_blind))) { dfree(mem, st); return DAPOL_ERR_BUFFER; }
    if (cnt) {
        k_gather_shard<<<grid_for(cnt, 256), 256, 0, st>>>(bounds[0], cnt, shift >= 64 ? ~0ull : (1ull << shift) - 1, sc.keys_sorted, sc.who, d_values,
                                                           reinterpret_cast<const uint32_t *>(d_blind), d_out_idx, d_out_values,
                                                           reinterpret_cast<uint32_t *>(d_out_blind));
        ctx->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    dfree(mem, st);
    CUDA_TRY(e);
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}
extern "C" int dapol_tree_level_pad_counts_dev(dapol_ctx *ctx, int height, uint64_t n, const uint64_t *d_leaf_idx, uint64_t *counts) {
    if (!ctx || !d_leaf_idx || !counts || height < 0 || height > DAPOL_MAX_TREE_HEIGHT || n == 0 || n >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    unsigned long long h_hist[65];
    CUDA_TRY(cudaMemsetAsync(ctx->scratch, 0, 65 * 8, st));
    k_leaf_msb_hist<<<grid_for(n, 256), 256, 0, st>>>(d_leaf_idx, n, height, ctx->scratch);
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(h_hist, ctx->scratch, 65 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_hist[64]) return DAPOL_ERR_BAD_ARG;
    uint64_t acc = 0, prev_real = 1;
    counts[0] = 0;
    for (int h = 1; h <= height; h++) {  // same level plan as tree_build_dev
        acc += h_hist[height - h];
        uint64_t real = 1 + acc;
        counts[h] = 2 * prev_real - real;
        prev_real = real;
    }
    return DAPOL_OK;
}
extern "C" int dapol_tree_build_shard_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *d_leaf_idx, const uint64_t *d_values,
                                          const uint8_t *d_blindings, const uint8_t pad_seed[32], const uint64_t *pad_level_base, dapol_tree **out) {
    if (!pad_level_base) return DAPOL_ERR_BAD_ARG;
    return tree_build_dev(ctx, hash_id, height, n, d_leaf_idx, d_values, d_blindings, pad_seed, 0, out, pad_level_base, nullptr);
}
extern "C" int dapol_tree_root_record(const dapol_tree *t, uint8_t *rec) {
    if (!t || !rec) return DAPOL_ERR_BAD_ARG;
    if (t->ns.hash_hi) return DAPOL_ERR_INVALID_DIGEST_SIZE;  // the record carries a 32-byte hash (sharded builds = Dapol::new: 32-byte digests)
    memcpy(rec, t->root_ext, 128);
    uint64_t v = 0;
    int rc = dapol_tree_level_copy(t, 0, nullptr, &v, rec + 192, rec + 128, rec + 160, nullptr);
    memcpy(rec + 224, &v, 8);
    return rc;
}
extern "C" int dapol_tree_build_from_records(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *leaf_idx, const uint8_t *records,
                                             const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out) {
    if (!ctx || !leaf_idx || !records || n == 0 || n > (1ull << 20)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(n, 8) + Arena::need(n, DAPOL_RECORD_BYTES);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_idx = ar.take<uint64_t>(n);
    uint8_t *d_rec = ar.take<uint8_t>(n * DAPOL_RECORD_BYTES);
    cudaMemcpyAsync(d_idx, leaf_idx, n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_rec, records, n * DAPOL_RECORD_BYTES, cudaMemcpyHostToDevice, st);
    int rc = tree_build_dev(ctx, hash_id, height, n, d_idx, nullptr, nullptr, pad_seed, pad_base, out, nullptr, reinterpret_cast<const uint32_t *>(d_rec));
    dfree(mem, st);
    return rc;
}
extern "C" int dapol_tree_attach_top(dapol_tree *t, const dapol_tree *top, uint64_t prefix) {
    if (!t || !top || t->ctx != top->ctx || t->hash_id != top->hash_id || t->height + top->height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_BAD_ARG;
    if (top->height < 64 && (prefix >> top->height)) return DAPOL_ERR_BAD_ARG;
    t->top = top;
    t->prefix = prefix;
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ persistence
namespace {
constexpr char TREE_MAGIC[8] = {'D', 'A', 'P', 'O', 'L', 'T', '0', '1'};
constexpr size_t IO_CHUNK = 64u << 20;
struct TreeFileHeader {  // little-endian, 64 bytes
    char magic[8];
    uint32_t version;
    int32_t hash_id, height;
    uint32_t flags;  // bit 0: id -> leaf-index map present; bit 1: the leaves carry id / salt hashes
    uint64_t n_leaves, T, n_pads, pos_words;
    uint64_t reserved;
};
static_assert(sizeof(TreeFileHeader) == 64, "tree file header layout");
struct PinnedBuf {
    void *p = nullptr;
    PinnedBuf() { if (cudaMallocHost(&p, IO_CHUNK) != cudaSuccess) p = nullptr; }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};
// device array -> file / file -> device array through a pinned staging buffer
int dev_to_file(FILE *f, const void *d, size_t bytes, PinnedBuf &buf, cudaStream_t st) {
    const uint8_t *src = static_cast<const uint8_t *>(d);
    for (size_t off = 0; off < bytes; off += IO_CHUNK) {
        size_t n = std::min(IO_CHUNK, bytes - off);
        CUDA_TRY(cudaMemcpyAsync(buf.p, src + off, n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (fwrite(buf.p, 1, n, f) != n) return DAPOL_ERR_IO;
    }
    return DAPOL_OK;
}
int file_to_dev(FILE *f, void *d, size_t bytes, PinnedBuf &buf, cudaStream_t st) {
    uint8_t *dst = static_cast<uint8_t *>(d);
    for (size_t off = 0; off < bytes; off += IO_CHUNK) {
        size_t n = std::min(IO_CHUNK, bytes - off);
        if (fread(buf.p, 1, n, f) != n) return DAPOL_ERR_IO;
        CUDA_TRY(cudaMemcpyAsync(dst + off, buf.p, n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return DAPOL_OK;
}
// words of the slot-map allocation (tree_build_dev's layout: 64-word aligned pieces for levels 1..H)
uint64_t pos_words_of(const std::vector<uint64_t> &n_real, int H, std::vector<uint64_t> *pos_off) {
    std::vector<uint64_t> off(H + 2, 0);
    for (int h = 1; h <= H; h++) off[h + 1] = off[h] + ((n_real[h] + 63) & ~63ull);
    if (pos_off) *pos_off = off;
    return off[H + 1] + 64;
}
}  // namespace

// A loaded file is data from outside: before any kernel indexes the node store through the slot maps, check that the maps are what
// the build would have produced -- every slot inside its level and increasing, indexes increasing inside a level, the two children
// of a parent adjacent with the parent's index = child index >> 1.  One thread per real node of a level.
__global__ void k_validate_level(uint64_t n_real, uint64_t level_n, const uint32_t *pos, const uint64_t *idx_level, const uint64_t *idx_children,
                                 uint64_t children_n, int is_root_level, int *bad) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_real) return;
    uint64_t slot = is_root_level ? 0 : pos[j];
    if (slot >= level_n || (!is_root_level && j && pos[j - 1] >= slot)) { *bad = 1; return; }
    if (idx_children) {  // children of the j-th real node sit at 2j, 2j + 1 of the level below
        if (2 * j + 1 >= children_n) { *bad = 1; return; }
        uint64_t l = idx_children[2 * j], r = idx_children[2 * j + 1];
        if ((l ^ 1) != r || (l & 1) || (l >> 1) != idx_level[slot] || (j && idx_children[2 * j - 1] >= l)) *bad = 1;
    }
}
static int validate_loaded_tree(dapol_ctx *ctx, const dapol_tree *t) {
    cudaStream_t st = ctx->stream;
    const int H = t->height;
    int *d_bad = reinterpret_cast<int *>(ctx->scratch), bad = 0;
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4, st));
    for (int h = 0; h <= H; h++) {
        if (t->n_real[h] == 0 || t->n_real[h] > t->level_n[h] || (h < H && t->level_n[h + 1] != 2 * t->n_real[h])) return DAPOL_ERR_IO;
        k_validate_level<<<grid_for(t->n_real[h], 256), 256, 0, st>>>(t->n_real[h], t->level_n[h], h ? t->pos[h] : nullptr, t->ns.idx + t->level_off[h],
                                                                      h < H ? t->ns.idx + t->level_off[h + 1] : nullptr, h < H ? t->level_n[h + 1] : 0, h == 0, d_bad);
        ctx->launches++;
    }
    CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return bad ? DAPOL_ERR_IO : DAPOL_OK;
}

extern "C" int dapol_tree_save(const dapol_tree *t, const char *path) {
    if (!t || !path) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(t->ctx->device));
    cudaStream_t st = t->ctx->stream;
    CUDA_TRY(cudaStreamSynchronize(st));
    PinnedBuf buf;
    if (!buf.p) { g_cuda_err = "cudaMallocHost failed"; return DAPOL_ERR_CUDA; }
    FILE *f = fopen(path, "wb");
    if (!f) return DAPOL_ERR_IO;
    const int H = t->height;
    const uint64_t T = t->T;
    TreeFileHeader h = {};
    memcpy(h.magic, TREE_MAGIC, 8);
    const bool save_map = t->leaf_index_of && t->index_map_n == 0;  // a shard's local slice of the map is not saved
    h.version = 1; h.hash_id = t->hash_id; h.height = H; h.flags = (save_map ? 1u : 0u) | (t->custom_leaf_hashes ? 2u : 0u);
    h.n_leaves = t->n_leaves; h.T = T; h.n_pads = t->n_pads; h.pos_words = pos_words_of(t->n_real, H, nullptr);
    int rc = DAPOL_OK;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    for (int l = 0; l <= H && ok; l++) {
        uint64_t row[4] = {t->level_off[l], t->level_n[l], t->n_real[l], t->npads[l]};
        ok = fwrite(row, sizeof row, 1, f) == 1;
    }
    ok = ok && fwrite(t->root_ext, sizeof t->root_ext, 1, f) == 1;
    if (!ok) rc = DAPOL_ERR_IO;
    const struct { const void *p; size_t bytes; } parts[] = {
        {t->ns.idx, T * 8}, {t->ns.v, T * 8}, {t->ns.r, T * 32}, {t->ns.comc, T * 32}, {t->ns.hash, T * 32}, {t->ns.is_pad, T},
        {t->pos_all, h.pos_words * 4}, {t->leaf_index_of, save_map ? t->n_leaves * 8 : 0}, {t->ns.hash_hi, t->ns.hash_hi ? T * 32 : 0}};
    for (const auto &pt : parts)
        if (rc == DAPOL_OK && pt.bytes) rc = dev_to_file(f, pt.p, pt.bytes, buf, st);
    if (rc == DAPOL_OK && fwrite(TREE_MAGIC, 8, 1, f) != 1) rc = DAPOL_ERR_IO;  // trailer: a truncated file does not load
    if (fclose(f) != 0 && rc == DAPOL_OK) rc = DAPOL_ERR_IO;
    return rc;
}
extern "C" int dapol_tree_load(dapol_ctx *ctx, const char *path, dapol_tree **out) {
    if (!ctx || !path || !out) return DAPOL_ERR_BAD_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PinnedBuf buf;
    if (!buf.p) { g_cuda_err = "cudaMallocHost failed"; return DAPOL_ERR_CUDA; }
    FILE *f = fopen(path, "rb");
    if (!f) return DAPOL_ERR_IO;
    TreeFileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, TREE_MAGIC, 8) != 0 || h.version != 1 || h.height < 0 || h.height > DAPOL_MAX_TREE_HEIGHT ||
        (h.hash_id != DAPOL_HASH_BLAKE3 && h.hash_id != DAPOL_HASH_BLAKE2S && h.hash_id != DAPOL_HASH_BLAKE2B) || h.T == 0 || h.n_leaves == 0 || h.n_leaves >= (1ull << 31)) {
        fclose(f);
        return DAPOL_ERR_IO;
    }
    const int H = h.height;
    dapol_tree *t = new dapol_tree();
    t->ctx = ctx; t->hash_id = h.hash_id; t->height = H; t->n_leaves = h.n_leaves; t->T = h.T; t->n_pads = h.n_pads;
    t->custom_leaf_hashes = (h.flags & 2u) != 0;
    t->level_off.assign(H + 1, 0); t->level_n.assign(H + 1, 0); t->n_real.assign(H + 1, 0); t->npads.assign(H + 1, 0);
    t->pos.assign(H + 1, nullptr);
    int rc = DAPOL_OK;
    uint64_t total = 0;
    for (int l = 0; l <= H; l++) {
        uint64_t row[4];
        if (fread(row, sizeof row, 1, f) != 1) { rc = DAPOL_ERR_IO; break; }
        t->level_off[l] = row[0]; t->level_n[l] = row[1]; t->n_real[l] = row[2]; t->npads[l] = row[3];
        if (row[0] != total || row[2] > row[1]) { rc = DAPOL_ERR_IO; break; }  // levels are laid out back to back
        total += row[1];
    }
    std::vector<uint64_t> pos_off;
    if (rc == DAPOL_OK && (total != h.T || t->n_real[H] != h.n_leaves || pos_words_of(t->n_real, H, &pos_off) != h.pos_words)) rc = DAPOL_ERR_IO;
    if (rc == DAPOL_OK && fread(t->root_ext, sizeof t->root_ext, 1, f) != 1) rc = DAPOL_ERR_IO;
    const uint64_t T = h.T;
    auto alloc = [&](auto **p, size_t bytes) {
        if (rc != DAPOL_OK) return;
        if (dmalloc(p, bytes, st) != cudaSuccess) { g_cuda_err = "tree load: out of device memory"; cudaGetLastError(); rc = DAPOL_ERR_CUDA; }
    };
    alloc(&t->ns.idx, T * 8); alloc(&t->ns.v, T * 8); alloc(&t->ns.r, T * 32); alloc(&t->ns.comc, T * 32); alloc(&t->ns.hash, T * 32);
    alloc(&t->ns.is_pad, T); alloc(&t->pos_all, h.pos_words * 4);
    if (h.flags & 1u) alloc(&t->leaf_index_of, h.n_leaves * 8);
    const bool b2b = h.hash_id == DAPOL_HASH_BLAKE2B;
    if (b2b) alloc(&t->ns.hash_hi, T * 32);
    const struct { void *p; size_t bytes; } parts[] = {
        {t->ns.idx, T * 8}, {t->ns.v, T * 8}, {t->ns.r, T * 32}, {t->ns.comc, T * 32}, {t->ns.hash, T * 32}, {t->ns.is_pad, T},
        {t->pos_all, h.pos_words * 4}, {t->leaf_index_of, (h.flags & 1u) ? h.n_leaves * 8 : 0}, {t->ns.hash_hi, b2b ? T * 32 : 0}};
    for (const auto &pt : parts)
        if (rc == DAPOL_OK && pt.bytes) rc = file_to_dev(f, pt.p, pt.bytes, buf, st);
    char tail[8];
    if (rc == DAPOL_OK && (fread(tail, 8, 1, f) != 1 || memcmp(tail, TREE_MAGIC, 8) != 0)) rc = DAPOL_ERR_IO;
    fclose(f);
    if (rc == DAPOL_OK) {
        if (H == 0) t->pos[0] = t->pos_all;
        for (int l = 1; l <= H; l++) t->pos[l] = t->pos_all + pos_off[l];
        if (dmalloc(&t->d_pos, (H + 1) * sizeof(uint32_t *), st) != cudaSuccess || dmalloc(&t->d_level_off, (H + 1) * 8, st) != cudaSuccess ||
            cudaMemcpyAsync(t->d_pos, t->pos.data(), (H + 1) * sizeof(uint32_t *), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(t->d_level_off, t->level_off.data(), (H + 1) * 8, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(t->root_comc, t->ns.comc, 32, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(t->root_hash, t->ns.hash, 32, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            (b2b && cudaMemcpyAsync(t->root_hash + 8, t->ns.hash_hi, 32, cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            g_cuda_err = "tree load: device tables"; cudaGetLastError(); rc = DAPOL_ERR_CUDA;
        }
    }
    if (rc == DAPOL_OK) rc = validate_loaded_tree(ctx, t);  // a corrupted or crafted file must not make later kernels read out of bounds
    if (rc != DAPOL_OK) { dapol_tree_destroy(t); return rc; }
    *out = t;
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ accessors
extern "C" int dapol_tree_height(const dapol_tree *t) { return t ? t->height : -1; }
extern "C" int dapol_tree_hash_id(const dapol_tree *t) { return t ? t->hash_id : -1; }
extern "C" int dapol_digest_len(int hash_id) {
    return hash_id == DAPOL_HASH_BLAKE2B ? 64 : (hash_id == DAPOL_HASH_BLAKE3 || hash_id == DAPOL_HASH_BLAKE2S) ? 32 : 0;
}
extern "C" uint64_t dapol_tree_num_nodes(const dapol_tree *t) { return t ? t->T : 0; }
extern "C" uint64_t dapol_tree_num_padding(const dapol_tree *t) { return t ? t->n_pads : 0; }
extern "C" uint64_t dapol_tree_level_size(const dapol_tree *t, int level) {
    return (t && level >= 0 && level <= t->height) ? t->level_n[level] : 0;
}
extern "C" int dapol_tree_level_copy(const dapol_tree *t, int level, uint64_t *idx, uint64_t *values, uint8_t *blindings, uint8_t *coms,
                                     uint8_t *hashes, uint8_t *is_padding) {
    if (!t || level < 0 || level > t->height) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(t->ctx->device));
    uint64_t o = t->level_off[level], n = t->level_n[level];
    if (idx) CUDA_TRY(cudaMemcpy(idx, t->ns.idx + o, n * 8, cudaMemcpyDeviceToHost));
    if (values) CUDA_TRY(cudaMemcpy(values, t->ns.v + o, n * 8, cudaMemcpyDeviceToHost));
    if (blindings) CUDA_TRY(cudaMemcpy(blindings, t->ns.r + 8 * o, n * 32, cudaMemcpyDeviceToHost));
    if (coms) CUDA_TRY(cudaMemcpy(coms, t->ns.comc + 8 * o, n * 32, cudaMemcpyDeviceToHost));
    if (hashes && !t->ns.hash_hi) CUDA_TRY(cudaMemcpy(hashes, t->ns.hash + 8 * o, n * 32, cudaMemcpyDeviceToHost));
    if (hashes && t->ns.hash_hi) {  // 64-byte digests: the halves are kept apart on the device
        CUDA_TRY(cudaMemcpy2D(hashes, 64, t->ns.hash + 8 * o, 32, 32, n, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy2D(hashes + 32, 64, t->ns.hash_hi + 8 * o, 32, 32, n, cudaMemcpyDeviceToHost));
    }
    if (is_padding) CUDA_TRY(cudaMemcpy(is_padding, t->ns.is_pad + o, n, cudaMemcpyDeviceToHost));
    return DAPOL_OK;
}
extern "C" int dapol_tree_root(const dapol_tree *t, uint8_t com[32], uint8_t hash[32], uint64_t *value, uint8_t blinding[32]) {
    return dapol_tree_level_copy(t, 0, nullptr, value, blinding, com, hash, nullptr);
}
extern "C" int dapol_tree_leaf_index_of(const dapol_tree *t, uint64_t input_pos, uint64_t *leaf_idx) {
    if (!t || !leaf_idx) return DAPOL_ERR_BAD_ARG;
    const uint64_t first = t->index_map_first, cnt = t->index_map_n ? t->index_map_n : t->n_leaves;
    if (!t->leaf_index_of || input_pos < first || input_pos - first >= cnt) return DAPOL_ERR_NOT_FOUND;
    CUDA_TRY(cudaSetDevice(t->ctx->device));
    CUDA_TRY(cudaMemcpy(leaf_idx, t->leaf_index_of + (input_pos - first), 8, cudaMemcpyDeviceToHost));
    return DAPOL_OK;
}
// device-resident path extraction (shared with the inclusion-proof path): all outputs are device pointers
int dapol_tree_paths_dev(const dapol_tree *t, uint64_t k, const uint64_t *d_leaf_idx, uint64_t *d_v, uint32_t *d_r, uint32_t *d_c, uint32_t *d_h,
                         uint32_t *d_lc, uint32_t *d_lh, int *d_not_found, uint32_t *d_hh, uint32_t *d_lhh) {
    dapol_ctx *ctx = t->ctx;
    const int Ht = dapol_total_height(t);
    k_paths<<<grid_for(k, 128), 128, 0, ctx->stream>>>(k, d_leaf_idx, 0, t->top != nullptr, t->prefix, t->ns, t->d_level_off, t->level_n[t->height],
                                                       t->d_pos, t->height, Ht, 0, d_v, d_r, d_c, d_h, d_lc, d_lh, d_not_found, d_hh, d_lhh);
    ctx->launches++;
    if (t->top && t->top->height > 0) {  // the upper levels come from the (replicated) top tree: siblings of this shard's root
        const dapol_tree *u = t->top;
        k_paths<<<grid_for(k, 128), 128, 0, ctx->stream>>>(k, nullptr, t->prefix, 0, 0, u->ns, u->d_level_off, u->level_n[u->height], u->d_pos,
                                                           u->height, Ht, t->height, d_v, d_r, d_c, d_h, nullptr, nullptr, d_not_found, d_hh, nullptr);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}
extern "C" int dapol_tree_paths(const dapol_tree *t, uint64_t k, const uint64_t *leaf_idx, uint64_t *values, uint8_t *blindings,
                                uint8_t *coms, uint8_t *hashes, uint8_t *leaf_coms, uint8_t *leaf_hashes) {
    if (!t || !leaf_idx || !values || !blindings || !coms || !hashes || k == 0) return DAPOL_ERR_BAD_ARG;
    dapol_ctx *ctx = t->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint64_t H = (uint64_t)dapol_total_height(t), kh = k * (H ? H : 1);
    uint8_t *mem = nullptr;
    Arena ar;
    const bool b2b = t->ns.hash_hi != nullptr;  // 64-byte digests: hashes / leaf_hashes are 64 bytes per node
    ar.size = Arena::need(k, 8) + Arena::need(kh, 8) + 4 * Arena::need(kh, 32) + 3 * Arena::need(k, 32) + 256;
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_li = ar.take<uint64_t>(k), *d_v = ar.take<uint64_t>(kh);
    uint32_t *d_r = ar.take<uint32_t>(kh * 8), *d_c = ar.take<uint32_t>(kh * 8), *d_h = ar.take<uint32_t>(kh * 8), *d_hh = ar.take<uint32_t>(kh * 8);
    uint32_t *d_lc = ar.take<uint32_t>(k * 8), *d_lh = ar.take<uint32_t>(k * 8), *d_lhh = ar.take<uint32_t>(k * 8);
    int *d_nf = ar.take<int>(1), nf = 0;
    cudaMemsetAsync(d_nf, 0, 4, st);
    cudaMemcpyAsync(d_li, leaf_idx, k * 8, cudaMemcpyHostToDevice, st);
    int rc = dapol_tree_paths_dev(t, k, d_li, d_v, d_r, d_c, d_h, d_lc, d_lh, d_nf, b2b ? d_hh : nullptr, b2b ? d_lhh : nullptr);
    if (rc) { dfree(mem, st); return rc; }
    cudaMemcpyAsync(&nf, d_nf, 4, cudaMemcpyDeviceToHost, st);
    if (H) {
        cudaMemcpyAsync(values, d_v, kh * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(blindings, d_r, kh * 32, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(coms, d_c, kh * 32, cudaMemcpyDeviceToHost, st);
        if (!b2b) cudaMemcpyAsync(hashes, d_h, kh * 32, cudaMemcpyDeviceToHost, st);
        else {
            cudaMemcpy2DAsync(hashes, 64, d_h, 32, 32, kh, cudaMemcpyDeviceToHost, st);
            cudaMemcpy2DAsync(hashes + 32, 64, d_hh, 32, 32, kh, cudaMemcpyDeviceToHost, st);
        }
    }
    if (leaf_coms) cudaMemcpyAsync(leaf_coms, d_lc, k * 32, cudaMemcpyDeviceToHost, st);
    if (leaf_hashes && !b2b) cudaMemcpyAsync(leaf_hashes, d_lh, k * 32, cudaMemcpyDeviceToHost, st);
    if (leaf_hashes && b2b) {
        cudaMemcpy2DAsync(leaf_hashes, 64, d_lh, 32, 32, k, cudaMemcpyDeviceToHost, st);
        cudaMemcpy2DAsync(leaf_hashes + 32, 64, d_lhh, 32, 32, k, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    dfree(mem, st);
    CUDA_TRY(e);
    return nf ? DAPOL_ERR_NOT_FOUND : DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ micro entry points
extern "C" int dapol_commit_batch(dapol_ctx *ctx, uint64_t n, const uint64_t *values, const uint8_t *blindings, uint8_t *coms) {
    if (!ctx || !values || !blindings || !coms || n == 0) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(n, 8) + 2 * Arena::need(n, 32);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_v = ar.take<uint64_t>(n);
    uint32_t *d_b = ar.take<uint32_t>(8 * n), *d_o = ar.take<uint32_t>(8 * n);
    cudaMemcpyAsync(d_v, values, n * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_b, blindings, n * 32, cudaMemcpyHostToDevice, st);
    unsigned g = grid_for(n, 128);
    switch (ctx->W) {
#define W_CASE(w) case w: k_commit<w><<<g, 128, 0, st>>>(n, d_v, d_b, d_o, ctx->tab_b, ctx->tab_bbl); break;
        DAPOL_W_CASES(W_CASE)
#undef W_CASE
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(coms, d_o, n * 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dfree(mem, st);  // released on the error path too
    CUDA_TRY(e);
    return DAPOL_OK;
}

template <typename F>
static int time_kernel(dapol_ctx *ctx, F launch, float *ms_best) {
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(a, ctx->stream));
        launch();
        ctx->launches++;
        CUDA_TRY(cudaEventRecord(b, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms; CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_best = best;
    return DAPOL_OK;
}
extern "C" int dapol_imad_peak(dapol_ctx *ctx, int variant, double *gmac_per_s) {
    if (!ctx || !gmac_per_s || variant < 0 || variant > 2) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    uint32_t *d_out; CUDA_TRY(cudaMalloc(&d_out, 4));
    const int iters = 4096, blocks = prop.multiProcessorCount * 8, threads = 256;
    float ms; int rc;
    if (variant == 0) rc = time_kernel(ctx, [&]() { k_imad_peak<0><<<blocks, threads, 0, ctx->stream>>>(d_out, 0x9e3779b1u, 0x7f4a7c15u, iters); }, &ms);
    else if (variant == 1) rc = time_kernel(ctx, [&]() { k_imad_peak<1><<<blocks, threads, 0, ctx->stream>>>(d_out, 0x9e3779b1u, 0x7f4a7c15u, iters); }, &ms);
    else rc = time_kernel(ctx, [&]() { k_imad_peak<2><<<blocks, threads, 0, ctx->stream>>>(d_out, 0x9e3779b1u, 0x7f4a7c15u, iters); }, &ms);
    cudaFree(d_out);
    if (rc) return rc;
    // instructions per thread: iters * 8 * 8; MAC32 per instruction: variant 0: 1 (32-bit low product only), 1: 1, 2: 0.5
    double instr = (double)blocks * threads * iters * 64.0;
    double macs = variant == 2 ? instr * 0.5 : instr;
    *gmac_per_s = macs / (ms * 1e-3) / 1e9;
    return DAPOL_OK;
}
extern "C" int dapol_fe_bench(dapol_ctx *ctx, int op, double *gop_per_s) {
    if (!ctx || !gop_per_s || op < 0 || op > 1) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    uint32_t *d_out; CUDA_TRY(cudaMalloc(&d_out, 4));
    const int iters = 2048, blocks = prop.multiProcessorCount * 8, threads = 256;
    float ms; int rc;
    if (op == 0) rc = time_kernel(ctx, [&]() { k_fe_bench<0><<<blocks, threads, 0, ctx->stream>>>(d_out, iters); }, &ms);
    else rc = time_kernel(ctx, [&]() { k_fe_bench<1><<<blocks, threads, 0, ctx->stream>>>(d_out, iters); }, &ms);
    cudaFree(d_out);
    if (rc) return rc;
    *gop_per_s = (double)blocks * threads * iters * 2.0 / (ms * 1e-3) / 1e9;
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ lookup by internal id
// Dapol::generate_proof_for_id / generate_proof_batch_for_ids (src/dapol/mod.rs:148-165) look the liability up in id_to_idx_map.
// Here the map is keyed by the audit id D(audit_seed || internal_id) every liability already has (mod.rs:347-353): the ids' 64-bit
// prefixes are sorted once, on the first lookup, and a query is a binary search + an exact 32-byte comparison.
__global__ void k_index_keys(uint64_t n, const uint32_t *audit, uint64_t *keys, uint32_t *iota) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint64_t)audit[8 * i] | ((uint64_t)audit[8 * i + 1] << 32);
    iota[i] = (uint32_t)i;
}
__global__ void k_lookup_ids(uint64_t k, const uint32_t *want /*[k][8]*/, uint64_t n, const uint64_t *keys_sorted, const uint32_t *who, const uint32_t *audit,
                             const uint64_t *leaf_index_of, uint64_t *out_idx, uint8_t *found) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= k) return;
    uint32_t w[8], a[8];
    load8(w, want + 8 * q);
    const uint64_t key = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (keys_sorted[mid] < key) lo = mid + 1; else hi = mid; }
    found[q] = 0; out_idx[q] = 0;
    for (; lo < n && keys_sorted[lo] == key; lo++) {
        const uint64_t u = who[lo];
        load8(a, audit + 8 * u);
        uint32_t d = 0;
        for (int i = 0; i < 8; i++) d |= a[i] ^ w[i];
        if (d == 0) { found[q] = 1; out_idx[q] = leaf_index_of[u]; return; }
    }
}
static int ensure_id_index(const dapol_tree *t) {
    std::lock_guard<std::mutex> lk(t->index_mu);
    if (t->akey_sorted) return DAPOL_OK;
    dapol_ctx *ctx = t->ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t n = t->index_map_n ? t->index_map_n : t->n_leaves;
    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *iota = nullptr, *who = nullptr;
    uint8_t *tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_sorted, iota, who, (int)n, 0, 64, st);
    cudaError_t e = dmalloc(&keys, n * 8, st);
    if (e == cudaSuccess) e = dmalloc(&keys_sorted, n * 8, st);
    if (e == cudaSuccess) e = dmalloc(&iota, n * 4, st);
    if (e == cudaSuccess) e = dmalloc(&who, n * 4, st);
    if (e == cudaSuccess) e = dmalloc(&tmp, tb, st);
    if (e == cudaSuccess) {
        k_index_keys<<<grid_for(n, 256), 256, 0, st>>>(n, t->audit_ids, keys, iota);
        cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_sorted, iota, who, (int)n, 0, 64, st);
        ctx->launches += 2;
        e = cudaStreamSynchronize(st);
    }
    dfree(keys, st); dfree(iota, st); dfree(tmp, st);
    if (e != cudaSuccess) { dfree(keys_sorted, st); dfree(who, st); CUDA_TRY(e); }
    t->akey_sorted = keys_sorted;
    t->akey_who = who;
    return DAPOL_OK;
}
extern "C" int dapol_tree_index_of_batch(const dapol_tree *t, uint64_t k, const uint8_t *id_blob, const uint64_t *id_off, uint64_t *leaf_idx, uint8_t *found) {
    if (!t || !id_off || !leaf_idx || !found || !k) return DAPOL_ERR_BAD_ARG;
    if (!t->audit_ids || !t->leaf_index_of) { memset(found, 0, k); return DAPOL_ERR_NOT_FOUND; }  // not built from liabilities: no id map (reference: empty map)
    dapol_ctx *ctx = t->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_id_index(t);
    if (rc) return rc;
    // audit ids of the queries on the host (the same D the tree was built with; ids of any length)
    std::vector<uint32_t> want(8 * k);
    for (uint64_t q = 0; q < k; q++) {
        if (id_off[q + 1] < id_off[q] || id_off[q + 1] - id_off[q] > 0xffffffffull) return DAPOL_ERR_BAD_ARG;
        dapol_hasher hs;
        hasher_init(hs, t->hash_id);
        if (!t->audit_seed.empty()) hasher_update(hs, t->audit_seed.data(), (uint32_t)t->audit_seed.size());
        if (id_off[q + 1] > id_off[q]) hasher_update(hs, id_blob + id_off[q], (uint32_t)(id_off[q + 1] - id_off[q]));
        if (hasher_final(hs, &want[8 * q])) return DAPOL_ERR_BAD_ARG;
    }
    cudaStream_t st = ctx->stream;
    const uint64_t n = t->index_map_n ? t->index_map_n : t->n_leaves;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(k, 32) + Arena::need(k, 8) + Arena::need(k, 1);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint32_t *d_want = ar.take<uint32_t>(8 * k);
    uint64_t *d_out = ar.take<uint64_t>(k);
    uint8_t *d_found = ar.take<uint8_t>(k);
    cudaMemcpyAsync(d_want, want.data(), k * 32, cudaMemcpyHostToDevice, st);
    k_lookup_ids<<<grid_for(k, 128), 128, 0, st>>>(k, d_want, n, t->akey_sorted, t->akey_who, t->audit_ids, t->leaf_index_of, d_out, d_found);
    ctx->launches++;
    cudaMemcpyAsync(leaf_idx, d_out, k * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(found, d_found, k, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    dfree(mem, st);
    CUDA_TRY(e);
    for (uint64_t q = 0; q < k; q++) if (!found[q]) return DAPOL_ERR_NOT_FOUND;  // reference: None if ANY id is unknown (mod.rs:155-165)
    return DAPOL_OK;
}
extern "C" int dapol_tree_index_of(const dapol_tree *t, const uint8_t *internal_id, uint64_t len, uint64_t *leaf_idx) {
    if (!leaf_idx) return DAPOL_ERR_BAD_ARG;
    const uint64_t off[2] = {0, len};
    uint8_t found = 0;
    return dapol_tree_index_of_batch(t, 1, internal_id, off, leaf_idx, &found);
}
