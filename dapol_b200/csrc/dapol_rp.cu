// Range proofs of the DAPOL+ hot path on the GPU: batched Bulletproofs prover and verifier (sm_100a), the generator
// window tables they run on, and the C-ABI entry points (include/dapol_b200.h).  Per-thread bodies: rp_kernels.cuh.
// Replaces /root/reference/src/range/mod.rs:48-119 (generate_/verify_{single,aggregated}_range_proof).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include "dapol_internal.h"
#include "rp_kernels.cuh"

#define RP_MSM_MAX_T 128
// minimum resident CTAs per SM the MSM kernels are compiled for (register cap = 65536 / (128 * RP_MSM_MINB))
#ifndef RP_MSM_MINB
#define RP_MSM_MINB 4      // single-warp .. 64-thread CTAs, called products: 128 registers, 16 warps per SM (m = 1: MSM passes 26.5 -> 25.1 ms)
#endif
#ifndef RP_MSM_MINB_INL
#define RP_MSM_MINB_INL 3  // 128-thread CTAs, inlined products
#endif
#ifndef RP_INL_MIN_T
#define RP_INL_MIN_T 128  // CTAs of at least this many threads use the MSM kernels with inlined products
#endif

// ------------------------------------------------------------------------------------------------ CTA reductions
// Sum of the per-thread partial points of a CTA, result in thread 0.  Inside a warp the partials move by shuffles; only
// the (<= 8) warp sums go through shared memory (1 KB), so the resident CTAs per SM are not capped by shared memory --
// the small shapes (N = 64: one or two warps per CTA) need many CTAs per SM to keep the multiply pipe busy.
#define RP_REDUCE_SH_WORDS (RP_MSM_MAX_T / 32 * 32)
__device__ __forceinline__ void shfl_down_ge(ge &o, const ge &a, int d) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        o.X.v[i] = __shfl_down_sync(0xffffffffu, a.X.v[i], d); o.Y.v[i] = __shfl_down_sync(0xffffffffu, a.Y.v[i], d);
        o.Z.v[i] = __shfl_down_sync(0xffffffffu, a.Z.v[i], d); o.T.v[i] = __shfl_down_sync(0xffffffffu, a.T.v[i], d);
    }
}
__device__ __forceinline__ void block_reduce_ge(ge &acc, uint32_t *sh /*[blockDim.x / 32][32]*/) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) {
        ge o;
        shfl_down_ge(o, acc, d);
        ge_add(acc, acc, o);  // lanes >= d compute a sum nobody reads
    }
    if (nwarps == 1) return;
    if (lane == 0) rp_store_ext(sh + 32 * warp, acc);
    __syncthreads();
    if (warp == 0) {
        if (lane < nwarps) rp_load_ext(acc, sh + 32 * lane); else ge_identity(acc);
#pragma unroll 1
        for (int d = (int)nwarps >> 1; d > 0; d >>= 1) {
            ge o;
            shfl_down_ge(o, acc, d);
            ge_add(acc, acc, o);
        }
    }
}
__device__ __forceinline__ void block_reduce_sc(sc &acc, uint32_t *sh /*[blockDim.x][8]*/) {
    uint32_t tid = threadIdx.x;
    store8(sh + 8 * tid, acc.v);
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            sc o;
            load8(o.v, sh + 8 * (tid + s));
            sc_add(acc, acc, o);
            store8(sh + 8 * tid, acc.v);
        }
        __syncthreads();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ table kernels
__global__ void k_rp_gen_chain(int mcap, uint32_t *uniform) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * mcap) rp_gen_chain_body((uint32_t)(t % mcap), t / mcap, uniform + (uint64_t)t * 64 * 16);
}
__global__ void __launch_bounds__(64) k_rp_gen_point(uint64_t n, const uint32_t *uniform, uint32_t *ext) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) rp_gen_point_body(g, uniform, ext);
    if (g == n) { ge p; ge_basepoint(p); rp_store_ext(ext + 32 * n, p); }
    if (g == n + 1) { ge p; ge_bblinding(p); rp_store_ext(ext + 32 * (n + 1), p); }
}
template <int W>
__global__ void __launch_bounds__(64) k_rp_tab_windows(uint64_t n, const uint32_t *ext, uint32_t *wb) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) rp_tab_windows_body<W>(g, ext, wb);
}
template <int W>
__global__ void __launch_bounds__(128) k_rp_tab_chunk(uint64_t items, const uint32_t *wb, ge_niels *tab) {
    uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it < items) rp_tab_chunk_body<W>(it, wb, tab);
}

// windows of the generator tables that are instantiated (dapol_ctx_set_rangeproof_window; 0 = pick by HBM budget)
#define RP_W_CASES(X) X(8) X(12) X(13) X(14) X(15) X(16)
static size_t rp_table_bytes(int W, int mcap) {
    return (128ull * mcap + 2) * (size_t)(253 / W + 1) * (1ull << (W - 1)) * sizeof(ge_niels);
}
template <int W>
static int rp_build_tables(dapol_ctx *ctx, int mcap) {
    constexpr int NW = 253 / W + 1;
    constexpr uint64_t HALF = 1ull << (W - 1);
    constexpr uint64_t C = HALF < RP_TAB_CHUNK ? HALF : RP_TAB_CHUNK;
    cudaStream_t st = ctx->stream;
    uint64_t ngen = 128ull * mcap, nb = ngen + 2;
    uint32_t *uniform = nullptr, *ext = nullptr, *wb = nullptr;
    ge_niels *tab = nullptr;
    if (cudaMalloc(&tab, nb * NW * HALF * sizeof(ge_niels)) != cudaSuccess || cudaMalloc(&uniform, 2ull * mcap * 64 * 64) != cudaSuccess ||
        cudaMalloc(&ext, nb * 128) != cudaSuccess || cudaMalloc(&wb, nb * NW * 128) != cudaSuccess) {
        dapol_cuda_err() = "range-proof generator tables: out of device memory";
        cudaFree(tab); cudaFree(uniform); cudaFree(ext); cudaFree(wb);
        return DAPOL_ERR_CUDA;
    }
    k_rp_gen_chain<<<grid_for(2 * mcap, 32), 32, 0, st>>>(mcap, uniform);
    k_rp_gen_point<<<grid_for(nb, 64), 64, 0, st>>>(ngen, uniform, ext);
    k_rp_tab_windows<W><<<grid_for(nb, 64), 64, 0, st>>>(nb, ext, wb);
    uint64_t items = nb * NW * (HALF / C);
    k_rp_tab_chunk<W><<<grid_for(items, 128), 128, 0, st>>>(items, wb, tab);
    ctx->launches += 4;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(uniform); cudaFree(ext); cudaFree(wb);
    ctx->rp_tab = tab;
    ctx->rp_mcap = mcap;
    return DAPOL_OK;
}
static int rp_ensure_tables(dapol_ctx *ctx, int m) {
    if (ctx->rp_mcap >= m) return DAPOL_OK;
    int mc = 1;
    while (mc < m) mc <<= 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
    // the old tables go first: they are being replaced, and their HBM may be what the new ones need
    cudaStreamSynchronize(ctx->stream);
    if (ctx->rp_tab) { cudaFree(ctx->rp_tab); ctx->rp_tab = nullptr; ctx->rp_mcap = 0; }
    auto build = [&](int w) {
        switch (w) {
#define W_CASE(w_) case w_: return rp_build_tables<w_>(ctx, mc);
            RP_W_CASES(W_CASE)
#undef W_CASE
        }
        return (int)DAPOL_ERR_BAD_ARG;
    };
    int rc = DAPOL_ERR_CUDA;
    if (ctx->rp_W_auto) {  // widest window whose tables fit the HBM budget: fewer additions per scalar multiplication
        // HBM budget of the tables: what the caller set (dapol_ctx_set_rangeproof_table_budget), else 70 % of the memory that is
        // free or cached-but-unused in the device's stream-ordered pool, at most 128 GB (the node store of a 2^21-user shard is
        // 16 GB, the batch scratch 6 GB).  The shared default pool is NOT trimmed up front (a co-tenant, e.g. torch, may rely on
        // its cache); only when an allocation that fits the budget fails is the pool's unused cache released, once.
        cudaMemPool_t pool = nullptr;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        uint64_t reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        }
        const size_t avail = free_b + (size_t)(reserved > used ? reserved - used : 0);
        const size_t budget = ctx->rp_budget ? ctx->rp_budget : std::min<size_t>(128ull << 30, avail / 10 * 7);
        static const int widths[] = {16, 15, 14, 13, 12, 8};
        bool trimmed = false;
        for (int cand : widths) {
            if (cand != 8 && rp_table_bytes(cand, mc) > budget) continue;
            ctx->rp_W = cand;
            rc = build(cand);
            if (rc == DAPOL_ERR_CUDA && !trimmed && pool) {  // the cache of earlier batches is in the way: release it and try this width again
                cudaGetLastError();
                cudaMemPoolTrimTo(pool, 0);
                trimmed = true;
                rc = build(cand);
            }
            if (rc != DAPOL_ERR_CUDA) break;
            cudaGetLastError();  // out of memory after all (fragmentation, another context): next narrower window
        }
        if (rc == DAPOL_OK) ctx->rp_table_bytes = rp_table_bytes(ctx->rp_W, mc);
    } else {
        rc = build(ctx->rp_W);
        if (rc == DAPOL_OK) ctx->rp_table_bytes = rp_table_bytes(ctx->rp_W, mc);
    }
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ctx->rp_last_ms[3], a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return rc;
}

// ------------------------------------------------------------------------------------------------ prover kernels
__global__ void __launch_bounds__(64) k_rp_p0(RpBatch b) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rp_p0_body(b, p);
}
template <int WT>
__global__ void __launch_bounds__(64) k_rp_p1(RpBatch b, const ge_niels *tab_b, const ge_niels *tab_bbl) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * b.m) rp_p1_body<WT>(b, t / b.m, (int)(t % b.m), tab_b, tab_bbl);
}
__global__ void __launch_bounds__(128) k_rp_p2(RpBatch b) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * b.N) rp_p2_body(b, t / b.N, (uint32_t)(t % b.N));
}
// result of CTA (proof x, which y, split z): the point itself, or one of gridDim.z partial sums
__device__ __forceinline__ void store_msm_point(const RpBatch &b, const ge &acc) {
    if (gridDim.z == 1) rp_store_point(b, blockIdx.x, blockIdx.y, acc);
    else rp_store_ext(b.parts + (((uint64_t)blockIdx.x * 2 + blockIdx.y) * RP_SPLIT_MAX + blockIdx.z) * 32, acc);
}
// MSM kernels (CTA per proof).  INL: 128-thread CTAs of the large shapes run the mixed additions with inlined products; the
// single-warp CTAs of the small shapes share the called multiplication (see ge25519.cuh, ge_madd_inl).
template <int W, bool INL>
__global__ void __launch_bounds__(RP_MSM_MAX_T, INL ? RP_MSM_MINB_INL : RP_MSM_MINB) k_rp_p3(RpBatch b) {
    __shared__ uint32_t sh[RP_REDUCE_SH_WORDS];
    ge acc;
    rp_p3_partial<W, INL>(acc, b, blockIdx.x, blockIdx.y, blockIdx.z * blockDim.x + threadIdx.x, gridDim.z * blockDim.x);
    block_reduce_ge(acc, sh);
    if (threadIdx.x == 0) store_msm_point(b, acc);
}
__global__ void __launch_bounds__(64) k_rp_p4(RpBatch b) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rp_p4_body(b, p);
}
// doubling expansion of one vector per proof (CTA per proof): vec[i + 2^s] = vec[i] * mult[s]
__global__ void __launch_bounds__(256) k_rp_expand(RpBatch b, uint32_t *vec, int slot) {
    uint64_t p = blockIdx.x;
    for (int s = 0; s < b.lg; s++) {
        for (uint32_t i = threadIdx.x; i < (1u << s); i += blockDim.x) rp_expand_step(vec + p * b.N * 8, b.mult + (p * 3 + slot) * 32 * 8, s, i);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(RP_MSM_MAX_T) k_rp_p5(RpBatch b) {
    __shared__ uint32_t sh[RP_MSM_MAX_T * 8];
    sc t0, t1, t2;
    uint64_t p = blockIdx.x;
    rp_p5_partial(t0, t1, t2, b, p, threadIdx.x, blockDim.x);
    block_reduce_sc(t0, sh); block_reduce_sc(t1, sh); block_reduce_sc(t2, sh);
    if (threadIdx.x == 0) { rp_st(rp_ch(b, p, CH_T0), t0); rp_st(rp_ch(b, p, CH_T1), t1); rp_st(rp_ch(b, p, CH_T2), t2); }
}
template <int W>
__global__ void __launch_bounds__(64) k_rp_p6(RpBatch b) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * b.K) rp_p6_body<W>(b, t >> 1, (int)(t & 1));
}
__global__ void __launch_bounds__(64) k_rp_p7(RpBatch b) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rp_p7_body(b, p);
}
__global__ void __launch_bounds__(128) k_rp_p8(RpBatch b) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * b.N) rp_p8_body(b, t / b.N, (uint32_t)(t % b.N));
}
__global__ void __launch_bounds__(RP_MSM_MAX_T) k_rp_p9(RpBatch b, int rnd) {
    __shared__ uint32_t sh[RP_MSM_MAX_T * 8];
    sc cl, cr;
    uint64_t p = blockIdx.x;
    rp_p9_partial(cl, cr, b, p, rnd, threadIdx.x, blockDim.x);
    block_reduce_sc(cl, sh); block_reduce_sc(cr, sh);
    if (threadIdx.x == 0) { rp_st(rp_ch(b, p, CH_CL), cl); rp_st(rp_ch(b, p, CH_CR), cr); }
}
template <int W, bool INL>
__global__ void __launch_bounds__(RP_MSM_MAX_T, INL ? RP_MSM_MINB_INL : RP_MSM_MINB) k_rp_p10(RpBatch b, int rnd) {
    __shared__ uint32_t sh[RP_REDUCE_SH_WORDS];
    ge acc;
    rp_p10_partial<W, INL>(acc, b, blockIdx.x, rnd, blockIdx.y, blockIdx.z * blockDim.x + threadIdx.x, gridDim.z * blockDim.x);
    block_reduce_ge(acc, sh);
    if (threadIdx.x == 0) store_msm_point(b, acc);
}
// compressed form of the two points of every proof (thread per (proof, point)), read by the transcript passes
__global__ void __launch_bounds__(64) k_rp_compress_pts(RpBatch b, int per_proof) {
    uint64_t pw = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (per_proof) {  // A/B knob (DAPOL_RP_COMPRESS_SEQ): both points of a proof by one thread, as the transcript passes used to do
        if (pw < b.K) { rp_compress_point_body(b, 2 * pw); rp_compress_point_body(b, 2 * pw + 1); }
        return;
    }
    if (pw < 2 * b.K) rp_compress_point_body(b, pw);
}
// small batches: the S partial sums of a split MSM (grid z) -> the point of (proof, L | R)
__global__ void __launch_bounds__(32) k_rp_sum_parts(RpBatch b, uint32_t S) {
    uint64_t pw = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pw < 2 * b.K) rp_sum_parts_body(b, pw, S);
}
// The smallest shapes in large batches (N <= 32: a warp per MSM gives every lane ONE term and then runs a 5-step shuffle reduction
// of 32 partial points): g = 8 lanes per (proof, L | R), four MSMs per warp, 4 terms per lane and a 3-step segmented reduction.
// Same partial-sum bodies, other (tid, T).
__device__ __forceinline__ void segmented_reduce_ge(ge &acc, uint32_t g) {
#pragma unroll 1
    for (uint32_t d = g >> 1; d > 0; d >>= 1) {
        ge o;
        shfl_down_ge(o, acc, (int)d);
        ge_add(acc, acc, o);
    }
}
template <int W>
__global__ void __launch_bounds__(RP_MSM_MAX_T, RP_MSM_MINB) k_rp_p3g(RpBatch b, uint32_t g) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pw = idx / g;
    ge acc;
    if (pw < 2 * b.K) rp_p3_partial<W, false>(acc, b, pw >> 1, (int)(pw & 1), (uint32_t)(idx % g), g);
    else ge_identity(acc);
    segmented_reduce_ge(acc, g);
    if (pw < 2 * b.K && idx % g == 0) rp_store_point(b, pw >> 1, (int)(pw & 1), acc);
}
template <int W>
__global__ void __launch_bounds__(RP_MSM_MAX_T, RP_MSM_MINB) k_rp_p10g(RpBatch b, int rnd, uint32_t g) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pw = idx / g;
    ge acc;
    if (pw < 2 * b.K) rp_p10_partial<W, false>(acc, b, pw >> 1, rnd, (int)(pw & 1), (uint32_t)(idx % g), g);
    else ge_identity(acc);
    segmented_reduce_ge(acc, g);
    if (pw < 2 * b.K && idx % g == 0) rp_store_point(b, pw >> 1, (int)(pw & 1), acc);
}
__global__ void __launch_bounds__(64) k_rp_p11(RpBatch b, int rnd) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rp_p11_body(b, p, rnd);
}
__global__ void __launch_bounds__(128) k_rp_p12(RpBatch b, int rnd, uint32_t cnt) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * cnt) rp_p12_body(b, t / cnt, rnd, (uint32_t)(t % cnt));
}
// hybrid rounds of the large aggregates (rp_kernels.cuh): materialise the folded generators, variable-base L / R, fold
template <int W>
__global__ void __launch_bounds__(64, 8) k_rp_pm(RpBatch b, int s) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * 2 * RP_FOLD_N) rp_pm_body<W, false>(b, t / (2 * RP_FOLD_N), (int)((t / RP_FOLD_N) & 1), (uint32_t)(t % RP_FOLD_N), s);
}
template <int W>
__global__ void __launch_bounds__(128) k_rp_pv(RpBatch b, int rnd) {
    const uint32_t g = 2u * ((uint32_t)b.N >> rnd);  // terms per (proof, L|R): a power of two <= 32, so groups never straddle a warp
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pw = idx / g;
    uint32_t t = (uint32_t)(idx % g);
    ge acc;
    if (pw < 2 * b.K) rp_pv_partial<W>(acc, b, pw >> 1, rnd, (int)(pw & 1), t, g);
    else ge_identity(acc);
#pragma unroll 1
    for (uint32_t d = g >> 1; d > 0; d >>= 1) {  // segmented reduction inside the group
        ge o;
        shfl_down_ge(o, acc, (int)d);
        ge_add(acc, acc, o);
    }
    if (pw < 2 * b.K && t == 0) rp_store_point(b, pw >> 1, (int)(pw & 1), acc);
}
__global__ void __launch_bounds__(64) k_rp_pf(RpBatch b, int rnd) {
    const uint32_t h = (uint32_t)b.N >> rnd;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * 2 * h) rp_pf_body(b, t / (2 * h), rnd, (int)((t / h) & 1), (uint32_t)(t % h));
}
// ------------------------------------------------------------------------------------------------ verifier kernels
__global__ void __launch_bounds__(64) k_rp_v0(RpBatch b) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rp_v0_body(b, p);
}
__global__ void __launch_bounds__(64) k_rp_v1(RpBatch b) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < b.K * b.vgroups) rp_v1_body(b, t / b.vgroups, (int)(t % b.vgroups));
}
template <int W, bool INL>
__global__ void __launch_bounds__(RP_MSM_MAX_T, INL ? RP_MSM_MINB_INL : RP_MSM_MINB) k_rp_v2(RpBatch b) {
    __shared__ uint32_t sh[RP_REDUCE_SH_WORDS];
    ge acc;
    uint64_t p = blockIdx.x;
    rp_v2_partial<W, INL>(acc, b, p, blockIdx.z * blockDim.x + threadIdx.x, gridDim.z * blockDim.x);
    block_reduce_ge(acc, sh);
    if (threadIdx.x == 0) {
        if (gridDim.z == 1) b.status[p] = b.status[p] && ge_is_identity(acc);
        else rp_store_ext(b.parts + ((p * 2) * RP_SPLIT_MAX + blockIdx.z) * 32, acc);  // small batches: one of gridDim.z partial sums
    }
}
// small batches: the verification equation's sum was split over S CTAs; add the parts and test the identity
__global__ void __launch_bounds__(32) k_rp_v2_finish(RpBatch b, uint32_t S) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.K) return;
    ge acc, q;
    rp_load_ext(acc, b.parts + (p * 2 * RP_SPLIT_MAX) * 32);
#pragma unroll 1
    for (uint32_t z = 1; z < S; z++) { rp_load_ext(q, b.parts + (p * 2 * RP_SPLIT_MAX + z) * 32); ge_add(acc, acc, q); }
    b.status[p] = b.status[p] && ge_is_identity(acc);
}

// ------------------------------------------------------------------------------------------------ host orchestration
static inline int ilog2(uint64_t x) { int l = 0; while ((1ull << l) < x) l++; return l; }
static inline unsigned msm_threads(uint64_t terms) {  // a few terms per thread, 32..RP_MSM_MAX_T threads
    unsigned t = 32;
    while (t < RP_MSM_MAX_T && (uint64_t)t * 4 < terms) t <<= 1;
    return t;
}
static bool rp_shape_ok(int nbits, int m) {
    return (nbits == 8 || nbits == 16 || nbits == 32 || nbits == 64) && m >= 1 && m <= 64 && (m & (m - 1)) == 0;
}
extern "C" uint64_t dapol_rangeproof_size(int nbits, int m) {
    if (!rp_shape_ok(nbits, m)) return 0;
    return 32ull * (9 + 2 * ilog2((uint64_t)nbits * m));
}
// scratch bytes per proof of a batch
static size_t rp_per_proof_bytes(int N, int m, int lg, bool verify) {
    size_t s = sizeof(merlin) + 2 * (size_t)m * 32 + CH_COUNT * 32 + (size_t)m * 32 + 3 * 32 * 32 + 256 + 4 + 4 * 256;
    s += (size_t)N * 32 * (verify ? 2 : 3);          // ypow, svec | vecA, vecB, ypow
    if (!verify) s += 4 * (size_t)(N / 2) * 32 + 2 * RP_FOLD_N * 128;  // cu, cui ping-pong; folded generators
    if (verify) s += (size_t)rp_nvar(lg, m) * (128 + 32 + 8 * 128);  // partial points, scalars, cached multiples
    return s;
}
#define RP_SPLIT_MAX_K 64  // batches up to this many proofs split their MSMs over several CTAs
// CTAs per (proof, L | R) MSM of `terms` terms on T threads: enough CTAs to cover the SMs twice, at least ~2 terms per thread
static unsigned msm_split(uint64_t K, uint64_t terms, unsigned T) {
    if (K > RP_SPLIT_MAX_K) return 1;
    uint64_t s = (2 * 148 + 2 * K - 1) / (2 * K);
    s = std::min<uint64_t>(s, std::max<uint64_t>(1, terms / (2ull * T)));
    return (unsigned)std::min<uint64_t>(s, RP_SPLIT_MAX);
}
struct RpPlan {
    RpBatch b;
    uint8_t *mem = nullptr;
};
// carve the scratch of a batch of K proofs out of one pool allocation
static int rp_plan(dapol_ctx *ctx, RpPlan &pl, int nbits, int m, uint64_t K, bool verify) {
    RpBatch &b = pl.b;
    memset(&b, 0, sizeof b);
    b.nbits = nbits; b.m = m; b.N = nbits * m; b.lg = ilog2((uint64_t)b.N); b.K = K;
    b.plen = 32 * (9 + 2 * b.lg);
    const int nv = rp_nvar(b.lg, m);
    const uint64_t N = b.N;
    Arena ar;
    ar.size = Arena::need(K, sizeof(merlin)) + 3 * Arena::need(K * m, 32) + Arena::need(K * CH_COUNT, 32) + Arena::need(K * 3 * 32, 32) +
              4 * Arena::need(K * N, 32) + 4 * Arena::need(K * (N / 2 + 1), 32) + Arena::need(K * 2, 128) + Arena::need(K * 2, 32) + Arena::need(K * nv, 128) +
              Arena::need(K * nv, 32) + Arena::need(K, b.plen) + Arena::need(K, 4) + (verify ? Arena::need(K * nv * 8, 128) : Arena::need(K * 2 * RP_FOLD_N, 128)) +
              (K <= RP_SPLIT_MAX_K ? Arena::need(K * 2 * RP_SPLIT_MAX, 128) : 0);
    CUDA_TRY(dmalloc(&pl.mem, ar.size, ctx->stream));
    ar.base = pl.mem;
    b.tr = ar.take<merlin>(K);
    b.Vc = ar.take<uint32_t>(K * m * 8); b.blr = ar.take<uint32_t>(K * m * 8); b.zpow = ar.take<uint32_t>(K * m * 8);
    b.chal = ar.take<uint32_t>(K * CH_COUNT * 8);
    b.mult = ar.take<uint32_t>(K * 3 * 32 * 8);
    b.ypow = ar.take<uint32_t>(K * N * 8);
    if (verify) {
        b.svec = ar.take<uint32_t>(K * N * 8);
        b.varpts = ar.take<uint32_t>(K * nv * 32);
        b.varsc = ar.take<uint32_t>(K * nv * 8);
        b.vartab = ar.take<uint32_t>(K * nv * 8 * 32);
        // threads per proof of V1: enough threads to fill the GPU for small batches, one shared doubling chain per thread
        // (a small batch of many-point proofs ends at one point per thread, a large batch at one thread per proof)
        int g = 1;
        while (g < nv && K * g < 65536) g <<= 1;
        while (g * RP_V1_PMAX < nv) g <<= 1;
        b.vgroups = g < nv ? g : nv;
        if (K <= RP_SPLIT_MAX_K) b.parts = ar.take<uint32_t>(K * 2 * RP_SPLIT_MAX * 32);
    } else {
        b.vecA = ar.take<uint32_t>(K * N * 8); b.vecB = ar.take<uint32_t>(K * N * 8);
        for (int i = 0; i < 2; i++) { b.cu[i] = ar.take<uint32_t>(K * (N / 2 + 1) * 8); b.cui[i] = ar.take<uint32_t>(K * (N / 2 + 1) * 8); }
        b.pts = ar.take<uint32_t>(K * 2 * 32);
        b.ptc = ar.take<uint32_t>(K * 2 * 8);
        b.gfold = ar.take<uint32_t>(K * 2 * RP_FOLD_N * 32);
        if (K <= RP_SPLIT_MAX_K) b.parts = ar.take<uint32_t>(K * 2 * RP_SPLIT_MAX * 32);
        b.proof = ar.take<uint32_t>(K * b.plen / 4);
    }
    b.status = ar.take<int>(K);
    const int W = ctx->rp_W;
    uint64_t per = (uint64_t)(253 / W + 1) * (1ull << (W - 1));
    b.tabG = ctx->rp_tab;
    b.tabH = b.tabG + 64ull * ctx->rp_mcap * per;
    b.tabB = b.tabG + 128ull * ctx->rp_mcap * per;
    b.tabBbl = b.tabB + per;
    return DAPOL_OK;
}
// proofs per chunk so that the scratch stays within a fixed HBM budget
static uint64_t rp_chunk(int N, int m, int lg, bool verify, uint64_t K) {
    const size_t budget = 6ull << 30;
    uint64_t c = budget / rp_per_proof_bytes(N, m, lg, verify);
    if (c < 1) c = 1;
    return std::min<uint64_t>(K, c);
}

// device time per kernel class, on the ctx stream.  Classes: TM_OTHER = thread-per-proof / per-element passes; the MSM
// passes are split into TM_P10 (L / R of the table rounds), TM_P3 (A, S), TM_HYB (pm, pv, pf), TM_VER (v1, v2).
enum { TM_P10 = 0, TM_OTHER = 1, TM_P3 = 2, TM_HYB = 3, TM_VER = 4, TM_CLASSES = 5 };
struct EventPair {  // released on every exit path
    cudaEvent_t a = nullptr, b = nullptr;
    EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    EventPair(const EventPair &) = delete;
    EventPair &operator=(const EventPair &) = delete;
};
struct PhaseTimer {
    cudaStream_t st;
    cudaEvent_t e[2];
    float ms[TM_CLASSES] = {};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[TM_CLASSES];
    explicit PhaseTimer(cudaStream_t s) : st(s) {}
    ~PhaseTimer() { collect(); }  // error paths: the events do not leak
    void begin(int which) { cudaEventCreate(&e[0]); cudaEventCreate(&e[1]); cudaEventRecord(e[0], st); cur = which; }
    void end() { cudaEventRecord(e[1], st); spans[cur].push_back({e[0], e[1]}); }
    void collect() {
        for (int w = 0; w < TM_CLASSES; w++) {
            for (auto &s : spans[w]) { float t = 0; cudaEventElapsedTime(&t, s.first, s.second); ms[w] += t; cudaEventDestroy(s.first); cudaEventDestroy(s.second); }
            spans[w].clear();
        }
    }
    void publish(float out[8]) const {
        out[1] = ms[TM_P10] + ms[TM_P3] + ms[TM_HYB] + ms[TM_VER]; out[2] = ms[TM_OTHER];
        out[4] = ms[TM_P10]; out[5] = ms[TM_P3]; out[6] = ms[TM_HYB]; out[7] = ms[TM_VER];
    }
    int cur = 0;
};

template <int W>
static int rp_prove_chunk(dapol_ctx *ctx, RpBatch &b, PhaseTimer &tm) {
    cudaStream_t st = ctx->stream;
    const uint64_t K = b.K, N = b.N;
    const unsigned T = msm_threads(N), TS = msm_threads(2 * N);
    static const int cseq = getenv("DAPOL_RP_COMPRESS_SEQ") ? atoi(getenv("DAPOL_RP_COMPRESS_SEQ")) : 0;
    tm.begin(1);
    k_rp_p0<<<grid_for(K, 64), 64, 0, st>>>(b);
    switch (ctx->W) {
#define W_CASE(w) case w: k_rp_p1<w><<<grid_for(K * b.m, 64), 64, 0, st>>>(b, ctx->tab_b, ctx->tab_bbl); break;
        DAPOL_W_CASES(W_CASE)
#undef W_CASE
        default: return DAPOL_ERR_BAD_ARG;
    }
    k_rp_p2<<<grid_for(K * N, 128), 128, 0, st>>>(b);
    tm.end();
    tm.begin(TM_P3);
    // lanes per MSM of the packed kernels: shapes of at most 32 generators per vector in batches that fill the GPU several times over
    // even at 8 lanes per MSM.  Measured (profiles/r02_variants.txt 9): n = 32, m = 1: MSM passes 40.4 -> 23.7 ms per 32768 proofs;
    // N = 64 and 128 gain nothing (26.1 vs 26.3 ms, 54.3 vs 56.9 ms) and keep a warp per MSM.
    const unsigned GP = (N <= (uint64_t)ctx->rp_pack_max_n && K >= (uint64_t)ctx->rp_pack_min_k && ctx->rp_pack_lanes) ? (unsigned)ctx->rp_pack_lanes : 0;
    const unsigned S3 = (b.parts && !GP) ? msm_split(K, N, TS) : 1, S10 = (b.parts && !GP) ? msm_split(K, N, T) : 1;
    if (GP) k_rp_p3g<W><<<grid_for(2 * K * GP, RP_MSM_MAX_T), RP_MSM_MAX_T, 0, st>>>(b, GP);
    else if (TS >= RP_INL_MIN_T) k_rp_p3<W, true><<<dim3((unsigned)K, 2, S3), TS, 0, st>>>(b);
    else k_rp_p3<W, false><<<dim3((unsigned)K, 2, S3), TS, 0, st>>>(b);
    if (S3 > 1) { k_rp_sum_parts<<<grid_for(2 * K, 32), 32, 0, st>>>(b, S3); ctx->launches++; }
    k_rp_compress_pts<<<grid_for(2 * K, 64), 64, 0, st>>>(b, cseq);
    tm.end();
    tm.begin(1);
    k_rp_p4<<<grid_for(K, 64), 64, 0, st>>>(b);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.ypow, 0);
    k_rp_p5<<<(unsigned)K, T, 0, st>>>(b);
    k_rp_p6<W><<<grid_for(2 * K, 64), 64, 0, st>>>(b);
    k_rp_compress_pts<<<grid_for(2 * K, 64), 64, 0, st>>>(b, cseq);
    k_rp_p7<<<grid_for(K, 64), 64, 0, st>>>(b);
    k_rp_p8<<<grid_for(K * N, 128), 128, 0, st>>>(b);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.ypow, 1);
    tm.end();
    ctx->launches += 12;
    // first round over folded generators (large aggregates), else lg + 1.  Small batches stay on the generator tables: the
    // variable-base chains of the folded rounds (252 doublings each) are latency a handful of proofs cannot hide, while a table
    // round split over the whole GPU is a few additions per thread -- same L_k, R_k, so the same bytes.
    const int sw = S10 > 1 ? b.lg + 1 : rp_switch_round(b.N, b.lg);
    for (int rnd = 1; rnd <= b.lg; rnd++) {
        uint32_t h = (uint32_t)(N >> rnd), cnt = std::max<uint32_t>(h, 1u << (rnd - 1));
        tm.begin(1);
        k_rp_p9<<<(unsigned)K, msm_threads(h), 0, st>>>(b, rnd);
        tm.end();
        tm.begin(rnd >= sw ? TM_HYB : TM_P10);
        if (rnd == sw) { k_rp_pm<W><<<grid_for(K * 2 * RP_FOLD_N, 64), 64, 0, st>>>(b, sw); ctx->launches++; }
        if (rnd >= sw) k_rp_pv<W><<<grid_for(K * 4 * h, 128), 128, 0, st>>>(b, rnd);
        else if (GP) k_rp_p10g<W><<<grid_for(2 * K * GP, RP_MSM_MAX_T), RP_MSM_MAX_T, 0, st>>>(b, rnd, GP);
        else if (T >= RP_INL_MIN_T) k_rp_p10<W, true><<<dim3((unsigned)K, 2, S10), T, 0, st>>>(b, rnd);
        else k_rp_p10<W, false><<<dim3((unsigned)K, 2, S10), T, 0, st>>>(b, rnd);
        if (rnd < sw && S10 > 1) { k_rp_sum_parts<<<grid_for(2 * K, 32), 32, 0, st>>>(b, S10); ctx->launches++; }
        k_rp_compress_pts<<<grid_for(2 * K, 64), 64, 0, st>>>(b, cseq);
        tm.end();
        tm.begin(1);
        k_rp_p11<<<grid_for(K, 64), 64, 0, st>>>(b, rnd);
        k_rp_p12<<<grid_for(K * cnt, 128), 128, 0, st>>>(b, rnd, cnt);
        tm.end();
        ctx->launches += 5;
        if (rnd >= sw && rnd < b.lg) {
            tm.begin(TM_HYB);
            k_rp_pf<<<grid_for(K * 2 * h, 64), 64, 0, st>>>(b, rnd);
            tm.end();
            ctx->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}
template <int W>
static int rp_verify_chunk(dapol_ctx *ctx, RpBatch &b, PhaseTimer &tm) {
    cudaStream_t st = ctx->stream;
    const uint64_t K = b.K, N = b.N;
    const int nv = rp_nvar(b.lg, b.m);
    tm.begin(1);
    k_rp_v0<<<grid_for(K, 64), 64, 0, st>>>(b);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.svec, 2);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.ypow, 1);
    tm.end();
    tm.begin(TM_VER);
    k_rp_v1<<<grid_for(K * b.vgroups, 64), 64, 0, st>>>(b);
    const unsigned TV = msm_threads(2 * N), SV = b.parts ? msm_split(K, 2 * N, TV) : 1;
    if (TV >= RP_INL_MIN_T) k_rp_v2<W, true><<<dim3((unsigned)K, 1, SV), TV, 0, st>>>(b);
    else k_rp_v2<W, false><<<dim3((unsigned)K, 1, SV), TV, 0, st>>>(b);
    if (SV > 1) { k_rp_v2_finish<<<grid_for(K, 32), 32, 0, st>>>(b, SV); ctx->launches++; }
    tm.end();
    ctx->launches += 5;
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}

// Device-resident batch prove.  d_values [K][m], d_blind [K][m][32], d_stream / d_base [K], d_proofs [K][plen] out.
int dapol_rp_prove_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint64_t *d_values, const uint8_t *d_blind, const uint8_t seed[32],
                       const uint64_t *d_stream, const uint64_t *d_base, uint8_t *d_proofs) {
    if (!ctx || !rp_shape_ok(nbits, m) || !K) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = rp_ensure_tables(ctx, m);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const int N = nbits * m, lg = ilog2((uint64_t)N);
    const uint64_t plen = 32ull * (9 + 2 * lg), chunk = rp_chunk(N, m, lg, false, K);
    PhaseTimer tm(st);
    EventPair evp;
    cudaEvent_t e0 = evp.a, e1 = evp.b;
    cudaEventRecord(e0, st);
    int bad = 0;
    for (uint64_t first = 0; first < K; first += chunk) {
        uint64_t kc = std::min(chunk, K - first);
        RpPlan pl;
        rc = rp_plan(ctx, pl, nbits, m, kc, false);
        if (rc) return rc;
        RpBatch &b = pl.b;
        b.values = d_values + first * m;
        b.blind = reinterpret_cast<const uint32_t *>(d_blind) + first * m * 8;
        b.stream = d_stream + first; b.base_block = d_base + first;
        memcpy(b.seed, seed, 32);
        switch (ctx->rp_W) {
#define W_CASE(w) case w: rc = rp_prove_chunk<w>(ctx, b, tm); break;
            RP_W_CASES(W_CASE)
#undef W_CASE
            default: rc = DAPOL_ERR_BAD_ARG;
        }
        if (rc) { dfree(pl.mem, st); return rc; }
        std::vector<int> status(kc);
        cudaError_t ce = cudaMemcpyAsync(d_proofs + first * plen, b.proof, kc * plen, cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(status.data(), b.status, kc * 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { dfree(pl.mem, st); CUDA_TRY(ce); }
        for (int s : status) bad |= s;
        dfree(pl.mem, st);
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ctx->rp_last_ms[0], e0, e1);
    tm.collect();
    tm.publish(ctx->rp_last_ms);
    return bad ? DAPOL_ERR_BAD_ARG : DAPOL_OK;
}
// ------------------------------------------------------------------------------------------------ batched verifier (bucket method)
// kernels over the bodies of rp_kernels.cuh ("batched verification"): weights, terms, buckets, bucket fold, fixed part, verdicts
__global__ void k_rpb_weights(RpBatch b, RpbPlan pl) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rpb_weight_body(pl, p);
}
__global__ void __launch_bounds__(64) k_rpb_terms(RpBatch b, RpbPlan pl) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nv = (uint64_t)rp_nvar(b.lg, b.m);
    if (t < b.K * nv) rpb_terms_body(b, pl, t / nv, (int)(t % nv));
}
__global__ void __launch_bounds__(128) k_rpb_buckets(RpbPlan pl, uint64_t n_buckets, uint64_t n_terms) {
    uint64_t bk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bk < n_buckets) rpb_bucket_body(pl, bk, n_terms);
}
__global__ void __launch_bounds__(128) k_rpb_chunks(RpbPlan pl, uint64_t n_chunks) {
    uint64_t ch = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < n_chunks) rpb_chunk_body(pl, ch);
}
__global__ void __launch_bounds__(64) k_rpb_windows(RpbPlan pl, uint64_t n_gw) {
    uint64_t gw = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw < n_gw) rpb_window_body(pl, gw);
}
__global__ void __launch_bounds__(128) k_rpb_combine(RpBatch b, RpbPlan pl) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = 2ull * b.N + 2, parts = rpb_parts(pl);
    if (t < pl.groups * parts * per) rpb_combine_body(b, pl, t / (parts * per), (t / per) % parts, (uint32_t)(t % per));
}
__global__ void __launch_bounds__(128) k_rpb_combine_sum(RpBatch b, RpbPlan pl) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = 2ull * b.N + 2;
    if (t < pl.groups * per) rpb_combine_sum_body(b, pl, t / per, (uint32_t)(t % per));
}
__global__ void k_rpb_status(RpBatch b, RpbPlan pl) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.K) rpb_status_body(b, pl, p);
}
template <int W, bool INL>
__global__ void __launch_bounds__(RP_MSM_MAX_T, INL ? RP_MSM_MINB_INL : RP_MSM_MINB) k_rpb_fixed(RpBatch b, RpbPlan pl) {
    __shared__ uint32_t sh[RP_REDUCE_SH_WORDS];
    ge acc;
    rpb_fixed_partial<W, INL>(acc, b, pl, blockIdx.x, threadIdx.x, blockDim.x);
    block_reduce_ge(acc, sh);
    if (threadIdx.x == 0) rp_store_ext(pl.gfix + (uint64_t)blockIdx.x * 32, acc);
}
__global__ void __launch_bounds__(32) k_rpb_groups(RpBatch b, RpbPlan pl) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < pl.groups) rpb_group_body(b, pl, g);
}
__global__ void k_rpb_group_ok(uint64_t K, int G, const int *gok, uint8_t *ok) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < K) ok[p] = gok[p / (uint64_t)G] ? 1 : 0;
}
// proofs and commitments of the listed groups, packed back to back (the re-verification input of the failed groups)
__global__ void k_rpb_gather(uint64_t n, const uint32_t *src_p, const uint32_t *src_c, const uint64_t *which, uint32_t pw, uint32_t cw, uint32_t *dst_p, uint32_t *dst_c) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (pw + cw)) return;
    const uint64_t i = t / (pw + cw), k = t % (pw + cw), p = which[i];
    if (k < pw) dst_p[i * pw + k] = src_p[p * pw + k]; else dst_c[i * cw + (k - pw)] = src_c[p * cw + (k - pw)];
}
__global__ void k_rpb_scatter(uint64_t n, const uint64_t *which, const uint8_t *src, uint8_t *ok) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ok[which[i]] = src[i];
}
// window of the bucket method for groups of M points: about 16 .. 32 points per bucket on average keeps a thread per bucket busy
// while the fold of the 2^(c-1) buckets per window stays a small share
static int rpb_window_for(uint64_t M) {
    int lgm = 0;
    while ((2ull << lgm) <= M) lgm++;
    int c = lgm - 3;
    return c < 3 ? 3 : (c > 16 ? 16 : c);
}
int dapol_rp_verify_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint8_t *d_proofs, const uint8_t *d_coms, uint8_t *d_ok);
// one chunk: V0 + expansions as the per-proof verifier, then the groups' combined checks; groups that fail are re-verified
// proof by proof (Straus path) so that d_ok holds the per-proof verdicts
template <int W>
static int rp_verify_chunk_batched(dapol_ctx *ctx, RpBatch &b, PhaseTimer &tm, const uint8_t *d_proofs, const uint8_t *d_coms, uint8_t *d_ok) {
    cudaStream_t st = ctx->stream;
    const uint64_t K = b.K, N = b.N, nv = (uint64_t)rp_nvar(b.lg, b.m);
    RpbPlan pl;
    memset(&pl, 0, sizeof pl);
    pl.G = (int)std::min<uint64_t>((uint64_t)ctx->rp_verify_group, K);
    pl.groups = (K + pl.G - 1) / pl.G;
    rpb_set_windows(pl, ctx->rp_verify_window ? ctx->rp_verify_window : rpb_window_for((uint64_t)pl.G * nv));
    const uint64_t nb = 1ull << (pl.c - 1);
    pl.L = 1;  // chunks of about sqrt(2 nb) buckets: the chunk pass (2 L additions) and the per-window pass (3 nb / L) stay short chains
    while ((uint64_t)pl.L * pl.L < 2 * nb) pl.L <<= 1;
    pl.L = (int)std::min<uint64_t>(pl.L, nb);
    pl.P = 16;
    const uint64_t parts = rpb_parts(pl);
    const uint64_t n_terms = K * nv * pl.NW, n_buckets = pl.groups * pl.NW * nb, n_chunks = n_buckets / pl.L, n_gw = pl.groups * pl.NW;
    if (n_terms >= (1ull << 31) || n_buckets >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    {   // weights: fresh secret randomness per call (the reference's verifier draws its batching weight from thread_rng)
        std::random_device rd;
        for (int i = 0; i < 8; i++) pl.wseed[i] = ctx->rp_verify_seeded ? ctx->rp_verify_seed[i] : (uint32_t)rd();
    }
    size_t sort_bytes = 0;
    int key_bits = 1;
    while ((1ull << key_bits) <= n_buckets) key_bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)n_terms, 0, key_bits, st);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(K, 32) + Arena::need(K * nv, 128) + 4 * Arena::need(n_terms, 4) + Arena::need(n_buckets, 128) + 2 * Arena::need(n_chunks, 128) +
              Arena::need(n_gw, 128) + Arena::need(pl.groups * (2 * N + 2), 32) + Arena::need(pl.groups * parts * (2 * N + 2), 32) +
              Arena::need(pl.groups, 128) + 2 * Arena::need(pl.groups, 4) + Arena::need(sort_bytes, 1);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    pl.rho = ar.take<uint32_t>(K * 8);
    pl.cached = ar.take<uint32_t>(K * nv * 32);
    pl.keys_in = ar.take<uint32_t>(n_terms); pl.keys = ar.take<uint32_t>(n_terms);
    pl.vals_in = ar.take<uint32_t>(n_terms); pl.vals = ar.take<uint32_t>(n_terms);
    pl.bucket = ar.take<uint32_t>(n_buckets * 32);
    pl.chunk_run = ar.take<uint32_t>(n_chunks * 32); pl.chunk_tot = ar.take<uint32_t>(n_chunks * 32);
    pl.window = ar.take<uint32_t>(n_gw * 32);
    pl.gsc = ar.take<uint32_t>(pl.groups * (2 * N + 2) * 8);
    pl.gpart = ar.take<uint32_t>(pl.groups * parts * (2 * N + 2) * 8);
    pl.gbad = ar.take<int>(pl.groups);
    pl.gfix = ar.take<uint32_t>(pl.groups * 32);
    pl.gok = ar.take<int>(pl.groups);
    uint8_t *sort_tmp = ar.take<uint8_t>(sort_bytes);
    tm.begin(1);
    k_rp_v0<<<grid_for(K, 64), 64, 0, st>>>(b);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.svec, 2);
    k_rp_expand<<<(unsigned)K, 256, 0, st>>>(b, b.ypow, 1);
    k_rpb_weights<<<grid_for(K, 128), 128, 0, st>>>(b, pl);
    cudaMemsetAsync(pl.gbad, 0, pl.groups * 4, st);
    tm.end();
    tm.begin(TM_VER);
    k_rpb_terms<<<grid_for(K * nv, 64), 64, 0, st>>>(b, pl);
    cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, pl.keys_in, pl.keys, pl.vals_in, pl.vals, (int)n_terms, 0, key_bits, st);
    k_rpb_buckets<<<grid_for(n_buckets, 128), 128, 0, st>>>(pl, n_buckets, n_terms);
    k_rpb_chunks<<<grid_for(n_chunks, 128), 128, 0, st>>>(pl, n_chunks);
    k_rpb_windows<<<grid_for(n_gw, 64), 64, 0, st>>>(pl, n_gw);
    k_rpb_combine<<<grid_for(pl.groups * parts * (2 * N + 2), 128), 128, 0, st>>>(b, pl);
    k_rpb_combine_sum<<<grid_for(pl.groups * (2 * N + 2), 128), 128, 0, st>>>(b, pl);
    k_rpb_status<<<grid_for(K, 128), 128, 0, st>>>(b, pl);
    if (msm_threads(2 * N) >= RP_INL_MIN_T) k_rpb_fixed<W, true><<<(unsigned)pl.groups, msm_threads(2 * N), 0, st>>>(b, pl);
    else k_rpb_fixed<W, false><<<(unsigned)pl.groups, msm_threads(2 * N), 0, st>>>(b, pl);
    k_rpb_groups<<<grid_for(pl.groups, 32), 32, 0, st>>>(b, pl);
    k_rpb_group_ok<<<grid_for(K, 128), 128, 0, st>>>(K, pl.G, pl.gok, d_ok);
    tm.end();
    ctx->launches += 17;  // 14 kernels + the radix sort's passes (counted as 3)
    std::vector<int> gok(pl.groups);
    cudaError_t ce = cudaMemcpyAsync(gok.data(), pl.gok, pl.groups * 4, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    dfree(mem, st);
    CUDA_TRY(ce);
    // failed groups: per-proof verdicts from the Straus path
    std::vector<uint64_t> redo;
    for (uint64_t g = 0; g < pl.groups; g++)
        if (!gok[g]) for (uint64_t p = g * pl.G; p < std::min<uint64_t>(K, (g + 1) * pl.G); p++) redo.push_back(p);
    ctx->rp_verify_redone += redo.size();
    if (redo.empty()) return DAPOL_OK;
    const uint64_t nr = redo.size();
    const uint32_t pw = b.plen / 4, cw = (uint32_t)b.m * 8;
    Arena a2;
    a2.size = Arena::need(nr, 8) + Arena::need(nr, b.plen) + Arena::need(nr * b.m, 32) + Arena::need(nr, 1);
    CUDA_TRY(dmalloc(&mem, a2.size, st));
    a2.base = mem;
    uint64_t *d_which = a2.take<uint64_t>(nr);
    uint32_t *r_p = a2.take<uint32_t>(nr * pw), *r_c = a2.take<uint32_t>(nr * cw);
    uint8_t *r_ok = a2.take<uint8_t>(nr);
    cudaMemcpyAsync(d_which, redo.data(), nr * 8, cudaMemcpyHostToDevice, st);
    k_rpb_gather<<<grid_for(nr * (pw + cw), 256), 256, 0, st>>>(nr, reinterpret_cast<const uint32_t *>(d_proofs), reinterpret_cast<const uint32_t *>(d_coms), d_which, pw, cw, r_p, r_c);
    const int saved = ctx->rp_verify_group;
    ctx->rp_verify_group = 0;
    float keep[8];
    memcpy(keep, ctx->rp_last_ms, sizeof keep);
    int rc = dapol_rp_verify_dev(ctx, b.nbits, b.m, nr, reinterpret_cast<const uint8_t *>(r_p), reinterpret_cast<const uint8_t *>(r_c), r_ok);
    ctx->rp_verify_group = saved;
    memcpy(ctx->rp_last_ms, keep, sizeof keep);
    if (rc == DAPOL_OK) {
        k_rpb_scatter<<<grid_for(nr, 128), 128, 0, st>>>(nr, d_which, r_ok, d_ok);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = DAPOL_ERR_CUDA;
    }
    ctx->launches += 2;
    dfree(mem, st);
    return rc;
}

// Device-resident batch verify.  d_proofs [K][plen], d_coms [K][m][32], d_ok [K] out (1 accept, 0 reject).
int dapol_rp_verify_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint8_t *d_proofs, const uint8_t *d_coms, uint8_t *d_ok);
__global__ void k_status_to_ok(uint64_t K, const int *status, uint8_t *ok) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < K) ok[p] = status[p] ? 1 : 0;
}
int dapol_rp_verify_dev(dapol_ctx *ctx, int nbits, int m, uint64_t K, const uint8_t *d_proofs, const uint8_t *d_coms, uint8_t *d_ok) {
    if (!ctx || !rp_shape_ok(nbits, m) || !K) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = rp_ensure_tables(ctx, m);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const int N = nbits * m, lg = ilog2((uint64_t)N);
    const uint64_t plen = 32ull * (9 + 2 * lg), chunk = rp_chunk(N, m, lg, true, K);
    PhaseTimer tm(st);
    EventPair evp;
    cudaEvent_t e0 = evp.a, e1 = evp.b;
    cudaEventRecord(e0, st);
    for (uint64_t first = 0; first < K; first += chunk) {
        uint64_t kc = std::min(chunk, K - first);
        RpPlan pl;
        rc = rp_plan(ctx, pl, nbits, m, kc, true);
        if (rc) return rc;
        RpBatch &b = pl.b;
        b.proof_in = reinterpret_cast<const uint32_t *>(d_proofs + first * plen);
        b.coms = reinterpret_cast<const uint32_t *>(d_coms) + first * m * 8;
        const bool batched = ctx->rp_verify_group > 1 && kc > 1;
        switch (ctx->rp_W) {
#define W_CASE(w) case w: rc = batched ? rp_verify_chunk_batched<w>(ctx, b, tm, d_proofs + first * plen, d_coms + first * m * 32, d_ok + first) \
                                       : rp_verify_chunk<w>(ctx, b, tm); break;
            RP_W_CASES(W_CASE)
#undef W_CASE
            default: rc = DAPOL_ERR_BAD_ARG;
        }
        if (rc) { dfree(pl.mem, st); return rc; }
        if (!batched) {
            k_status_to_ok<<<grid_for(kc, 128), 128, 0, st>>>(kc, b.status, d_ok + first);
            ctx->launches++;
        }
        dfree(pl.mem, st);
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ctx->rp_last_ms[0], e0, e1);
    tm.collect();
    tm.publish(ctx->rp_last_ms);
    CUDA_TRY(cudaGetLastError());
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int dapol_rangeproof_prove_batch_dev(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint64_t *d_values, const uint8_t *d_blindings,
                                                const uint8_t seed[32], const uint64_t *d_streams, const uint64_t *d_base_blocks,
                                                uint8_t *d_proofs) {
    if (!d_values || !d_blindings || !seed || !d_streams || !d_base_blocks || !d_proofs) return DAPOL_ERR_BAD_ARG;
    return dapol_rp_prove_dev(ctx, nbits, m, k, d_values, d_blindings, seed, d_streams, d_base_blocks, d_proofs);
}
extern "C" int dapol_rangeproof_verify_batch_dev(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint8_t *d_proofs, uint64_t proof_len,
                                                 const uint8_t *d_commitments, uint8_t *d_ok) {
    if (!ctx || !d_proofs || !d_commitments || !d_ok || !rp_shape_ok(nbits, m)) return DAPOL_ERR_BAD_ARG;
    if (proof_len != dapol_rangeproof_size(nbits, m)) {  // RangeProof::from_bytes / verify: wrong size for (n, m) is a reject, not an error
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaMemsetAsync(d_ok, 0, k, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return DAPOL_OK;
    }
    return dapol_rp_verify_dev(ctx, nbits, m, k, d_proofs, d_commitments, d_ok);
}
extern "C" int dapol_rangeproof_prove_batch(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint64_t *values, const uint8_t *blindings,
                                            const uint8_t seed[32], const uint64_t *streams, const uint64_t *base_blocks, uint8_t *proofs) {
    if (!ctx || !values || !blindings || !seed || !streams || !base_blocks || !proofs || !rp_shape_ok(nbits, m) || !k) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t plen = dapol_rangeproof_size(nbits, m);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(k * m, 8) + Arena::need(k * m, 32) + 2 * Arena::need(k, 8) + Arena::need(k, plen);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_v = ar.take<uint64_t>(k * m);
    uint8_t *d_b = ar.take<uint8_t>(k * m * 32);
    uint64_t *d_s = ar.take<uint64_t>(k), *d_bb = ar.take<uint64_t>(k);
    uint8_t *d_p = ar.take<uint8_t>(k * plen);
    cudaMemcpyAsync(d_v, values, k * m * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_b, blindings, k * m * 32, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_s, streams, k * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_bb, base_blocks, k * 8, cudaMemcpyHostToDevice, st);
    int rc = dapol_rp_prove_dev(ctx, nbits, m, k, d_v, d_b, seed, d_s, d_bb, d_p);
    if (rc == DAPOL_OK) {
        cudaMemcpyAsync(proofs, d_p, k * plen, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = DAPOL_ERR_CUDA;
    }
    dfree(mem, st);
    return rc;
}
extern "C" int dapol_rangeproof_verify_batch(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint8_t *proofs, uint64_t proof_len,
                                             const uint8_t *commitments, uint8_t *ok) {
    if (!ctx || !proofs || !commitments || !ok || !rp_shape_ok(nbits, m) || !k) return DAPOL_ERR_BAD_ARG;
    if (proof_len != dapol_rangeproof_size(nbits, m)) { memset(ok, 0, k); return DAPOL_OK; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(k, proof_len) + Arena::need(k * m, 32) + Arena::need(k, 1);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint8_t *d_p = ar.take<uint8_t>(k * proof_len), *d_c = ar.take<uint8_t>(k * m * 32), *d_ok = ar.take<uint8_t>(k);
    cudaMemcpyAsync(d_p, proofs, k * proof_len, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_c, commitments, k * m * 32, cudaMemcpyHostToDevice, st);
    int rc = dapol_rp_verify_dev(ctx, nbits, m, k, d_p, d_c, d_ok);
    if (rc == DAPOL_OK) {
        cudaMemcpyAsync(ok, d_ok, k, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = DAPOL_ERR_CUDA;
    }
    dfree(mem, st);
    return rc;
}
extern "C" int dapol_rangeproof_last_times(const dapol_ctx *ctx, float ms[4]) {
    if (!ctx || !ms) return DAPOL_ERR_BAD_ARG;
    memcpy(ms, ctx->rp_last_ms, sizeof(float) * 4);
    return DAPOL_OK;
}
extern "C" int dapol_rangeproof_last_kernel_times(const dapol_ctx *ctx, float ms[8]) {
    if (!ctx || !ms) return DAPOL_ERR_BAD_ARG;
    memcpy(ms, ctx->rp_last_ms, sizeof(float) * 8);
    return DAPOL_OK;
}
extern "C" int dapol_ctx_set_verify_mode(dapol_ctx *ctx, uint64_t group, int window_bits, const uint8_t *weight_seed) {
    if (!ctx || group > (1u << 20) || window_bits < 0 || (window_bits && (window_bits < 3 || window_bits > 16))) return DAPOL_ERR_BAD_ARG;
    ctx->rp_verify_group = (int)group;
    ctx->rp_verify_window = window_bits;
    ctx->rp_verify_seeded = weight_seed != nullptr;
    if (weight_seed) memcpy(ctx->rp_verify_seed, weight_seed, 32);
    return DAPOL_OK;
}
extern "C" uint64_t dapol_ctx_verify_fallbacks(const dapol_ctx *ctx) { return ctx ? ctx->rp_verify_redone : 0; }
extern "C" int dapol_ctx_set_rangeproof_table_budget(dapol_ctx *ctx, uint64_t bytes) {
    if (!ctx) return DAPOL_ERR_BAD_ARG;
    ctx->rp_budget = bytes;
    return DAPOL_OK;
}
extern "C" uint64_t dapol_ctx_rangeproof_table_bytes(const dapol_ctx *ctx) { return ctx && ctx->rp_tab ? ctx->rp_table_bytes : 0; }
extern "C" int dapol_ctx_set_rangeproof_window(dapol_ctx *ctx, int window) {
    if (!ctx || (window != 0 && window != 8 && window != 12 && window != 13 && window != 14 && window != 15 && window != 16)) return DAPOL_ERR_BAD_ARG;
    ctx->rp_W_auto = window == 0;
    if (window == 0) {  // the width is chosen when the tables are (re)built
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->rp_tab) cudaFree(ctx->rp_tab);
        ctx->rp_tab = nullptr; ctx->rp_mcap = 0;
        return DAPOL_OK;
    }
    if (ctx->rp_W != window) {
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->rp_tab) cudaFree(ctx->rp_tab);
        ctx->rp_tab = nullptr; ctx->rp_mcap = 0; ctx->rp_W = window;
    }
    return DAPOL_OK;
}
