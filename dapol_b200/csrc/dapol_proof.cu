// Inclusion proofs of the DAPOL+ hot path: Dapol::generate_proof_batch (src/dapol/mod.rs:172-190) with the two
// aggregation policies (src/range/padding.rs:88-118, src/range/splitting.rs:100-129), DapolProof::serialize /
// deserialize (src/proof/mod.rs:68-84) and DapolProof::verify (src/proof/mod.rs:41-47,89-95: Merkle fold with
// DapolProofNode::merge, src/proof/node.rs:56-69, then R::verify, padding.rs:168-197 / splitting.rs:180-211).
// The range proofs of a batch of leaves run as a few LARGE batches on the GPU (one per aggregate size + one of singles).
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>
#include "dapol_internal.h"
#include "rp_kernels.cuh"

#define SINGLE_PROOF_BYTE_NUM 672  // src/range/mod.rs:18

static inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static inline void put_be(uint8_t *o, uint64_t x, int k) { for (int i = 0; i < k; i++) o[i] = (uint8_t)(x >> (8 * (k - 1 - i))); }
static inline uint64_t get_be(const uint8_t *o, int k) { uint64_t x = 0; for (int i = 0; i < k; i++) x = (x << 8) | o[i]; return x; }

// which siblings go into which aggregated proof; the rest are single proofs (padding.rs:88-118, splitting.rs:100-129)
struct AggGroup {
    uint64_t start, count, m;
};
static int policy_plan(uint64_t nsib, uint64_t agg, int policy, std::vector<AggGroup> &g, uint64_t &single_from) {
    g.clear();
    if (agg > nsib) return DAPOL_ERR_BAD_ARG;  // reference: slice out of bounds panic (padding.rs:95-97, splitting.rs:110-113)
    if (policy == DAPOL_POLICY_PADDING) {
        g.push_back({0, agg, next_pow2(agg)});  // agg = 0: next_power_of_two(0) = 1 -> one proof of a single dummy party
        single_from = agg;
    } else if (policy == DAPOL_POLICY_SPLITTING) {
        uint64_t base = next_pow2(agg), pos = 0;
        while (pos < agg) {
            if (agg & base) { g.push_back({pos, base, base}); pos += base; }
            base >>= 1;
        }
        single_from = pos;
    } else return DAPOL_ERR_BAD_ARG;
    for (auto &x : g) if (x.m > 64) return DAPOL_ERR_BAD_ARG;
    return DAPOL_OK;
}
// a sibling on the wire is DapolProofNode::serialize = com (32) || hash (Dlen), src/proof/node.rs:74-79
static uint64_t merkle_bytes(uint64_t H, uint64_t dl) { return 2 + 8 + (H + 7) / 8 + 8 + (32 + dl) * H; }
extern "C" uint64_t dapol_inclusion_proof_size_d(int height, uint64_t aggregation_factor, int policy, int hash_id) {
    std::vector<AggGroup> g;
    uint64_t sf;
    if (!dapol_digest_len(hash_id) || height < 0 || height > 64 || policy_plan((uint64_t)height, aggregation_factor, policy, g, sf)) return 0;
    uint64_t sz = policy == DAPOL_POLICY_SPLITTING ? 2 : 0;
    for (auto &x : g) sz += 8 + dapol_rangeproof_size(64, (int)x.m);
    sz += 8 + SINGLE_PROOF_BYTE_NUM * ((uint64_t)height - sf);
    return sz + merkle_bytes((uint64_t)height, (uint64_t)dapol_digest_len(hash_id));
}
extern "C" uint64_t dapol_inclusion_proof_size(int height, uint64_t aggregation_factor, int policy) {
    return dapol_inclusion_proof_size_d(height, aggregation_factor, policy, DAPOL_HASH_BLAKE3);
}

// ------------------------------------------------------------------------------------------------ prove
// values / blindings of one aggregated group for every leaf: party j < count is sibling start + j, the rest are the
// (0, Scalar::one()) dummy parties of the Padding policy (padding.rs:98-101)
__global__ void k_gather_group(uint64_t k, int H, uint64_t start, uint64_t count, uint64_t m, const uint64_t *pv, const uint32_t *pr,
                               uint64_t *ov, uint32_t *orr) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * m) return;
    uint64_t p = t / m, j = t % m;
    uint32_t w[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    uint64_t v = 0;
    if (j < count) { v = pv[p * H + start + j]; load8(w, pr + (p * H + start + j) * 8); }
    ov[t] = v;
    store8(orr + t * 8, w);
}
// ChaCha20 key of the prover's nonce streams of one tree: BLAKE3(label || seed || root commitment || root hash || le64 policy ||
// le64 aggregation factor || le64 height).  The root commits to every witness of every proof, so the caller's seed re-used for
// another tree (next audit), another policy or another aggregation factor never meets a nonce twice with different witnesses
// (two e_blinding = alpha + rho x for one (alpha, rho) would leak both), and it is domain-separated from the padding stream.
static void prover_nonce_key(uint8_t key[32], const uint8_t seed[32], const dapol_tree *whole, int policy, uint64_t agg, uint64_t height) {
    static const uint8_t label[30] = {'d', 'a', 'p', 'o', 'l', '-', 'b', '2', '0', '0', ' ', 'p', 'r', 'o', 'v', 'e', 'r', ' ', 'n', 'o', 'n', 'c', 'e', ' ',
                                      'k', 'e', 'y', ' ', 'v', '1'};
    const uint64_t w[3] = {(uint64_t)policy, agg, height};
    uint8_t tail[24];
    for (int i = 0; i < 3; i++) for (int b = 0; b < 8; b++) tail[8 * i + b] = (uint8_t)(w[i] >> (8 * b));
    dapol_hasher hs;
    uint32_t out[8];
    hasher_init(hs, DAPOL_HASH_BLAKE3);
    hasher_update(hs, label, 30);
    hasher_update(hs, seed, 32);
    hasher_update_words(hs, whole->root_comc, 8);
    hasher_update_words(hs, whole->root_hash, dapol_dlen(whole->hash_id) / 4);
    hasher_update(hs, tail, 24);
    hasher_final(hs, out);
    memcpy(key, out, 32);
}
// RNG contract: range proof #q of the inclusion proof of leaf x draws from ChaCha20(prover_nonce_key) stream x from block q << 32
__global__ void k_proof_streams(uint64_t k, uint64_t per, uint64_t q0, const uint64_t *leaf_idx, uint64_t *stream, uint64_t *base) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * per) return;
    stream[t] = leaf_idx[t / per];
    base[t] = (q0 + t % per) << 32;
}

extern "C" int dapol_prove_batch(const dapol_tree *t, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                 const uint8_t seed[32], uint8_t *out, uint64_t cap, uint64_t *proof_size) {
    if (!t || !leaf_idx || !seed || !k) return DAPOL_ERR_BAD_ARG;
    dapol_ctx *ctx = t->ctx;
    const uint64_t H = (uint64_t)dapol_total_height(t);  // a shard with its top tree attached proves against the whole tree
    std::vector<AggGroup> groups;
    uint64_t sf;
    int rc = policy_plan(H, aggregation_factor, policy, groups, sf);
    if (rc) return rc;
    const uint64_t size = dapol_inclusion_proof_size_d((int)H, aggregation_factor, policy, t->hash_id);
    const bool b2b = t->ns.hash_hi != nullptr;  // 64-byte digests: a sibling is com || hash lo || hash hi
    if (proof_size) *proof_size = size;
    if (!out || cap < k * size) return DAPOL_ERR_BUFFER;
    uint8_t key[32];
    prover_nonce_key(key, seed, t->top ? t->top : t, policy, aggregation_factor, H);
    seed = key;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t kh = k * (H ? H : 1), nsingle = H - sf;
    uint64_t max_m = 1;
    for (auto &g : groups) max_m = std::max(max_m, g.m);
    const uint64_t rows = std::max(k * max_m, k * std::max<uint64_t>(nsingle, 1));  // proofs x parties of the largest range-proof batch
    uint64_t agg_bytes = 0;
    for (auto &g : groups) agg_bytes += dapol_rangeproof_size(64, (int)g.m);
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(k, 8) + Arena::need(kh, 8) + 4 * Arena::need(kh, 32) + 2 * Arena::need(k, 32) + 256 + Arena::need(rows, 8) +
              Arena::need(rows, 32) + 2 * Arena::need(rows, 8) + Arena::need(k * agg_bytes, 1) + Arena::need(k * nsingle + 1, SINGLE_PROOF_BYTE_NUM);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint64_t *d_li = ar.take<uint64_t>(k), *d_v = ar.take<uint64_t>(kh);
    uint32_t *d_r = ar.take<uint32_t>(kh * 8), *d_c = ar.take<uint32_t>(kh * 8), *d_h = ar.take<uint32_t>(kh * 8), *d_hh = ar.take<uint32_t>(kh * 8);
    uint32_t *d_lc = ar.take<uint32_t>(k * 8), *d_lh = ar.take<uint32_t>(k * 8);
    int *d_nf = ar.take<int>(1), nf = 0;
    uint64_t *g_v = ar.take<uint64_t>(rows);
    uint32_t *g_r = ar.take<uint32_t>(rows * 8);
    uint64_t *g_s = ar.take<uint64_t>(rows), *g_b = ar.take<uint64_t>(rows);
    uint8_t *d_agg = ar.take<uint8_t>(k * agg_bytes), *d_single = ar.take<uint8_t>((k * nsingle + 1) * SINGLE_PROOF_BYTE_NUM);
#define FAIL(code) do { dfree(mem, st); return (code); } while (0)
    cudaMemsetAsync(d_nf, 0, 4, st);
    cudaMemcpyAsync(d_li, leaf_idx, k * 8, cudaMemcpyHostToDevice, st);
    rc = dapol_tree_paths_dev(t, k, d_li, d_v, d_r, d_c, d_h, d_lc, d_lh, d_nf, b2b ? d_hh : nullptr, nullptr);
    if (rc) FAIL(rc);
    cudaMemcpyAsync(&nf, d_nf, 4, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) FAIL(DAPOL_ERR_CUDA);
    if (nf) FAIL(DAPOL_ERR_NOT_FOUND);  // reference: None for an index that is not a leaf (mod.rs:173)
    // aggregated proofs, one GPU batch of k proofs per group
    uint64_t q = 0, off = 0;
    std::vector<uint64_t> agg_off;
    for (auto &g : groups) {
        k_gather_group<<<grid_for(k * g.m, 128), 128, 0, st>>>(k, (int)H, g.start, g.count, g.m, d_v, d_r, g_v, g_r);
        k_proof_streams<<<grid_for(k, 128), 128, 0, st>>>(k, 1, q, d_li, g_s, g_b);
        ctx->launches += 2;
        rc = dapol_rp_prove_dev(ctx, 64, (int)g.m, k, g_v, reinterpret_cast<const uint8_t *>(g_r), seed, g_s, g_b, d_agg + k * off);
        if (rc) FAIL(rc);
        agg_off.push_back(off);
        off += dapol_rangeproof_size(64, (int)g.m);
        q++;
    }
    // single proofs of the remaining siblings: one batch of k * nsingle proofs (sibling-major per leaf)
    if (nsingle) {
        k_gather_group<<<grid_for(k * nsingle, 128), 128, 0, st>>>(k, (int)H, sf, nsingle, nsingle, d_v, d_r, g_v, g_r);
        k_proof_streams<<<grid_for(k * nsingle, 128), 128, 0, st>>>(k, nsingle, q, d_li, g_s, g_b);
        ctx->launches += 2;
        rc = dapol_rp_prove_dev(ctx, 64, 1, k * nsingle, g_v, reinterpret_cast<const uint8_t *>(g_r), seed, g_s, g_b, d_single);
        if (rc) FAIL(rc);
    }
    // host assembly of DapolProof::serialize = R::serialize || MerkleProof::serialize
    std::vector<uint8_t> h_agg(k * agg_bytes), h_single(k * nsingle * SINGLE_PROOF_BYTE_NUM), h_c(kh * 32), h_h(kh * 32), h_hh(b2b ? kh * 32 : 0);
    if (agg_bytes) cudaMemcpyAsync(h_agg.data(), d_agg, h_agg.size(), cudaMemcpyDeviceToHost, st);
    if (nsingle) cudaMemcpyAsync(h_single.data(), d_single, h_single.size(), cudaMemcpyDeviceToHost, st);
    if (H) { cudaMemcpyAsync(h_c.data(), d_c, kh * 32, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(h_h.data(), d_h, kh * 32, cudaMemcpyDeviceToHost, st); }
    if (H && b2b) cudaMemcpyAsync(h_hh.data(), d_hh, kh * 32, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) FAIL(DAPOL_ERR_CUDA);
    dfree(mem, st);
#undef FAIL
    for (uint64_t p = 0; p < k; p++) {
        uint8_t *o = out + p * size;
        if (policy == DAPOL_POLICY_SPLITTING) { put_be(o, groups.size(), 2); o += 2; }         // splitting.rs:38-59
        for (size_t gi = 0; gi < groups.size(); gi++) {                                        // padding.rs:40-53
            uint64_t plen = dapol_rangeproof_size(64, (int)groups[gi].m);
            put_be(o, plen, 8); o += 8;
            memcpy(o, h_agg.data() + k * agg_off[gi] + p * plen, plen); o += plen;
        }
        put_be(o, nsingle, 8); o += 8;
        memcpy(o, h_single.data() + p * nsingle * SINGLE_PROOF_BYTE_NUM, nsingle * SINGLE_PROOF_BYTE_NUM);
        o += nsingle * SINGLE_PROOF_BYTE_NUM;
        // MerkleProof::serialize (smtree ^0.1.2; prefix widths are UPSTREAM-RECALL): height, #indexes,
        // path bits MSB-first, #siblings, siblings (com || hash, src/proof/node.rs:74-79) leaf level first
        put_be(o, H, 2); o += 2;
        put_be(o, 1, 8); o += 8;
        uint64_t nb = (H + 7) / 8;
        if (nb) { put_be(o, H == 64 ? leaf_idx[p] : leaf_idx[p] << (8 * nb - H), (int)nb); o += nb; }
        put_be(o, H, 8); o += 8;
        for (uint64_t s = 0; s < H; s++) {
            memcpy(o, h_c.data() + (p * H + s) * 32, 32);
            memcpy(o + 32, h_h.data() + (p * H + s) * 32, 32);
            o += 64;
            if (b2b) { memcpy(o, h_hh.data() + (p * H + s) * 32, 32); o += 32; }
        }
    }
    return DAPOL_OK;
}

// ------------------------------------------------------------------------------------------------ verify
// MerkleProof::verify: fold upward from the leaf with DapolProofNode::merge (src/proof/node.rs:56-69):
// hash = D(C_l || C_r || H_l || H_r), com = com_l + com_r; sibs = [H][com 32 || hash Dlen].
// The expensive steps run side by side (a thread per proof would run 2 H inverse square roots one after the other -- 3.8 ms for
// ONE height-16 proof -- and leave the GPU idle on small batches): every point of a proof is decompressed by its
// own thread, the running commitment of level lv = leaf + sib_0 + .. + sib_(lv-1) is summed and compressed by its own thread, and
// only the hash chain (H short hashes) stays sequential.  pts: [k][hmax + 1][32] (0 = leaf, j + 1 = sibling j); comc: [k][hmax + 1][8].
__global__ void __launch_bounds__(64) k_merkle_decompress(uint64_t k, uint32_t hmax, uint32_t ss, const uint8_t *blob, const uint64_t *sib_off, const uint32_t *heights,
                                                          const uint32_t *leaf_c, uint32_t *pts, uint8_t *ok) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * (hmax + 1)) return;
    const uint64_t p = t / (hmax + 1);
    const uint32_t j = (uint32_t)(t % (hmax + 1));
    if (!ok[p] || j > heights[p]) return;
    uint32_t w[8];
    if (j == 0) load8(w, leaf_c + 8 * p);
    else {
        const uint8_t *a = blob + sib_off[p] + (uint64_t)ss * (j - 1);  // the blob is byte-aligned only
        for (int i = 0; i < 8; i++) w[i] = (uint32_t)a[4 * i] | ((uint32_t)a[4 * i + 1] << 8) | ((uint32_t)a[4 * i + 2] << 16) | ((uint32_t)a[4 * i + 3] << 24);
    }
    ge q;
    if (!ge_decompress(q, w)) { ok[p] = 0; return; }
    rp_store_ext(pts + (p * (hmax + 1) + j) * 32, q);
}
__global__ void __launch_bounds__(64) k_merkle_prefix(uint64_t k, uint32_t hmax, const uint32_t *heights, const uint32_t *pts, uint32_t *comc, const uint8_t *ok) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * (hmax + 1)) return;
    const uint64_t p = t / (hmax + 1);
    const uint32_t lv = (uint32_t)(t % (hmax + 1));
    if (!ok[p] || lv == 0 || lv > heights[p]) return;
    ge acc, q;
    rp_load_ext(acc, pts + (p * (hmax + 1)) * 32);
#pragma unroll 1
    for (uint32_t j = 1; j <= lv; j++) { rp_load_ext(q, pts + (p * (hmax + 1) + j) * 32); ge_add(acc, acc, q); }
    uint32_t c[8];
    ge_compress(c, acc);
    store8(comc + (p * (hmax + 1) + lv) * 8, c);
}
__global__ void __launch_bounds__(64) k_merkle_hash_chain(uint64_t k, uint32_t hmax, int hash_id, uint32_t ss, const uint8_t *blob, const uint64_t *sib_off,
                                                          const uint32_t *heights, const uint64_t *idx, const uint32_t *leaf_c, const uint32_t *leaf_h,
                                                          const uint32_t *comc, const uint32_t *root, uint8_t *ok) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= k || !ok[p]) return;
    const bool wide = hash_id == DAPOL_HASH_BLAKE2B;  // 64-byte digests: leaf_h = [k][16], root = com (8) | hash (16)
    const int hw = wide ? 16 : 8;
    uint32_t curc[8], curl[8], curu[8], sc_[8], sl[8], su[8];
    load8(curc, leaf_c + 8 * p); load8(curl, leaf_h + (uint64_t)hw * p);
    if (wide) load8(curu, leaf_h + 16 * p + 8);
    const uint8_t *sib = blob + sib_off[p];
    const uint32_t H = heights[p];
    const uint64_t x = idx[p];
#pragma unroll 1
    for (uint32_t lv = 0; lv < H; lv++) {
        const uint8_t *a = sib + (uint64_t)ss * lv;
        for (int i = 0; i < 8; i++) {
            sc_[i] = (uint32_t)a[4 * i] | ((uint32_t)a[4 * i + 1] << 8) | ((uint32_t)a[4 * i + 2] << 16) | ((uint32_t)a[4 * i + 3] << 24);
            sl[i] = (uint32_t)a[32 + 4 * i] | ((uint32_t)a[33 + 4 * i] << 8) | ((uint32_t)a[34 + 4 * i] << 16) | ((uint32_t)a[35 + 4 * i] << 24);
            if (wide) su[i] = (uint32_t)a[64 + 4 * i] | ((uint32_t)a[65 + 4 * i] << 8) | ((uint32_t)a[66 + 4 * i] << 16) | ((uint32_t)a[67 + 4 * i] << 24);
        }
        uint32_t lo[8], hi[8];
        const bool right = (x >> lv) & 1;  // the running node is the right child
        if (wide) {
            if (right) dapol_b2b_hash192(lo, hi, sc_, curc, sl, su, curl, curu); else dapol_b2b_hash192(lo, hi, curc, sc_, curl, curu, sl, su);
        } else {
            if (right) dapol_hash128(hash_id, lo, sc_, curc, sl, curl); else dapol_hash128(hash_id, lo, curc, sc_, curl, sl);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { curl[i] = lo[i]; if (wide) curu[i] = hi[i]; }
        load8(curc, comc + (p * (hmax + 1) + lv + 1) * 8);
    }
    uint32_t d = 0;
    for (int i = 0; i < 8; i++) d |= (curc[i] ^ root[i]) | (curl[i] ^ root[8 + i]) | (wide ? (curu[i] ^ root[16 + i]) : 0u);
    ok[p] = (uint8_t)(d == 0);
}

struct ParsedProof {
    bool ok = false;
    std::vector<std::pair<uint64_t, uint64_t>> agg;  // (offset, length) inside the blob
    uint64_t nind = 0, ind_off = 0, H = 0, idx = 0, sib_off = 0;
};
// DapolProof::deserialize (src/proof/mod.rs:76-84): R::deserialize then MerkleProof::deserialize; any framing error rejects
static ParsedProof parse_proof(const uint8_t *p, uint64_t base, uint64_t len, int policy, uint64_t ss /* bytes of one sibling: 32 + Dlen */) {
    ParsedProof r;
    uint64_t pos = 0, nagg = 1;
#define NEED(n) do { if (len - pos < (uint64_t)(n)) return r; } while (0)
    if (policy == DAPOL_POLICY_SPLITTING) { NEED(2); nagg = get_be(p, 2); pos += 2; if (nagg > 64) return r; }
    for (uint64_t i = 0; i < nagg; i++) {
        NEED(8);
        uint64_t sz = get_be(p + pos, 8); pos += 8;
        if (sz > len - pos) return r;
        r.agg.push_back({base + pos, sz}); pos += sz;
    }
    NEED(8);
    r.nind = get_be(p + pos, 8); pos += 8;
    if (r.nind > 64 || (len - pos) / SINGLE_PROOF_BYTE_NUM < r.nind) return r;
    r.ind_off = base + pos; pos += SINGLE_PROOF_BYTE_NUM * r.nind;
    NEED(10);
    r.H = get_be(p + pos, 2); pos += 2;
    if (get_be(p + pos, 8) != 1 || r.H > 64) return r;
    pos += 8;
    uint64_t nb = (r.H + 7) / 8;
    NEED(nb + 8);
    r.idx = nb ? get_be(p + pos, (int)nb) : 0;
    if (nb && r.H != 64) r.idx >>= (8 * nb - r.H);
    pos += nb;
    uint64_t nsib = get_be(p + pos, 8); pos += 8;
    if (nsib != r.H || (len - pos) / ss < nsib) return r;
    r.sib_off = base + pos;
    if (r.nind > nsib) return r;  // reference: usize underflow panic (padding.rs:171, splitting.rs:182)
#undef NEED
    r.ok = true;
    return r;
}

extern "C" int dapol_verify_batch(dapol_ctx *ctx, int hash_id, int policy, uint64_t k, const uint8_t root_com[32], const uint8_t root_hash[32],
                                  const uint8_t *leaf_coms, const uint8_t *leaf_hashes, const uint8_t *proofs, const uint64_t *offsets,
                                  uint8_t *ok) {
    if (!ctx || !root_com || !root_hash || !leaf_coms || !leaf_hashes || !proofs || !offsets || !ok || !k) return DAPOL_ERR_BAD_ARG;
    const uint64_t dl = (uint64_t)dapol_digest_len(hash_id), ss = 32 + dl;  // root_hash / leaf_hashes: dl bytes each
    if (!dl) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (policy != DAPOL_POLICY_PADDING && policy != DAPOL_POLICY_SPLITTING) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t total = offsets[k];
    std::vector<ParsedProof> pp(k);
    std::vector<uint64_t> h_sib(k), h_idx(k);
    std::vector<uint32_t> h_H(k);
    for (uint64_t p = 0; p < k; p++) {
        if (offsets[p + 1] < offsets[p] || offsets[p + 1] > total) return DAPOL_ERR_BAD_ARG;
        pp[p] = parse_proof(proofs + offsets[p], offsets[p], offsets[p + 1] - offsets[p], policy, ss);
        ok[p] = pp[p].ok ? 1 : 0;
        h_sib[p] = pp[p].sib_off; h_idx[p] = pp[p].idx; h_H[p] = (uint32_t)pp[p].H;
    }
    // range-proof work items grouped by aggregate size m: (proof, blob offset of the range proof, first sibling, real parties)
    struct Item { uint64_t p, off, first, count; };
    std::map<uint64_t, std::vector<Item>> by_m;
    for (uint64_t p = 0; p < k; p++) {
        if (!pp[p].ok) continue;
        const ParsedProof &r = pp[p];
        uint64_t nagg_coms = r.H - r.nind;
        std::vector<AggGroup> groups;
        uint64_t sf;
        if (policy_plan(r.H, nagg_coms, policy, groups, sf) || groups.size() > r.agg.size() ||
            (policy == DAPOL_POLICY_PADDING && r.agg.size() != 1)) { ok[p] = 0; continue; }
        bool good = true;
        for (size_t gi = 0; gi < groups.size() && good; gi++)
            if (r.agg[gi].second != dapol_rangeproof_size(64, (int)groups[gi].m)) good = false;  // RangeProof::from_bytes / size mismatch
        if (!good) { ok[p] = 0; continue; }
        for (size_t gi = 0; gi < groups.size(); gi++) by_m[groups[gi].m].push_back({p, r.agg[gi].first, groups[gi].start, groups[gi].count});
        for (uint64_t s = 0; s < r.nind; s++) by_m[1].push_back({p, r.ind_off + s * SINGLE_PROOF_BYTE_NUM, nagg_coms + s, 1});
    }
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(total + 1, 1) + 2 * Arena::need(k, 8) + Arena::need(k, 4) + 3 * Arena::need(k, 32) + 256 + Arena::need(k, 1);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint8_t *d_blob = ar.take<uint8_t>(total + 1);
    uint64_t *d_sib = ar.take<uint64_t>(k), *d_idx = ar.take<uint64_t>(k);
    uint32_t *d_H = ar.take<uint32_t>(k), *d_lc = ar.take<uint32_t>(k * 8), *d_lh = ar.take<uint32_t>(k * 16), *d_root = ar.take<uint32_t>(24);
    uint8_t *d_ok = ar.take<uint8_t>(k);
    cudaMemcpyAsync(d_blob, proofs, total, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_sib, h_sib.data(), k * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_idx, h_idx.data(), k * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_H, h_H.data(), k * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_lc, leaf_coms, k * 32, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_lh, leaf_hashes, k * dl, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_root, root_com, 32, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_root + 8, root_hash, dl, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_ok, ok, k, cudaMemcpyHostToDevice, st);
    uint32_t hmax = 0;
    for (uint64_t p = 0; p < k; p++) if (pp[p].ok) hmax = std::max(hmax, h_H[p]);
    {   // Merkle fold in slices of at most 2^20 (proof, level) items = 160 MB of scratch
        const uint64_t per = hmax + 1, slice = std::max<uint64_t>(1, (1ull << 20) / per), items = std::min(k, slice) * per;
        const int hw = hash_id == DAPOL_HASH_BLAKE2B ? 16 : 8;
        uint8_t *fold_mem = nullptr;
        if (dmalloc(&fold_mem, Arena::need(items, 128) + Arena::need(items, 32), st) != cudaSuccess) { dfree(mem, st); CUDA_TRY(cudaGetLastError()); return DAPOL_ERR_CUDA; }
        uint32_t *pts = reinterpret_cast<uint32_t *>(fold_mem), *comc = reinterpret_cast<uint32_t *>(fold_mem + Arena::need(items, 128));
        for (uint64_t p0 = 0; p0 < k; p0 += slice) {
            const uint64_t kc = std::min(slice, k - p0), it = kc * per;
            k_merkle_decompress<<<grid_for(it, 64), 64, 0, st>>>(kc, hmax, (uint32_t)ss, d_blob, d_sib + p0, d_H + p0, d_lc + 8 * p0, pts, d_ok + p0);
            k_merkle_prefix<<<grid_for(it, 64), 64, 0, st>>>(kc, hmax, d_H + p0, pts, comc, d_ok + p0);
            k_merkle_hash_chain<<<grid_for(kc, 64), 64, 0, st>>>(kc, hmax, hash_id, (uint32_t)ss, d_blob, d_sib + p0, d_H + p0, d_idx + p0, d_lc + 8 * p0,
                                                                 d_lh + (uint64_t)hw * p0, comc, d_root, d_ok + p0);
            ctx->launches += 3;
        }
        dfree(fold_mem, st);
    }
    cudaMemcpyAsync(ok, d_ok, k, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { dfree(mem, st); CUDA_TRY(cudaGetLastError()); return DAPOL_ERR_CUDA; }
    dfree(mem, st);
    // R::verify: one GPU batch per aggregate size.  The Padding policy fills the missing parties with
    // commit(0, 1) = B_blinding (padding.rs:176-180).
    uint8_t com_padding[32];
    {
        ge bb;
        uint32_t w[8];
        ge_bblinding(bb);
        ge_compress(w, bb);
        memcpy(com_padding, w, 32);
    }
    for (auto &kv : by_m) {
        const uint64_t m = kv.first, plen = dapol_rangeproof_size(64, (int)m);
        const std::vector<Item> &items = kv.second;
        std::vector<uint8_t> hp(items.size() * plen), hc(items.size() * m * 32), hok(items.size());
        for (size_t i = 0; i < items.size(); i++) {
            const Item &it = items[i];
            memcpy(hp.data() + i * plen, proofs + it.off, plen);
            const uint8_t *sib = proofs + pp[it.p].sib_off;
            for (uint64_t j = 0; j < m; j++) memcpy(hc.data() + (i * m + j) * 32, j < it.count ? sib + ss * (it.first + j) : com_padding, 32);
        }
        int rc = dapol_rangeproof_verify_batch(ctx, 64, (int)m, items.size(), hp.data(), plen, hc.data(), hok.data());
        if (rc) return rc;
        for (size_t i = 0; i < items.size(); i++) if (!hok[i]) ok[items[i].p] = 0;
    }
    return DAPOL_OK;
}

// Inclusion proofs streamed to a file (the reference's TODO "write the proofs to a local file", src/dapol/mod.rs:250):
// dapol_prove_batch in chunks, proof i at byte i * size.
extern "C" int dapol_prove_to_file(const dapol_tree *t, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                   const uint8_t seed[32], uint64_t chunk, const char *path, uint64_t *proof_size) {
    if (!t || !leaf_idx || !seed || !k || !path) return DAPOL_ERR_BAD_ARG;
    const uint64_t size = dapol_inclusion_proof_size_d(dapol_total_height(t), aggregation_factor, policy, t->hash_id);
    if (proof_size) *proof_size = size;
    if (size == 0) return DAPOL_ERR_BAD_ARG;
    if (chunk == 0) chunk = 8192;
    FILE *f = fopen(path, "wb");
    if (!f) return DAPOL_ERR_IO;
    std::vector<uint8_t> buf(std::min(chunk, k) * size);
    int rc = DAPOL_OK;
    for (uint64_t s = 0; s < k && rc == DAPOL_OK; s += chunk) {
        const uint64_t n = std::min(chunk, k - s);
        uint64_t got = 0;
        rc = dapol_prove_batch(t, n, leaf_idx + s, aggregation_factor, policy, seed, buf.data(), buf.size(), &got);
        if (rc == DAPOL_OK && fwrite(buf.data(), 1, n * size, f) != n * size) rc = DAPOL_ERR_IO;
    }
    if (fclose(f) != 0 && rc == DAPOL_OK) rc = DAPOL_ERR_IO;
    return rc;
}

// ================================================================================================ batch proofs
// ONE DapolProof for several leaves: Dapol::generate_proof_batch (src/dapol/mod.rs:172-190) and DapolProof::verify_batch
// (src/proof/mod.rs:49-54); test shape src/proof/tests.rs:6-35.  smtree's get_merkle_path_ref_batch / MerkleProof::verify_batch
// are UPSTREAM-RECALL: level by level from the leaves up, left to right, a node's sibling is part of the
// proof only if it is not itself on the way up from the batch.  The sibling order and the MerkleProof framing live in
// batch_sibling_plan / batch_merkle_bytes / parse_batch_proof only.
struct SibRef { int h; uint64_t idx; };
static std::vector<SibRef> batch_sibling_plan(int height, uint64_t k, const uint64_t *idx) {
    std::vector<SibRef> plan;
    std::vector<uint64_t> cur(idx, idx + k), nxt;
    for (int h = height; h >= 1; h--) {
        const size_t n = cur.size();
        for (size_t i = 0; i < n; i++) {
            const uint64_t s = cur[i] ^ 1;
            if (!((i > 0 && cur[i - 1] == s) || (i + 1 < n && cur[i + 1] == s))) plan.push_back({h, s});
        }
        nxt.clear();
        for (uint64_t x : cur) if (nxt.empty() || nxt.back() != x >> 1) nxt.push_back(x >> 1);
        cur.swap(nxt);
    }
    return plan;
}
static uint64_t batch_merkle_bytes(uint64_t H, uint64_t k, uint64_t nsib, uint64_t dl) { return 2 + 8 + k * ((H + 7) / 8) + 8 + (32 + dl) * nsib; }
static uint64_t range_part_bytes(uint64_t nsib, uint64_t agg, int policy) {
    std::vector<AggGroup> g;
    uint64_t sf;
    if (policy_plan(nsib, agg, policy, g, sf)) return 0;
    uint64_t sz = policy == DAPOL_POLICY_SPLITTING ? 2 : 0;
    for (auto &x : g) sz += 8 + dapol_rangeproof_size(64, (int)x.m);
    return sz + 8 + SINGLE_PROOF_BYTE_NUM * (nsib - sf);
}
extern "C" uint64_t dapol_batch_proof_size(int height, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy) {
    return dapol_batch_proof_size_d(height, k, leaf_idx, aggregation_factor, policy, DAPOL_HASH_BLAKE3);
}
extern "C" uint64_t dapol_batch_proof_size_d(int height, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy, int hash_id) {
    const uint64_t dl = (uint64_t)dapol_digest_len(hash_id);
    if (!dl || !leaf_idx || k == 0 || height < 0 || height > 64) return 0;
    if (k == 1) return dapol_inclusion_proof_size_d(height, aggregation_factor, policy, hash_id);
    for (uint64_t i = 0; i < k; i++) if ((i && leaf_idx[i] <= leaf_idx[i - 1]) || (height < 64 && (leaf_idx[i] >> height))) return 0;
    const uint64_t nsib = batch_sibling_plan(height, k, leaf_idx).size(), rs = range_part_bytes(nsib, aggregation_factor, policy);
    return rs ? rs + batch_merkle_bytes((uint64_t)height, k, nsib, dl) : 0;
}
// node (level h, tree index x) of the store: levels are kept in tree order, so a binary search over the level's indexes finds it
__global__ void k_fetch_nodes(uint64_t n, const int *lvl, const uint64_t *idx, NodeStore ns, const uint64_t *level_off, const uint64_t *level_n,
                              uint64_t *o_v, uint32_t *o_r, uint32_t *o_c, uint32_t *o_h, uint8_t *o_pad, int *not_found, uint32_t *o_hh) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t off = level_off[lvl[j]];
    int64_t slot = find_leaf_slot(ns.idx + off, level_n[lvl[j]], idx[j]);
    if (slot < 0) { *not_found = 1; return; }
    const uint64_t g = off + (uint64_t)slot;
    uint32_t w[8];
    o_v[j] = ns.v[g];
    load8(w, ns.r + 8 * g); store8(o_r + 8 * j, w);
    load8(w, ns.comc + 8 * g); store8(o_c + 8 * j, w);
    load8(w, ns.hash + 8 * g); store8(o_h + 8 * j, w);
    if (o_hh && ns.hash_hi) { load8(w, ns.hash_hi + 8 * g); store8(o_hh + 8 * j, w); }
    o_pad[j] = ns.is_pad[g];
}
// nonce key of a batch of more than one leaf: the tree's prover key chained over the leaf indexes, 64 per link; stream 0
static void batch_nonce_key(uint8_t key[32], uint64_t k, const uint64_t *idx) {
    static const char label[] = "dapol-b200 batch proof nonce key v1";
    dapol_hasher hs;
    uint32_t st[8];
    uint8_t le[8];
    hasher_init(hs, DAPOL_HASH_BLAKE3);
    hasher_update(hs, reinterpret_cast<const uint8_t *>(label), 35);
    hasher_update(hs, key, 32);
    for (int b = 0; b < 8; b++) le[b] = (uint8_t)(k >> (8 * b));
    hasher_update(hs, le, 8);
    hasher_final(hs, st);
    for (uint64_t i = 0; i < k; i += 64) {
        hasher_init(hs, DAPOL_HASH_BLAKE3);
        hasher_update_words(hs, st, 8);
        for (uint64_t j = i; j < k && j < i + 64; j++) {
            for (int b = 0; b < 8; b++) le[b] = (uint8_t)(idx[j] >> (8 * b));
            hasher_update(hs, le, 8);
        }
        hasher_final(hs, st);
    }
    memcpy(key, st, 32);
}
extern "C" int dapol_generate_proof_batch(const dapol_tree *t, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                          const uint8_t seed[32], uint8_t *out, uint64_t cap, uint64_t *proof_size) {
    if (!t || !leaf_idx || !seed || !k) return DAPOL_ERR_BAD_ARG;
    if (k == 1) return dapol_prove_batch(t, 1, leaf_idx, aggregation_factor, policy, seed, out, cap, proof_size);  // mod.rs:167-169
    if (t->top) return DAPOL_ERR_BAD_ARG;  // a batch may straddle shards: build it on the rank that holds a single tree
    dapol_ctx *ctx = t->ctx;
    const int H = t->height;
    const uint64_t size = dapol_batch_proof_size_d(H, k, leaf_idx, aggregation_factor, policy, t->hash_id);
    const bool b2b = t->ns.hash_hi != nullptr;
    if (proof_size) *proof_size = size;
    if (!size) return DAPOL_ERR_BAD_ARG;  // unsorted indexes (smtree rejects), aggregation_factor > #siblings (reference: slice panic)
    if (!out || cap < size) return DAPOL_ERR_BUFFER;
    const std::vector<SibRef> plan = batch_sibling_plan(H, k, leaf_idx);
    const uint64_t nsib = plan.size(), nf_items = k + nsib;
    std::vector<AggGroup> groups;
    uint64_t sf;
    policy_plan(nsib, aggregation_factor, policy, groups, sf);
    const uint64_t nsingle = nsib - sf;
    uint8_t key[32];
    prover_nonce_key(key, seed, t, policy, aggregation_factor, (uint64_t)H);
    batch_nonce_key(key, k, leaf_idx);
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<int> h_lvl(nf_items);
    std::vector<uint64_t> h_idx(nf_items);
    for (uint64_t i = 0; i < k; i++) { h_lvl[i] = H; h_idx[i] = leaf_idx[i]; }
    for (uint64_t j = 0; j < nsib; j++) { h_lvl[k + j] = plan[j].h; h_idx[k + j] = plan[j].idx; }
    uint64_t max_m = 1, agg_bytes = 0;
    for (auto &g : groups) { max_m = std::max(max_m, g.m); agg_bytes += dapol_rangeproof_size(64, (int)g.m); }
    const uint64_t rows = std::max<uint64_t>(max_m, std::max<uint64_t>(nsingle, 1));
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = Arena::need(nf_items, 4) + 2 * Arena::need(nf_items, 8) + 4 * Arena::need(nf_items, 32) + Arena::need(nf_items, 1) + 3 * 256 + Arena::need(65, 8) +
              Arena::need(rows, 8) + Arena::need(rows, 32) + 2 * Arena::need(rows, 8) + Arena::need(agg_bytes + 1, 1) + Arena::need(nsingle + 1, SINGLE_PROOF_BYTE_NUM);
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    int *d_lvl = ar.take<int>(nf_items);
    uint64_t *d_idx = ar.take<uint64_t>(nf_items), *d_v = ar.take<uint64_t>(nf_items);
    uint32_t *d_r = ar.take<uint32_t>(nf_items * 8), *d_c = ar.take<uint32_t>(nf_items * 8), *d_h = ar.take<uint32_t>(nf_items * 8);
    uint32_t *d_hh = ar.take<uint32_t>(nf_items * 8);
    uint8_t *d_pad = ar.take<uint8_t>(nf_items);
    int *d_nf = ar.take<int>(1);
    uint64_t *d_zero = ar.take<uint64_t>(1), *d_level_n = ar.take<uint64_t>(65);
    uint64_t *g_v = ar.take<uint64_t>(rows);
    uint32_t *g_r = ar.take<uint32_t>(rows * 8);
    uint64_t *g_s = ar.take<uint64_t>(rows), *g_b = ar.take<uint64_t>(rows);
    uint8_t *d_agg = ar.take<uint8_t>(agg_bytes + 1), *d_single = ar.take<uint8_t>((nsingle + 1) * SINGLE_PROOF_BYTE_NUM);
#define FAILB(code) do { dfree(mem, st); return (code); } while (0)
    cudaMemsetAsync(d_nf, 0, 4, st); cudaMemsetAsync(d_zero, 0, 8, st);
    cudaMemcpyAsync(d_lvl, h_lvl.data(), nf_items * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_idx, h_idx.data(), nf_items * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_level_n, t->level_n.data(), (H + 1) * 8, cudaMemcpyHostToDevice, st);
    k_fetch_nodes<<<grid_for(nf_items, 128), 128, 0, st>>>(nf_items, d_lvl, d_idx, t->ns, t->d_level_off, d_level_n, d_v, d_r, d_c, d_h, d_pad, d_nf, b2b ? d_hh : nullptr);
    ctx->launches++;
    int nf = 0;
    std::vector<uint8_t> h_pad(nf_items);
    cudaMemcpyAsync(&nf, d_nf, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_pad.data(), d_pad, nf_items, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) FAILB(DAPOL_ERR_CUDA);
    for (uint64_t i = 0; i < k; i++) if (h_pad[i]) nf = 1;  // a padding node is not a leaf of the liability set (reference: None)
    if (nf) FAILB(DAPOL_ERR_NOT_FOUND);
    const uint64_t *s_v = d_v + k;
    const uint32_t *s_r = d_r + 8 * k;
    uint64_t q = 0, off = 0;
    std::vector<uint64_t> agg_off;
    for (auto &g : groups) {
        k_gather_group<<<grid_for(g.m, 128), 128, 0, st>>>(1, (int)nsib, g.start, g.count, g.m, s_v, s_r, g_v, g_r);
        k_proof_streams<<<1, 128, 0, st>>>(1, 1, q, d_zero, g_s, g_b);
        ctx->launches += 2;
        int rc = dapol_rp_prove_dev(ctx, 64, (int)g.m, 1, g_v, reinterpret_cast<const uint8_t *>(g_r), key, g_s, g_b, d_agg + off);
        if (rc) FAILB(rc);
        agg_off.push_back(off);
        off += dapol_rangeproof_size(64, (int)g.m);
        q++;
    }
    if (nsingle) {
        k_gather_group<<<grid_for(nsingle, 128), 128, 0, st>>>(1, (int)nsib, sf, nsingle, nsingle, s_v, s_r, g_v, g_r);
        k_proof_streams<<<grid_for(nsingle, 128), 128, 0, st>>>(1, nsingle, q, d_zero, g_s, g_b);
        ctx->launches += 2;
        int rc = dapol_rp_prove_dev(ctx, 64, 1, nsingle, g_v, reinterpret_cast<const uint8_t *>(g_r), key, g_s, g_b, d_single);
        if (rc) FAILB(rc);
    }
    std::vector<uint8_t> h_agg(agg_bytes + 1), h_single(nsingle * SINGLE_PROOF_BYTE_NUM + 1), h_c(nsib * 32 + 1), h_h(nsib * 32 + 1), h_hh(nsib * 32 + 1);
    if (agg_bytes) cudaMemcpyAsync(h_agg.data(), d_agg, agg_bytes, cudaMemcpyDeviceToHost, st);
    if (nsingle) cudaMemcpyAsync(h_single.data(), d_single, nsingle * SINGLE_PROOF_BYTE_NUM, cudaMemcpyDeviceToHost, st);
    if (nsib) { cudaMemcpyAsync(h_c.data(), d_c + 8 * k, nsib * 32, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(h_h.data(), d_h + 8 * k, nsib * 32, cudaMemcpyDeviceToHost, st); }
    if (nsib && b2b) cudaMemcpyAsync(h_hh.data(), d_hh + 8 * k, nsib * 32, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) FAILB(DAPOL_ERR_CUDA);
    dfree(mem, st);
#undef FAILB
    uint8_t *o = out;
    if (policy == DAPOL_POLICY_SPLITTING) { put_be(o, groups.size(), 2); o += 2; }
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const uint64_t plen = dapol_rangeproof_size(64, (int)groups[gi].m);
        put_be(o, plen, 8); o += 8;
        memcpy(o, h_agg.data() + agg_off[gi], plen); o += plen;
    }
    put_be(o, nsingle, 8); o += 8;
    memcpy(o, h_single.data(), nsingle * SINGLE_PROOF_BYTE_NUM); o += nsingle * SINGLE_PROOF_BYTE_NUM;
    // MerkleProof::serialize of a batch proof: height, #indexes, the path bits of every index, #siblings, siblings
    const uint64_t Hu = (uint64_t)H, nb = (Hu + 7) / 8;
    put_be(o, Hu, 2); o += 2;
    put_be(o, k, 8); o += 8;
    for (uint64_t i = 0; i < k; i++) if (nb) { put_be(o, Hu == 64 ? leaf_idx[i] : leaf_idx[i] << (8 * nb - Hu), (int)nb); o += nb; }
    put_be(o, nsib, 8); o += 8;
    for (uint64_t s = 0; s < nsib; s++) {
        memcpy(o, h_c.data() + s * 32, 32); memcpy(o + 32, h_h.data() + s * 32, 32); o += 64;
        if (b2b) { memcpy(o, h_hh.data() + s * 32, 32); o += 32; }
    }
    return DAPOL_OK;
}

// ---- verify_batch: MerkleProof::verify_batch as level-synchronous merges on the device (DapolProofNode::merge,
// src/proof/node.rs:56-69), then R::verify over the siblings' commitments in proof order.
struct BatchNode { uint32_t ext[32], c[8], h[16]; };  // h: 8 words, or 16 (lo | hi) for 64-byte digests
// working nodes 0..k-1 = the leaves, k.. = the proof's siblings: decompress (non-canonical points reject, proof/node.rs:81-102)
__global__ void __launch_bounds__(64) k_batch_init(uint64_t n, const uint32_t *coms, const uint32_t *hashes, BatchNode *nodes, int *bad, int hw /* words per hash */) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t c[8], h[8];
    load8(c, coms + 8 * j); load8(h, hashes + (uint64_t)hw * j);
    ge p;
    if (!ge_decompress(p, c)) { *bad = 1; return; }
    rp_store_ext(nodes[j].ext, p);
    store8(nodes[j].c, c); store8(nodes[j].h, h);
    if (hw == 16) { load8(h, hashes + 16 * j + 8); store8(nodes[j].h + 8, h); }
}
// one level: parent j = merge(nodes[left[j]], nodes[right[j]]) written at nodes[dst0 + j]
__global__ void __launch_bounds__(64) k_batch_merge_level(uint64_t n, const uint32_t *left, const uint32_t *right, uint64_t dst0, BatchNode *nodes, int hash_id) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const BatchNode &l = nodes[left[j]], &r = nodes[right[j]];
    ge a, b, s;
    rp_load_ext(a, l.ext); rp_load_ext(b, r.ext);
    ge_add(s, a, b);
    uint32_t cl[8], cr[8], hl[8], hr[8], hh[8], cc[8];
    load8(cl, l.c); load8(cr, r.c); load8(hl, l.h); load8(hr, r.h);
    BatchNode &o = nodes[dst0 + j];
    if (hash_id == DAPOL_HASH_BLAKE2B) {
        uint32_t hlu[8], hru[8], hu[8];
        load8(hlu, l.h + 8); load8(hru, r.h + 8);
        dapol_b2b_hash192(hh, hu, cl, cr, hl, hlu, hr, hru);
        store8(o.h + 8, hu);
    } else dapol_hash128(hash_id, hh, cl, cr, hl, hr);
    ge_compress(cc, s);
    rp_store_ext(o.ext, s);
    store8(o.c, cc); store8(o.h, hh);
}
struct ParsedBatch {
    bool ok = false;
    std::vector<std::pair<uint64_t, uint64_t>> agg;
    uint64_t nind = 0, ind_off = 0, H = 0, nsib = 0, sib_off = 0;
    std::vector<uint64_t> idx;
};
static ParsedBatch parse_batch_proof(const uint8_t *p, uint64_t len, int policy, uint64_t ss) {
    ParsedBatch r;
    uint64_t pos = 0, nagg = 1;
#define NEED(n) do { if (len - pos < (uint64_t)(n)) return r; } while (0)
    if (policy == DAPOL_POLICY_SPLITTING) { NEED(2); nagg = get_be(p, 2); pos += 2; if (nagg > 64) return r; }
    for (uint64_t i = 0; i < nagg; i++) {
        NEED(8);
        uint64_t sz = get_be(p + pos, 8); pos += 8;
        if (sz > len - pos) return r;
        r.agg.push_back({pos, sz}); pos += sz;
    }
    NEED(8);
    r.nind = get_be(p + pos, 8); pos += 8;
    if ((len - pos) / SINGLE_PROOF_BYTE_NUM < r.nind) return r;
    r.ind_off = pos; pos += SINGLE_PROOF_BYTE_NUM * r.nind;
    NEED(10);
    r.H = get_be(p + pos, 2); pos += 2;
    const uint64_t k = get_be(p + pos, 8); pos += 8;
    const uint64_t nb = (r.H + 7) / 8;
    if (r.H > 64 || k == 0 || (nb && (len - pos) / nb < k)) return r;
    for (uint64_t i = 0; i < k; i++) {
        uint64_t x = nb ? get_be(p + pos, (int)nb) : 0;
        if (nb && r.H != 64) x >>= (8 * nb - r.H);
        if (i && x <= r.idx.back()) return r;
        r.idx.push_back(x); pos += nb;
    }
    NEED(8);
    r.nsib = get_be(p + pos, 8); pos += 8;
    if ((len - pos) / ss < r.nsib || r.nind > r.nsib) return r;
    r.sib_off = pos;
#undef NEED
    r.ok = true;
    return r;
}
extern "C" int dapol_proof_verify_batch(dapol_ctx *ctx, int hash_id, int policy, uint64_t k, const uint8_t root_com[32], const uint8_t root_hash[32],
                                        const uint8_t *leaf_coms, const uint8_t *leaf_hashes, const uint8_t *proof, uint64_t proof_len, uint8_t *ok) {
    if (!ctx || !root_com || !root_hash || !leaf_coms || !leaf_hashes || !proof || !ok || !k) return DAPOL_ERR_BAD_ARG;
    const uint64_t dl = (uint64_t)dapol_digest_len(hash_id), ss = 32 + dl;  // root_hash / leaf_hashes: dl bytes each
    if (!dl) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (policy != DAPOL_POLICY_PADDING && policy != DAPOL_POLICY_SPLITTING) return DAPOL_ERR_BAD_ARG;
    *ok = 0;
    const ParsedBatch pb = parse_batch_proof(proof, proof_len, policy, ss);
    if (!pb.ok || pb.idx.size() != k) return DAPOL_OK;  // malformed bytes / wrong number of leaves: a reject, never an error
    const std::vector<SibRef> plan = batch_sibling_plan((int)pb.H, k, pb.idx.data());
    if (plan.size() != pb.nsib) return DAPOL_OK;
    // merge plan: working nodes [0, k) leaves, [k, k + nsib) siblings, then the parents level by level
    const uint64_t nsib = pb.nsib;
    std::vector<uint32_t> left, right;
    std::vector<uint64_t> lvl_first, lvl_count;
    {
        std::vector<std::pair<uint64_t, uint32_t>> cur, nxt;  // (tree index, working node)
        for (uint64_t i = 0; i < k; i++) cur.push_back({pb.idx[i], (uint32_t)i});
        uint64_t sib_at = 0, next_node = k + nsib;
        for (int h = (int)pb.H; h >= 1; h--) {
            nxt.clear();
            lvl_first.push_back(left.size());
            for (size_t i = 0; i < cur.size();) {
                uint32_t me = cur[i].second, other;
                size_t step = 1;
                if (i + 1 < cur.size() && cur[i + 1].first == (cur[i].first ^ 1)) { other = cur[i + 1].second; step = 2; }
                else other = (uint32_t)(k + sib_at++);  // plan order == consumption order: level by level, left to right
                if (cur[i].first & 1) { left.push_back(other); right.push_back(me); } else { left.push_back(me); right.push_back(other); }
                nxt.push_back({cur[i].first >> 1, (uint32_t)next_node++});
                i += step;
            }
            lvl_count.push_back(left.size() - lvl_first.back());
            cur.swap(nxt);
        }
        if (sib_at != nsib) return DAPOL_OK;
    }
    const uint64_t n_work = k + nsib + left.size();
    if (n_work >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<uint8_t> h_c((k + nsib) * 32), h_h((k + nsib) * dl);
    memcpy(h_c.data(), leaf_coms, k * 32); memcpy(h_h.data(), leaf_hashes, k * dl);
    for (uint64_t s = 0; s < nsib; s++) {
        memcpy(h_c.data() + (k + s) * 32, proof + pb.sib_off + ss * s, 32);
        memcpy(h_h.data() + (k + s) * dl, proof + pb.sib_off + ss * s + 32, dl);
    }
    uint8_t *mem = nullptr;
    Arena ar;
    ar.size = 3 * Arena::need(k + nsib, 32) + Arena::need(n_work, sizeof(BatchNode)) + 2 * Arena::need(left.size() + 1, 4) + 256;
    CUDA_TRY(dmalloc(&mem, ar.size, st));
    ar.base = mem;
    uint32_t *d_c = ar.take<uint32_t>((k + nsib) * 8), *d_h = ar.take<uint32_t>((k + nsib) * 16);
    BatchNode *nodes = ar.take<BatchNode>(n_work);
    uint32_t *d_l = ar.take<uint32_t>(left.size() + 1), *d_r = ar.take<uint32_t>(left.size() + 1);
    int *d_bad = ar.take<int>(1), bad = 0;
    cudaMemsetAsync(d_bad, 0, 4, st);
    cudaMemcpyAsync(d_c, h_c.data(), h_c.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_h, h_h.data(), h_h.size(), cudaMemcpyHostToDevice, st);
    if (!left.empty()) {
        cudaMemcpyAsync(d_l, left.data(), left.size() * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_r, right.data(), right.size() * 4, cudaMemcpyHostToDevice, st);
    }
    k_batch_init<<<grid_for(k + nsib, 64), 64, 0, st>>>(k + nsib, d_c, d_h, nodes, d_bad, (int)(dl / 4));
    ctx->launches++;
    uint64_t dst = k + nsib;
    for (size_t l = 0; l < lvl_first.size(); l++) {
        if (!lvl_count[l]) continue;
        k_batch_merge_level<<<grid_for(lvl_count[l], 64), 64, 0, st>>>(lvl_count[l], d_l + lvl_first[l], d_r + lvl_first[l], dst, nodes, hash_id);
        ctx->launches++;
        dst += lvl_count[l];
    }
    uint32_t top[24];
    const BatchNode *root_node = nodes + (n_work - 1);  // height 0: the single leaf itself (k == 1, no merges)
    cudaMemcpyAsync(top, root_node->c, 32, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(top + 8, root_node->h, dl, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    dfree(mem, st);
    CUDA_TRY(e);
    if (bad || memcmp(top, root_com, 32) || memcmp(top + 8, root_hash, dl)) return DAPOL_OK;
    // R::verify on the siblings' commitments (padding.rs:168-197 / splitting.rs:180-211)
    const uint64_t n_agg_coms = nsib - pb.nind;
    std::vector<AggGroup> groups;
    uint64_t sf;
    if (policy_plan(nsib, n_agg_coms, policy, groups, sf) || groups.size() > pb.agg.size() || (policy == DAPOL_POLICY_PADDING && pb.agg.size() != 1)) return DAPOL_OK;
    uint8_t com_padding[32];
    {
        ge bb;
        uint32_t w[8];
        ge_bblinding(bb);
        ge_compress(w, bb);
        memcpy(com_padding, w, 32);
    }
    const uint8_t *sib = proof + pb.sib_off;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const AggGroup &g = groups[gi];
        if (pb.agg[gi].second != dapol_rangeproof_size(64, (int)g.m)) return DAPOL_OK;
        std::vector<uint8_t> hc(g.m * 32);
        uint8_t one = 0;
        for (uint64_t j = 0; j < g.m; j++) memcpy(hc.data() + j * 32, j < g.count ? sib + ss * (g.start + j) : com_padding, 32);
        int rc = dapol_rangeproof_verify_batch(ctx, 64, (int)g.m, 1, proof + pb.agg[gi].first, pb.agg[gi].second, hc.data(), &one);
        if (rc) return rc;
        if (!one) return DAPOL_OK;
    }
    if (pb.nind) {
        std::vector<uint8_t> hc(pb.nind * 32), oks(pb.nind);
        for (uint64_t s = 0; s < pb.nind; s++) memcpy(hc.data() + s * 32, sib + ss * (n_agg_coms + s), 32);
        int rc = dapol_rangeproof_verify_batch(ctx, 64, 1, pb.nind, proof + pb.ind_off, SINGLE_PROOF_BYTE_NUM, hc.data(), oks.data());
        if (rc) return rc;
        for (uint8_t v : oks) if (!v) return DAPOL_OK;
    }
    *ok = 1;
    return DAPOL_OK;
}
