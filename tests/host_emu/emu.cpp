// TEST HARNESS ONLY: compiles the product's device headers (dapol_b200/csrc/*.cuh) for the HOST with the
// PTX carry-chain instructions emulated (fe25519.cuh EMU_* macros) and drives the per-thread kernel bodies
// in serial loops, so the arithmetic and tree logic can be checked on a CPU-only machine against the oracle.
// Never linked into or loaded by the product library; the GPU tests exercise the real kernels.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../dapol_b200/csrc/tree_kernels.cuh"

#define EX extern "C" __attribute__((visibility("default")))

static void w2b(uint8_t *b, const uint32_t *w, int n) { memcpy(b, w, 4 * n); }
static void b2w(uint32_t *w, const uint8_t *b, int n) { memcpy(w, b, 4 * n); }

EX void emu_fe_op(int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fe x, y, r;
    b2w(x.v, a, 8); b2w(y.v, b, 8);
    switch (op) {
        case 0: fe_mul(r, x, y); break;
        case 1: fe_sq(r, x); break;
        case 2: fe_add(r, x, y); break;
        case 3: fe_sub(r, x, y); break;
        case 4: fe_invert(r, x); break;
        case 5: fe_neg(r, x); break;
        case 6: fe_pow22523(r, x); break;
        default: fe_set0(r);
    }
    fe_tobytes(out, r);
}
EX void emu_mul_wide(const uint8_t a[32], const uint8_t b[32], uint8_t out[64], int sq) {
    uint32_t x[8], y[8], R[16];
    b2w(x, a, 8); b2w(y, b, 8);
    if (sq) sqr_wide_8(R, x); else mul_wide_8x8(R, x, y);
    w2b(out, R, 16);
}
EX void emu_sc_op(int op, const uint8_t a[64], const uint8_t b[32], uint8_t out[32]) {
    sc x, y, r;
    b2w(x.v, a, 8); b2w(y.v, b, 8);
    switch (op) {
        case 0: sc_mul(r, x, y); break;           // x any, y < l
        case 1: sc_add(r, x, y); break;
        case 2: sc_sub(r, x, y); break;
        case 3: sc_invert(r, x); break;
        case 4: { uint32_t w[16]; b2w(w, a, 16); sc_from_wide(r, w); break; }
        case 5: sc_reduce256(r, x); break;
        case 6: sc_neg(r, x); break;
        case 7: { uint32_t w[16]; sc h; b2w(w, a, 16); sc_from_wide_with_half(r, h, w); break; }  // the wide reduction ...
        case 8: { uint32_t w[16]; sc h; b2w(w, a, 16); sc_from_wide_with_half(h, r, w); break; }  // ... and its half
        case 9: sc_half256(r, x); break;
        default: sc_set_u64(r, 0);
    }
    w2b(out, r.v, 8);
}
EX int emu_sc_is_canonical(const uint8_t a[32]) { uint32_t w[8]; b2w(w, a, 8); return sc_is_canonical(w); }

static void scalarmult(ge &r, const uint32_t k[8], const ge &p) {
    ge acc; ge_identity(acc);
    for (int i = 255; i >= 0; i--) { ge_dbl(acc, acc); if ((k[i >> 5] >> (i & 31)) & 1) ge_add(acc, acc, p); }
    r = acc;
}
EX void emu_scalarmult_base(const uint8_t s[32], uint8_t out[32], int use_bbl) {
    uint32_t k[8], c[8]; b2w(k, s, 8);
    ge b, r;
    if (use_bbl) ge_bblinding(b); else ge_basepoint(b);
    scalarmult(r, k, b);
    ge_compress(c, r); w2b(out, c, 8);
}
EX int emu_decompress_recompress(const uint8_t s[32], uint8_t out[32]) {
    uint32_t w[8], c[8]; b2w(w, s, 8);
    ge p;
    if (!ge_decompress(p, w)) return 0;
    ge q; ge_dbl(q, p); ge_sub(q, q, p);  // exercise dbl/sub: 2p - p
    ge_compress(c, q); w2b(out, c, 8);
    return 1;
}
EX void emu_from_uniform(const uint8_t b[64], uint8_t out[32]) {
    uint32_t w[16], c[8]; b2w(w, b, 16);
    ge p; ge_from_uniform(p, w);
    ge_compress(c, p); w2b(out, c, 8);
}
EX void emu_hash32(int hash_id, const uint8_t in[32], uint8_t out[32]) {
    uint32_t w[8], o[8]; b2w(w, in, 8); dapol_hash32(hash_id, o, w); w2b(out, o, 8);
}
EX void emu_hash128(int hash_id, const uint8_t in[128], uint8_t out[32]) {
    uint32_t w[32], o[8]; b2w(w, in, 32); dapol_hash128(hash_id, o, w, w + 8, w + 16, w + 24); w2b(out, o, 8);
}
EX int emu_hash_bytes(int hash_id, const uint8_t *in, uint32_t len, uint8_t out[32]) {
    dapol_hasher h; hasher_init(h, hash_id);
    hasher_update(h, in, len / 2); hasher_update(h, in + len / 2, len - len / 2);
    uint32_t o[8]; int rc = hasher_final(h, o); w2b(out, o, 8); return rc;
}
EX void emu_chacha(const uint8_t key[32], uint64_t counter, uint64_t stream, uint8_t out[64]) {
    uint32_t k[8], o[16]; b2w(k, key, 8); chacha20_block(o, k, counter, stream); w2b(out, o, 16);
}
EX void emu_keccak(uint8_t st[200]) { uint64_t s[25]; memcpy(s, st, 200); keccak_f1600(s); memcpy(st, s, 200); }

// ---- comb commit with tables built by the same body the init kernel uses
template <int W>
static void commit_w(uint64_t v, const uint8_t r[32], uint8_t out[32]) {
    constexpr int NWR = 253 / W + 1;
    const uint32_t half = 1u << (W - 1);
    static std::vector<ge_niels> tb, tbbl;
    if (tb.empty()) {
        tb.resize((size_t)NWR * half); tbbl.resize((size_t)NWR * half);
        if (W == 8) {  // the run builder of the wide windows (RUN consecutive multiples per thread, shared inversion)
            constexpr int RUN = W == 8 ? 16 : 1;
            std::vector<ge> bases(NWR);
            for (int which = 0; which < 2; which++) {
                comb_bases_body<W>(bases.data(), NWR, which);
                std::vector<ge_niels> &dst = which ? tbbl : tb;
                for (uint64_t t = 0; t < dst.size() / RUN; t++) comb_table_run_body<W, RUN>(t, dst.data(), NWR, bases.data());
            }
            // both builders must give identical table bytes
            std::vector<ge_niels> chk(tb.size());
            for (uint64_t t = 0; t < 300; t++) {
                comb_table_body<W>(t * 7, chk.data(), NWR, 1);
                if (memcmp(&chk[t * 7], &tbbl[t * 7], sizeof(ge_niels))) abort();
            }
        } else {
            for (uint64_t t = 0; t < tb.size(); t++) comb_table_body<W>(t, tb.data(), NWR, 0);
            for (uint64_t t = 0; t < tbbl.size(); t++) comb_table_body<W>(t, tbbl.data(), NWR, 1);
        }
    }
    // one-node store
    uint64_t idx = 0, vv = 0; uint32_t rr[8], cc[8], hh[8], ext[32], blind[8]; uint8_t pad = 0; uint32_t pos = 0;
    NodeStore ns{&idx, &vv, rr, cc, hh, ext, &pad};
    b2w(blind, r, 8);
    leaf_batch_body<W, 1>(0, 1, 1, ns, 0, &pos, 0, &v, blind, tb.data(), tbbl.data());
    w2b(out, cc, 8);
}
EX void emu_commit(int w, uint64_t v, const uint8_t r[32], uint8_t out[32]) {
    if (w == 4) commit_w<4>(v, r, out); else if (w == 5) commit_w<5>(v, r, out); else commit_w<8>(v, r, out);
}

// ---- serial tree build through the kernel bodies (mirrors the CUDA host orchestration)
struct EmuTree {
    int height; std::vector<uint64_t> level_off, level_n;
    std::vector<uint64_t> idx, v; std::vector<uint32_t> r, comc, hash, ext, hash_hi; std::vector<uint8_t> is_pad;
    uint64_t n_pads;
};
static int g_positional = 0;  // padding mode of the next emulated builds (dapol_ctx_set_padding_mode)
EX void emu_set_padding_mode(int positional) { g_positional = positional; }
EX EmuTree *emu_tree_build(int hash_id, int height, uint64_t n, const uint64_t *leaf_idx, const uint64_t *values,
                           const uint8_t *blind, const uint8_t pad_seed[32], uint64_t pad_base) {
    constexpr int W = 4;
    constexpr int NWR = 253 / W + 1;
    const uint32_t half = 1u << (W - 1);
    static std::vector<ge_niels> tb, tbbl;
    if (tb.empty()) {
        tb.resize((size_t)NWR * half); tbbl.resize((size_t)NWR * half);
        for (uint64_t t = 0; t < tb.size(); t++) comb_table_body<W>(t, tb.data(), NWR, 0);
        for (uint64_t t = 0; t < tbbl.size(); t++) comb_table_body<W>(t, tbbl.data(), NWR, 1);
    }
    // level sizes from the adjacent-leaf msb histogram (as the CUDA host code does)
    uint64_t hist[65] = {0};
    for (uint64_t k = 0; k < n; k++) { int bad = 0; int m = leaf_pair_msb(k, leaf_idx, height, &bad); if (bad) return nullptr; if (m >= 0) hist[m]++; }
    std::vector<uint64_t> n_real(height + 1), npads(height + 1, 0), nparents(height + 1, 0);
    { uint64_t acc = 0; for (int h = 0; h <= height; h++) { if (h >= 1) acc += hist[height - h]; n_real[h] = 1 + acc; } }
    if (n_real[height] != n) abort();
    auto *t = new EmuTree();
    t->height = height; t->level_off.resize(height + 1); t->level_n.resize(height + 1);
    uint64_t T = 1, total_pads = 0; t->level_off[0] = 0; t->level_n[0] = 1;
    for (int h = 1; h <= height; h++) {
        nparents[h] = n_real[h - 1]; t->level_n[h] = 2 * n_real[h - 1]; npads[h] = t->level_n[h] - n_real[h];
        t->level_off[h] = T; T += t->level_n[h]; total_pads += npads[h];
    }
    t->idx.assign(T, 0); t->v.assign(T, 0); t->r.assign(8 * T, 0); t->comc.assign(8 * T, 0); t->hash.assign(8 * T, 0);
    t->ext.assign(32 * T, 0); t->is_pad.assign(T, 0);
    NodeStore ns{t->idx.data(), t->v.data(), t->r.data(), t->comc.data(), t->hash.data(), t->ext.data(), t->is_pad.data()};
    const bool b2b = hash_id == DAPOL_HASH_BLAKE2B;  // 64-byte digests: the orchestration of tree_build_dev / dapol_launch_merges
    if (b2b) { t->hash_hi.assign(8 * T, 0); ns.hash_hi = t->hash_hi.data(); }
    std::vector<uint64_t> pad_dest(total_pads + 1), pad_rng(total_pads + 1);
    std::vector<std::vector<uint64_t>> real(height + 1);
    std::vector<std::vector<uint32_t>> pos(height + 1);
    real[height].assign(leaf_idx, leaf_idx + n);
    uint64_t ord = 0;
    PadStreams ps;
    memset(&ps, 0, sizeof ps);
    ps.positional = g_positional; ps.levels = height; ps.level0 = 0;
    for (int h = height; h >= 1; h--) {
        ps.start[h] = ord;
        uint64_t c = real[h].size();
        if (c != n_real[h]) abort();
        pos[h].resize(c); real[h - 1].assign(n_real[h - 1], ~0ull);
        uint64_t s = 0;
        for (uint64_t k = 0; k < c; k++) {
            uint64_t f = struct_flags_body(k, real[h].data(), c);
            struct_apply_body(k, real[h].data(), f, s, pos[h].data(), real[h - 1].data(), t->level_off[h], ns, pad_dest.data(), ord, pad_rng.data(), g_positional ? 0 : pad_base + ord, g_positional);
            s += f;
        }
        if ((s >> 32) != nparents[h] || (s & 0xffffffffull) != npads[h]) abort();
        ord += npads[h];
    }
    t->n_pads = total_pads;
    ps.start[0] = ord;
    uint32_t seed[8]; b2w(seed, pad_seed, 8);
    std::vector<uint32_t> bw(8 * n); b2w(bw.data(), blind, 8 * n);
    constexpr int BT = 3;  // units per emulated thread (exercises full and partial batches)
    auto threads = [](uint64_t units) { return (units + BT - 1) / BT; };
    if (height == 0) {
        uint32_t p0 = 0; leaf_batch_body<W, BT>(0, 1, 1, ns, 0, &p0, hash_id, values, bw.data(), tb.data(), tbbl.data()); t->idx[0] = leaf_idx[0];
        if (b2b) leafpad_hash_b2b_body(0, ns, 0);
        return t;
    }
    for (uint64_t i = 0, st = threads(n); i < st; i++)
        leaf_batch_body<W, BT>(i, st, n, ns, t->level_off[height], pos[height].data(), hash_id, values, bw.data(), tb.data(), tbbl.data());
    for (uint64_t g = 0, st = threads(total_pads); g < st; g++) pad_batch_body<W, BT>(g, st, total_pads, ns, pad_dest.data(), hash_id, seed, pad_rng.data(), tbbl.data(), ps);
    if (b2b) for (uint64_t g = 0; g < T; g++) leafpad_hash_b2b_body(g, ns, t->level_off[height]);
    // merges: sums per level, one compress pass over all internal nodes, hashes per level (as tree_build_dev)
    for (int h = height; h >= 1; h--)
        for (uint64_t j = 0; j < nparents[h]; j++)
            merge_sum_body(j, ns, t->level_off[h], t->level_off[h - 1], h - 1 == 0 ? nullptr : pos[h - 1].data());
    std::vector<uint32_t *> pos_ptr(height + 1, nullptr);
    for (int h = 1; h <= height; h++) pos_ptr[h] = pos[h].data();
    InternalMap im;
    memset(&im, 0, sizeof(im));
    im.levels = height; im.level_off = t->level_off.data(); im.pos = pos_ptr.data();
    uint64_t n_int = 0;
    for (int h = 0; h < height; h++) { im.start[h] = n_int; n_int += n_real[h]; }
    im.start[height] = n_int;
    for (uint64_t j = 0, st = threads(n_int); j < st; j++) compress_internal_body<BT>(j, st, n_int, ns, im);
    for (int h = height; h >= 1; h--)
        for (uint64_t j = 0; j < nparents[h]; j++) {
            if (b2b) merge_hash_b2b_body(j, ns, t->level_off[h], t->level_off[h - 1], h - 1 == 0 ? nullptr : pos[h - 1].data());
            else merge_hash_body(j, ns, t->level_off[h], t->level_off[h - 1], h - 1 == 0 ? nullptr : pos[h - 1].data(), hash_id);
        }
    return t;
}
// leaf derivation through the kernel bodies, with the sort/fix-point orchestration mirrored serially
#include <algorithm>
#include <numeric>
EX int emu_derive_leaves(int hash_id, uint64_t n, const uint8_t *iid_blob, const uint64_t *iid_off, const uint8_t *eid_blob,
                         const uint64_t *eid_off, const uint8_t *seed, uint32_t seed_len, int height, uint64_t *out_idx, uint8_t *out_blind,
                         uint64_t *err_pos) {
    std::vector<uint32_t> audit(8 * n), cur(8 * n), blind(8 * n), tries(n, 1);
    std::vector<uint64_t> cand(n);
    for (uint64_t i = 0; i < n; i++)
        if (derive_body(i, hash_id, iid_blob, iid_off, eid_blob, eid_off, seed, seed_len, height, audit.data(), cur.data(), cand.data(), blind.data())) return 16;
    uint64_t first_dup = ~0ull, first_fail = ~0ull;
    for (uint64_t i = 0; i < n; i++) for (uint64_t j = 0; j < i; j++) if (!memcmp(&audit[8 * i], &audit[8 * j], 32)) { first_dup = std::min(first_dup, i); break; }
    std::vector<uint32_t> who(n);
    for (;;) {
        std::iota(who.begin(), who.end(), 0u);
        std::stable_sort(who.begin(), who.end(), [&](uint32_t a, uint32_t b) { return cand[a] < cand[b]; });
        std::vector<uint64_t> sorted(n);
        for (uint64_t j = 0; j < n; j++) sorted[j] = cand[who[j]];
        uint64_t losers = 0;
        for (uint64_t j = 1; j < n; j++) {
            if (sorted[j] != sorted[j - 1]) continue;
            uint64_t u = who[j];
            if (tries[u] > 128) continue;
            if (rehash_body(u, hash_id, height, cur.data(), cand.data(), tries.data())) losers++;
            else { tries[u] = 129; first_fail = std::min(first_fail, u); }
        }
        if (!losers) break;
    }
    if (first_dup != ~0ull && first_dup <= first_fail) { *err_pos = first_dup; return 4; }
    if (first_fail != ~0ull) { *err_pos = first_fail; return 5; }
    memcpy(out_idx, cand.data(), 8 * n); memcpy(out_blind, blind.data(), 32 * n);
    return 0;
}
EX uint64_t emu_tree_level_size(EmuTree *t, int h) { return t->level_n[h]; }
EX uint64_t emu_tree_num_pads(EmuTree *t) { return t->n_pads; }
EX void emu_tree_level_copy(EmuTree *t, int h, uint64_t *idx, uint64_t *v, uint8_t *r, uint8_t *comc, uint8_t *hash, uint8_t *is_pad) {
    uint64_t o = t->level_off[h], n = t->level_n[h];
    memcpy(idx, &t->idx[o], 8 * n); memcpy(v, &t->v[o], 8 * n);
    memcpy(r, &t->r[8 * o], 32 * n); memcpy(comc, &t->comc[8 * o], 32 * n);
    if (t->hash_hi.empty()) memcpy(hash, &t->hash[8 * o], 32 * n);
    else for (uint64_t i = 0; i < n; i++) { memcpy(hash + 64 * i, &t->hash[8 * (o + i)], 32); memcpy(hash + 64 * i + 32, &t->hash_hi[8 * (o + i)], 32); }
    memcpy(is_pad, &t->is_pad[o], n);
}
EX void emu_tree_free(EmuTree *t) { delete t; }

// id / salt leaf hashes (opt-in mode) through the kernel body: audit ids as derive_body leaves them
EX int emu_leaf_id_hashes(int hash_id, uint64_t n, const uint8_t *iid_blob, const uint64_t *iid_off, const uint8_t *eid_blob, const uint64_t *eid_off,
                          const uint8_t *seed, uint32_t seed_len, int height, uint8_t *out) {
    std::vector<uint32_t> audit(8 * n), cur(8 * n), blind(8 * n), h(8 * n);
    std::vector<uint64_t> cand(n);
    int rc = 0;
    for (uint64_t i = 0; i < n; i++) rc |= derive_body(i, hash_id, iid_blob, iid_off, eid_blob, eid_off, seed, seed_len, height, audit.data(), cur.data(), cand.data(), blind.data());
    for (uint64_t i = 0; i < n; i++) rc |= leaf_id_hash_body(i, hash_id, audit.data(), eid_blob, eid_off, h.data());
    w2b(out, h.data(), 8 * n);
    return rc;
}
