// Per-thread bodies of the batched Bulletproofs range-proof prover and verifier.
//
// What the reference does one proof at a time on one CPU thread
//   (/root/reference/src/range/mod.rs:48-119 -> bulletproofs RangeProof::{prove,verify}_{single,multiple} over
//    merlin Transcript::new(&[]), regenerating BulletproofGens on every call)
// is split here into phase-synchronous passes over a BATCH of K proofs of one shape (n bits x m parties, N = n m):
//   thread-per-proof passes   : merlin transcript (STROBE-128 / Keccak-f), challenges, scalar inversions
//   thread-per-element passes : RNG draws, l(x) / r(x), folding of a and b, coefficient tables
//   CTA-per-proof passes      : inner products and the multi-scalar multiplications (IMAD-bound, dominant)
// All MSMs of the prover run over the ORIGINAL generators G_i, H_i (plus B, B_blinding): the inner-product rounds
// never fold generators, they fold the scalar coefficients (cu / cui tables) instead, so every MSM is fixed-base
// and uses signed-window tables (2^(W-1) multiples per window, affine Niels, 96 B) that live in HBM -- built once
// per context (the reference re-derives the generators for every proof, range/mod.rs:50,66,85,104).
#pragma once
#include "ge25519.cuh"
#include "hash_dev.cuh"

#define DAPOL_RP_ERR_CHALLENGE 1  // prover: a zero Fiat-Shamir challenge (bulletproofs aborts with MaliciousDealer)

// ------------------------------------------------------------------------------------------------ merlin
// merlin ^3.0.0 transcript.rs / strobe.rs: STROBE-128 (R = 166) over Keccak-f[1600]
struct merlin {
    uint64_t st[25];
    uint32_t pos, pos_begin;
};
#define STROBE_R 166u
#define SF_I 1u
#define SF_A 2u
#define SF_C 4u
#define SF_M 16u
#define SF_K 32u

DAPOL_HD_INLINE void strobe_xor(merlin &t, uint32_t pos, uint32_t b) { t.st[pos >> 3] ^= (uint64_t)(b & 0xffu) << (8 * (pos & 7)); }
DAPOL_HD_INLINE void strobe_run_f(merlin &t) {
    strobe_xor(t, t.pos, t.pos_begin);
    strobe_xor(t, t.pos + 1, 0x04);
    strobe_xor(t, STROBE_R + 1, 0x80);
    keccak_f1600(t.st);
    t.pos = 0; t.pos_begin = 0;
}
DAPOL_HD_INLINE void strobe_absorb_byte(merlin &t, uint32_t b) {
    strobe_xor(t, t.pos, b);
    t.pos++;
    if (t.pos == STROBE_R) strobe_run_f(t);
}
DAPOL_HD_INLINE void strobe_absorb(merlin &t, const uint8_t *d, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) strobe_absorb_byte(t, d[i]);
}
DAPOL_HD_INLINE void strobe_absorb_words(merlin &t, const uint32_t *w, uint32_t nbytes) {
    for (uint32_t i = 0; i < nbytes; i++) strobe_absorb_byte(t, w[i >> 2] >> (8 * (i & 3)));
}
DAPOL_HD_INLINE void strobe_begin_op(merlin &t, uint32_t flags) {
    uint32_t old = t.pos_begin;
    t.pos_begin = t.pos + 1;
    strobe_absorb_byte(t, old);
    strobe_absorb_byte(t, flags);
    if ((flags & (SF_C | SF_K)) && t.pos != 0) strobe_run_f(t);
}
DAPOL_HD_INLINE void strobe_init(merlin &t) {
#pragma unroll 1
    for (int i = 0; i < 25; i++) t.st[i] = 0;
    // [1, R+2, 1, 0, 1, 96] || "STROBEv1.0.2"
    const uint8_t hdr[18] = {1, STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (int i = 0; i < 18; i++) strobe_xor(t, i, hdr[i]);
    keccak_f1600(t.st);
    t.pos = 0; t.pos_begin = 0;
    const uint8_t lab[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
    strobe_begin_op(t, SF_M | SF_A);
    strobe_absorb(t, lab, 11);
}
// append_message(label, msg): meta_ad(label); meta_ad(le32(len), more); ad(msg)
DAPOL_HD_INLINE void tr_append_hdr(merlin &t, const char *label, uint32_t llen, uint32_t mlen) {
    strobe_begin_op(t, SF_M | SF_A);
    strobe_absorb(t, reinterpret_cast<const uint8_t *>(label), llen);
    uint32_t w = mlen;
    strobe_absorb_words(t, &w, 4);
    strobe_begin_op(t, SF_A);
}
DAPOL_HD_INLINE void tr_append_bytes(merlin &t, const char *label, uint32_t llen, const uint8_t *msg, uint32_t mlen) {
    tr_append_hdr(t, label, llen, mlen);
    strobe_absorb(t, msg, mlen);
}
DAPOL_HD_INLINE void tr_append_words(merlin &t, const char *label, uint32_t llen, const uint32_t *msg, uint32_t mlen) {
    tr_append_hdr(t, label, llen, mlen);
    strobe_absorb_words(t, msg, mlen);
}
DAPOL_HD_INLINE void tr_append_u64(merlin &t, const char *label, uint32_t llen, uint64_t x) {
    uint32_t w[2] = {(uint32_t)x, (uint32_t)(x >> 32)};
    tr_append_words(t, label, llen, w, 8);
}
// challenge_scalar(label): 64 PRF bytes reduced mod l
DAPOL_HD_INLINE void tr_challenge_scalar(merlin &t, const char *label, uint32_t llen, sc &out) {
    strobe_begin_op(t, SF_M | SF_A);
    strobe_absorb(t, reinterpret_cast<const uint8_t *>(label), llen);
    uint32_t n = 64;
    strobe_absorb_words(t, &n, 4);
    strobe_begin_op(t, SF_I | SF_A | SF_C);
    uint32_t w[16];
#pragma unroll 1
    for (int i = 0; i < 16; i++) w[i] = 0;
    for (uint32_t i = 0; i < 64; i++) {
        uint32_t b = (uint32_t)(t.st[t.pos >> 3] >> (8 * (t.pos & 7))) & 0xffu;
        t.st[t.pos >> 3] &= ~((uint64_t)0xff << (8 * (t.pos & 7)));
        w[i >> 2] |= b << (8 * (i & 3));
        t.pos++;
        if (t.pos == STROBE_R) strobe_run_f(t);
    }
    sc_from_wide(out, w);
}
// Transcript::new(&[]) + rangeproof_domain_sep(n, m)
DAPOL_HD_INLINE void tr_init_rangeproof(merlin &t, uint64_t n, uint64_t m) {
    strobe_init(t);
    tr_append_bytes(t, "dom-sep", 7, nullptr, 0);
    tr_append_bytes(t, "dom-sep", 7, reinterpret_cast<const uint8_t *>("rangeproof v1"), 13);
    tr_append_u64(t, "n", 1, n);
    tr_append_u64(t, "m", 1, m);
}
DAPOL_HD_INLINE void tr_ipp_domain_sep(merlin &t, uint64_t n) {
    tr_append_bytes(t, "dom-sep", 7, reinterpret_cast<const uint8_t *>("ipp v1"), 6);
    tr_append_u64(t, "n", 1, n);
}
DAPOL_HD_INLINE int words_are_zero(const uint32_t w[8]) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= w[i];
    return x == 0;
}

// ------------------------------------------------------------------------------------------------ generators
// BulletproofGens::new(64, m): party j, G_j[i] / H_j[i] = from_uniform_bytes(i-th 64-byte read of
// SHAKE256("GeneratorsChain" || 'G'|'H' || le32(j)))   (bulletproofs generators.rs GeneratorsChain)
// One thread squeezes the 64 x 64 bytes of one (party, tag) chain.
DAPOL_HD_INLINE void rp_gen_chain_body(uint32_t party, int is_h, uint32_t *uniform /*[64][16] words*/) {
    uint64_t st[25];
#pragma unroll 1
    for (int i = 0; i < 25; i++) st[i] = 0;
    const uint8_t lab[15] = {'G', 'e', 'n', 'e', 'r', 'a', 't', 'o', 'r', 's', 'C', 'h', 'a', 'i', 'n'};
    uint8_t in[20];
    for (int i = 0; i < 15; i++) in[i] = lab[i];
    in[15] = is_h ? 'H' : 'G';
    in[16] = (uint8_t)party; in[17] = (uint8_t)(party >> 8); in[18] = (uint8_t)(party >> 16); in[19] = (uint8_t)(party >> 24);
    for (int i = 0; i < 20; i++) st[i >> 3] ^= (uint64_t)in[i] << (8 * (i & 7));
    st[20 >> 3] ^= (uint64_t)0x1F << (8 * (20 & 7));
    st[135 >> 3] ^= (uint64_t)0x80 << (8 * (135 & 7));
    keccak_f1600(st);
    uint32_t pos = 0;  // word position inside the 136-byte (34-word) rate
#pragma unroll 1
    for (uint32_t o = 0; o < 64 * 16; o++) {
        if (pos == 34) { keccak_f1600(st); pos = 0; }
        uniform[o] = (uint32_t)(st[pos >> 1] >> (32 * (pos & 1)));
        pos++;
    }
}
DAPOL_HD_INLINE void rp_gen_point_body(uint64_t g, const uint32_t *uniform /*[g][16]*/, uint32_t *ext /*[g][32]*/) {
    uint32_t w[16];
    load8(w, uniform + 16 * g); load8(w + 8, uniform + 16 * g + 8);
    ge p;
    ge_from_uniform(p, w);
    uint32_t *o = ext + 32 * g;
    store8(o, p.X.v); store8(o + 8, p.Y.v); store8(o + 16, p.Z.v); store8(o + 24, p.T.v);
}

// ---- fixed-base window tables: tab[g][k][e] = (e+1) * 2^(W k) * P_g, affine Niels with canonical words
DAPOL_HD_INLINE void rp_load_ext(ge &p, const uint32_t *src) {
    load8(p.X.v, src); load8(p.Y.v, src + 8); load8(p.Z.v, src + 16); load8(p.T.v, src + 24);
}
DAPOL_HD_INLINE void rp_store_ext(uint32_t *dst, const ge &p) {
    store8(dst, p.X.v); store8(dst + 8, p.Y.v); store8(dst + 16, p.Z.v); store8(dst + 24, p.T.v);
}
// one thread per base point: wb[g][k] = 2^(W k) * P_g (extended)
template <int W>
DAPOL_HD_INLINE void rp_tab_windows_body(uint64_t g, const uint32_t *ext, uint32_t *wb) {
    constexpr int NW = 253 / W + 1;
    ge p;
    rp_load_ext(p, ext + 32 * g);
#pragma unroll 1
    for (int k = 0; k < NW; k++) {
        rp_store_ext(wb + 32 * (g * NW + k), p);
        if (k + 1 < NW) {
#pragma unroll 1
            for (int i = 0; i < W; i++) ge_dbl(p, p);
        }
    }
}
// one thread per (base point, window, chunk of C consecutive multiples); one shared inversion per chunk
#define RP_TAB_CHUNK 16
template <int W>
DAPOL_HD_INLINE void rp_tab_chunk_body(uint64_t item, const uint32_t *wb, ge_niels *tab) {
    constexpr int NW = 253 / W + 1;
    constexpr uint32_t HALF = 1u << (W - 1);
    constexpr uint32_t C = HALF < RP_TAB_CHUNK ? HALF : RP_TAB_CHUNK;
    constexpr uint32_t CHUNKS = HALF / C;
    uint64_t gk = item / CHUNKS;  // g * NW + k
    uint32_t e0 = (uint32_t)(item % CHUNKS) * C;
    ge base, cur;
    rp_load_ext(base, wb + 32 * gk);
    // cur = (e0 + 1) * base
    uint32_t mlt = e0 + 1;
    int top = 31;
    while (!((mlt >> top) & 1u)) top--;
    cur = base;
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        ge_dbl(cur, cur);
        if ((mlt >> b) & 1u) ge_add(cur, cur, base);
    }
    ge pts[C];
    fe pre[C];
    fe acc;
    fe_set1(acc);
#pragma unroll 1
    for (uint32_t c = 0; c < C; c++) {
        pts[c] = cur;
        pre[c] = acc;
        fe_mul(acc, acc, cur.Z);
        if (c + 1 < C) ge_add(cur, cur, base);
    }
    fe inv;
    fe_invert(inv, acc);
#pragma unroll 1
    for (int c = (int)C - 1; c >= 0; c--) {
        fe zi, x, y;
        fe_mul(zi, inv, pre[c]);
        fe_mul(inv, inv, pts[c].Z);
        fe_mul(x, pts[c].X, zi); fe_mul(y, pts[c].Y, zi);
        ge_niels nl;
        fe_add(nl.ypx, y, x); fe_sub(nl.ymx, y, x);
        fe_mul(nl.t2d, x, y); fe_mul(nl.t2d, nl.t2d, fe_const_d2());
        uint32_t w[8];
        uint32_t *o = reinterpret_cast<uint32_t *>(tab + (gk * HALF + e0 + (uint32_t)c));
        fe_canon(w, nl.ypx); store8(o, w);
        fe_canon(w, nl.ymx); store8(o + 8, w);
        fe_canon(w, nl.t2d); store8(o + 16, w);
    }
}

// acc += s * P_g using the window table of base g (s canonical); INL: mixed additions with inlined products (ge_madd_inl)
template <int W, bool INL = false>
DAPOL_HD_INLINE void rp_fixed_mul_acc(ge &acc, const ge_niels *tab, uint64_t g, const sc &s) {
    constexpr int NW = 253 / W + 1;
    int32_t d[NW];
    sc_signed_digits<W, NW>(d, s.v, 8);
    ge_comb_accumulate<W, NW, false, INL>(acc, tab + g * (uint64_t)NW * (1u << (W - 1)), d);
}
// The lone extra term of an MSM (c * B, s_bl * B_blinding, ...) spread over the CTA: thread tid adds window tid of
// s * P_g.  One more addition per thread instead of one thread doing all 253/W + 1 of them in an extra pass while
// the rest of the CTA idles (at N = 64 that extra pass was a third of the MSM's time).
template <int W>
DAPOL_HD_INLINE void rp_fixed_mul_spread(ge &acc, const ge_niels *tab, uint64_t g, const sc &s, uint32_t tid, uint32_t T) {
    constexpr int NW = 253 / W + 1;
    int32_t d[NW];
    sc_signed_digits<W, NW>(d, s.v, 8);
#pragma unroll 1
    for (uint32_t k = tid; k < (uint32_t)NW; k += T) {
        int32_t dk = d[k];
        if (dk != 0) {
            int neg = dk < 0;
            uint32_t e = (uint32_t)(neg ? -dk : dk) - 1u;
            ge_niels q;
            load_niels(q, tab + (g * (uint64_t)NW + k) * (1u << (W - 1)) + e);
            ge_madd(acc, acc, q, neg);
        }
    }
}
// acc += (+/-) P_g   (unit scalar: entry (window 0, multiple 1))
template <int W>
DAPOL_HD_INLINE void rp_fixed_unit_acc(ge &acc, const ge_niels *tab, uint64_t g, int neg) {
    constexpr int NW = 253 / W + 1;
    ge_niels q;
    load_niels(q, tab + g * (uint64_t)NW * (1u << (W - 1)));
    ge_madd(acc, acc, q, neg);
}

// variable-base s * P (s canonical): signed 4-bit windows over 8 cached multiples
DAPOL_HD_INLINE void ge_scalarmult_var(ge &r, const sc &s, const ge &p) {
    ge_cached tb[8];
    ge cur = p;
    ge_to_cached(tb[0], cur);
#pragma unroll 1
    for (int i = 1; i < 8; i++) {
        ge_cadd(cur, cur, tb[0], 0);
        ge_to_cached(tb[i], cur);
    }
    int32_t d[64];
    sc_signed_digits<4, 64>(d, s.v, 8);
    ge acc;
    ge_identity(acc);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        if (k != 63) { ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); }
        int32_t dk = d[k];
        if (dk != 0) {
            int neg = dk < 0;
            ge_cadd(acc, acc, tb[(neg ? -dk : dk) - 1], neg);
        }
    }
    r = acc;
}

// r = s1 * P1 + s2 * P2 (Straus: one doubling chain for both, signed 4-bit windows over 8 cached multiples each)
DAPOL_HD_INLINE void ge_double_scalarmult_var(ge &r, const sc &s1, const ge &p1, const sc &s2, const ge &p2) {
    ge_cached tb[2][8];
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        ge cur = j ? p2 : p1;
        ge_to_cached(tb[j][0], cur);
#pragma unroll 1
        for (int i = 1; i < 8; i++) {
            ge_cadd(cur, cur, tb[j][0], 0);
            ge_to_cached(tb[j][i], cur);
        }
    }
    int32_t d[2][64];
    sc_signed_digits<4, 64>(d[0], s1.v, 8);
    sc_signed_digits<4, 64>(d[1], s2.v, 8);
    ge acc;
    ge_identity(acc);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        if (k != 63) { ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); }
#pragma unroll 1
        for (int j = 0; j < 2; j++) {
            int32_t dk = d[j][k];
            if (dk != 0) {
                int neg = dk < 0;
                ge_cadd(acc, acc, tb[j][(neg ? -dk : dk) - 1], neg);
            }
        }
    }
    r = acc;
}

// ------------------------------------------------------------------------------------------------ batch layout
// Scalars are 8 LE words, canonical unless noted.  Per-proof challenge slots (RpBatch::chal, CH_* below).
enum {
    CH_Y = 0, CH_Z, CH_ZZ, CH_X, CH_W, CH_YINV, CH_U, CH_UINV, CH_ABL, CH_SBL, CH_T1B, CH_T2B, CH_T0, CH_T1, CH_T2, CH_CL, CH_CR,
    CH_A, CH_B, CH_ALLINV, CH_MZ, CH_SB, CH_SBBL, CH_COUNT
};
struct RpBatch {
    int nbits, m, N, lg;
    uint64_t K;
    const uint64_t *values;     // [K][m]                       (prover)
    const uint32_t *blind;      // [K][m][8] possibly unreduced (prover)
    const uint64_t *stream;     // [K] ChaCha stream id per proof
    const uint64_t *base_block; // [K] first keystream block per proof
    uint32_t seed[8];
    const uint32_t *coms;       // [K][m][8] commitments        (verifier)
    const uint32_t *proof_in;   // [K][plen/4]                  (verifier)
    uint32_t plen;              // proof bytes = 32 (9 + 2 lg)
    merlin *tr;                 // [K]
    uint32_t *Vc;               // [K][m][8]
    uint32_t *blr;              // [K][m][8] blindings reduced
    uint32_t *chal;             // [K][CH_COUNT][8]
    uint32_t *zpow;             // [K][m][8]   z^(2+j)
    uint32_t *mult;             // [K][3][32][8] per-step multipliers of the doubling expansions (y, y^-1, u^2)
    uint32_t *vecA, *vecB;      // [K][N][8]   s_L -> l -> a ;  s_R -> r -> b
    uint32_t *ypow;             // [K][N][8]   y^k, later y^-k
    uint32_t *svec;             // [K][N][8]   verifier s vector
    uint32_t *cu[2], *cui[2];   // [K][N/2][8] coefficient tables, ping-pong
    uint32_t *pts;              // [K][2][32]  extended points out of the MSM passes
    uint32_t *ptc;              // [K][2][8]   the same points compressed (rp_compress_point_body: a thread per (proof, point))
    uint32_t *parts;            // [K][2][RP_SPLIT_MAX][32] small batches: partial sums of an MSM split over several CTAs (else nullptr)
    uint32_t *gfold;            // [K][2][RP_FOLD_N][32] folded generators G^(k), H^(k) of the variable-base rounds (extended)
    uint32_t *varpts;           // [K][nvar][32] verifier: partial sums of s_q * P_q (the first vgroups entries are used)
    uint32_t *varsc;            // [K][nvar][8]  verifier: scalars of the variable points
    uint32_t *vartab;           // [K][nvar][8][32] verifier: cached multiples 1..8 of the variable points
    int vgroups;                // verifier: threads per proof of V1 (each runs Straus over every vgroups-th variable point)
    uint32_t *proof;            // [K][plen/4]  output (prover)
    int *status;                // [K] prover: 0 ok, else error; verifier: 1 accept / 0 reject
    // generator tables
    const ge_niels *tabG, *tabH;  // [m_cap*64] bases each, window W_rp
    const ge_niels *tabB, *tabBbl;  // 1 base each, window W_rp
};
DAPOL_HD_INLINE uint32_t *rp_ch(const RpBatch &b, uint64_t p, int slot) { return b.chal + (p * CH_COUNT + slot) * 8; }
DAPOL_HD_INLINE void rp_ld(sc &s, const uint32_t *src) { load8(s.v, src); }
DAPOL_HD_INLINE void rp_st(uint32_t *dst, const sc &s) { store8(dst, s.v); }
DAPOL_HD_INLINE void sc_set1(sc &s) { sc_set_u64(s, 1); }
DAPOL_HD_INLINE void rp_rng_scalar(sc &out, const RpBatch &b, uint64_t p, uint64_t draw) {
    uint32_t ks[16];
    chacha20_block(ks, b.seed, b.base_block[p] + draw, b.stream[p]);
    sc_from_wide(out, ks);
}
// generator table index of position k of the concatenated vectors: party j = k / n, bit i = k % n
DAPOL_HD_INLINE uint64_t rp_gen_of(const RpBatch &b, uint32_t k) { return (uint64_t)(k / (uint32_t)b.nbits) * 64u + (k % (uint32_t)b.nbits); }

// ------------------------------------------------------------------------------------------------ prover passes
// P0 (thread per proof): blinding draws and their sums, reduced value blindings, transcript start.
// Draw order of bulletproofs' party/dealer (App. A.4): party j: a_blinding, s_blinding, s_L[0..n), s_R[0..n); then per party
// t1_blinding, t2_blinding.
DAPOL_HD_INLINE void rp_p0_body(const RpBatch &b, uint64_t p) {
    const uint64_t per = 2 + 2 * (uint64_t)b.nbits;
    sc abl, sbl, t1b, t2b, t;
    sc_set_u64(abl, 0); sbl = abl; t1b = abl; t2b = abl;
#pragma unroll 1
    for (int j = 0; j < b.m; j++) {
        rp_rng_scalar(t, b, p, j * per); sc_add(abl, abl, t);
        rp_rng_scalar(t, b, p, j * per + 1); sc_add(sbl, sbl, t);
        rp_rng_scalar(t, b, p, b.m * per + 2 * j); sc_add(t1b, t1b, t);
        rp_rng_scalar(t, b, p, b.m * per + 2 * j + 1); sc_add(t2b, t2b, t);
        sc r;
        rp_ld(r, b.blind + (p * b.m + j) * 8);
        sc_reduce256(r, r);
        rp_st(b.blr + (p * b.m + j) * 8, r);
    }
    rp_st(rp_ch(b, p, CH_ABL), abl); rp_st(rp_ch(b, p, CH_SBL), sbl);
    rp_st(rp_ch(b, p, CH_T1B), t1b); rp_st(rp_ch(b, p, CH_T2B), t2b);
    merlin tr;
    tr_init_rangeproof(tr, (uint64_t)b.nbits, (uint64_t)b.m);
    b.tr[p] = tr;
    b.status[p] = 0;
}
// P1 (thread per (proof, party)): V_j = commit(v_j, r_j) compressed, with the tree's comb tables (window WT):
// tab_b holds multiples of B/2 and the blinding is halved mod l, so the sum is the half point of V_j and
// compress(V_j) is the batched double-and-compress with a batch of one
template <int WT>
DAPOL_HD_INLINE void rp_p1_body(const RpBatch &b, uint64_t p, int j, const ge_niels *tab_b, const ge_niels *tab_bbl) {
    constexpr int WV = comb_value_window<WT>::value;
    constexpr int NWR = 253 / WT + 1, NWV = 64 / WV + 1;
    sc r, rh;
    rp_ld(r, b.blind + (p * b.m + j) * 8);
    sc_half256(rh, r);
    uint64_t v = b.values[p * b.m + j];
    uint32_t vw[2] = {(uint32_t)v, (uint32_t)(v >> 32)};
    int32_t dr[NWR], dv[NWV];
    sc_signed_digits<WT, NWR>(dr, rh.v, 8);
    sc_signed_digits<WV, NWV>(dv, vw, 2);
    ge acc;
    ge_identity(acc);
    ge_comb_accumulate<WV, NWV>(acc, tab_b, dv);
    ge_comb_accumulate<WT, NWR>(acc, tab_bbl, dr);
    ge_dc_batch<1> dc;
    dc.init();
    dc.push(acc);
    dc.solve();
    uint32_t cc[8];
    dc.get(0, cc);
    store8(b.Vc + (p * b.m + j) * 8, cc);
}
// P2 (thread per (proof, k)): s_L[k], s_R[k]
DAPOL_HD_INLINE void rp_p2_body(const RpBatch &b, uint64_t p, uint32_t k) {
    const uint64_t n = (uint64_t)b.nbits, per = 2 + 2 * n;
    uint64_t j = k / n, i = k % n;
    sc s;
    rp_rng_scalar(s, b, p, j * per + 2 + i);
    rp_st(b.vecA + (p * b.N + k) * 8, s);
    rp_rng_scalar(s, b, p, j * per + 2 + n + i);
    rp_st(b.vecB + (p * b.N + k) * 8, s);
}
// P3 (CTA per (proof, which)): which = 0: A = a_bl * B_bl + sum_k (bit_k ? G_k : -H_k)
//                               which = 1: S = s_bl * B_bl + <s_L, G> + <s_R, H>          -- per-thread partial sums
template <int W, bool INL = false>
DAPOL_HD_INLINE void rp_p3_partial(ge &acc, const RpBatch &b, uint64_t p, int which, uint32_t tid, uint32_t T) {
    ge_identity(acc);
    const uint32_t N = (uint32_t)b.N, n = (uint32_t)b.nbits;
    if (which == 0) {
#pragma unroll 1
        for (uint32_t k = tid; k < N; k += T) {
            uint64_t v = b.values[p * b.m + k / n];
            int bit = (int)((v >> (k % n)) & 1);
            rp_fixed_unit_acc<W>(acc, bit ? b.tabG : b.tabH, rp_gen_of(b, k), !bit);
        }
        sc s;
        rp_ld(s, rp_ch(b, p, CH_ABL));
        rp_fixed_mul_spread<W>(acc, b.tabBbl, 0, s, tid, T);
    } else {
#pragma unroll 1
        for (uint32_t t = tid; t < 2 * N; t += T) {
            sc s;
            if (t < N) { rp_ld(s, b.vecA + (p * N + t) * 8); rp_fixed_mul_acc<W, INL>(acc, b.tabG, rp_gen_of(b, t), s); }
            else { rp_ld(s, b.vecB + (p * N + (t - N)) * 8); rp_fixed_mul_acc<W, INL>(acc, b.tabH, rp_gen_of(b, t - N), s); }
        }
        sc s;
        rp_ld(s, rp_ch(b, p, CH_SBL));
        rp_fixed_mul_spread<W>(acc, b.tabBbl, 0, s, tid, T);
    }
}
// Small batches (a single proof is the reference's own benchmark, benches/dapol.rs:59-141): one CTA per (proof, L | R) leaves
// most of the GPU idle, so an MSM is split over S CTAs -- the strided term loops take (tid, T) over all S * blockDim threads --
// and the S partial sums are added by rp_sum_parts_body (thread per (proof, L | R)).
#define RP_SPLIT_MAX 32
DAPOL_HD_INLINE void rp_sum_parts_body(const RpBatch &b, uint64_t pw, uint32_t S) {
    ge acc, q;
    rp_load_ext(acc, b.parts + (pw * RP_SPLIT_MAX) * 32);
#pragma unroll 1
    for (uint32_t z = 1; z < S; z++) { rp_load_ext(q, b.parts + (pw * RP_SPLIT_MAX + z) * 32); ge_add(acc, acc, q); }
    rp_store_ext(b.pts + pw * 32, acc);
}
// write the sum point of an MSM pass
DAPOL_HD_INLINE void rp_store_point(const RpBatch &b, uint64_t p, int which, const ge &pt) { rp_store_ext(b.pts + (p * 2 + which) * 32, pt); }

// (thread per (proof, which)) compress the point of an MSM pass: the two points of a proof side by side instead of one after
// the other inside the thread-per-proof transcript passes (an inverse square root each: a third of a single proof's latency)
DAPOL_HD_INLINE void rp_compress_point_body(const RpBatch &b, uint64_t pw) {
    ge pt;
    uint32_t cc[8];
    rp_load_ext(pt, b.pts + pw * 32);
    ge_compress(cc, pt);
    store8(b.ptc + pw * 8, cc);
}
// P4 (thread per proof): A, S compressed -> proof; transcript V.., A, S -> y, z; z powers; doubling multipliers of y
DAPOL_HD_INLINE void rp_p4_body(const RpBatch &b, uint64_t p) {
    merlin tr = b.tr[p];
    for (int j = 0; j < b.m; j++) tr_append_words(tr, "V", 1, b.Vc + (p * b.m + j) * 8, 32);
    uint32_t cc[8];
    uint32_t *out = b.proof + p * (b.plen / 4);
    load8(cc, b.ptc + (p * 2 + 0) * 8); store8(out, cc); tr_append_words(tr, "A", 1, cc, 32);
    load8(cc, b.ptc + (p * 2 + 1) * 8); store8(out + 8, cc); tr_append_words(tr, "S", 1, cc, 32);
    sc y, z, zz, t;
    tr_challenge_scalar(tr, "y", 1, y);
    tr_challenge_scalar(tr, "z", 1, z);
    sc_mul(zz, z, z);
    rp_st(rp_ch(b, p, CH_Y), y); rp_st(rp_ch(b, p, CH_Z), z); rp_st(rp_ch(b, p, CH_ZZ), zz);
    t = zz;
#pragma unroll 1
    for (int j = 0; j < b.m; j++) { rp_st(b.zpow + (p * b.m + j) * 8, t); sc_mul(t, t, z); }
    t = y;
#pragma unroll 1
    for (int s = 0; s < b.lg; s++) { rp_st(b.mult + ((p * 3 + 0) * 32 + s) * 8, t); sc_mul(t, t, t); }
    sc_set1(t);
    rp_st(b.ypow + p * b.N * 8, t);  // y^0; the doubling expansion fills the rest
    b.tr[p] = tr;
}
// doubling expansion step s (threads i < 2^s of a proof): vec[i + 2^s] = vec[i] * mult[s]; vec[0] preset
DAPOL_HD_INLINE void rp_expand_step(uint32_t *vec /*proof base*/, const uint32_t *mult /*proof, slot base*/, int s, uint32_t i) {
    sc a, m, r;
    rp_ld(a, vec + (uint64_t)i * 8);
    rp_ld(m, mult + s * 8);
    sc_mul(r, a, m);
    rp_st(vec + ((uint64_t)i + (1u << s)) * 8, r);
}
// z^(2+j) 2^i of position k
DAPOL_HD_INLINE void rp_zz2(sc &out, const RpBatch &b, uint64_t p, uint32_t k) {
    uint32_t j = k / (uint32_t)b.nbits, i = k % (uint32_t)b.nbits;
    sc zj, two;
    rp_ld(zj, b.zpow + (p * b.m + j) * 8);
    sc_set_u64(two, 1ull << i);
    sc_mul(out, zj, two);
}
// P5 (CTA per proof, partial sums): t0 = <l0, r0>, t1 = <l0, r1> + <l1, r0>, t2 = <l1, r1> with
// l0 = a_L - z, l1 = s_L, r0 = y^k (a_R + z) + z^(2+j) 2^i, r1 = y^k s_R
DAPOL_HD_INLINE void rp_p5_partial(sc &t0, sc &t1, sc &t2, const RpBatch &b, uint64_t p, uint32_t tid, uint32_t T) {
    sc_set_u64(t0, 0); t1 = t0; t2 = t0;
    sc z, one;
    rp_ld(z, rp_ch(b, p, CH_Z));
    sc_set1(one);
    const uint32_t N = (uint32_t)b.N, n = (uint32_t)b.nbits;
#pragma unroll 1
    for (uint32_t k = tid; k < N; k += T) {
        uint64_t v = b.values[p * b.m + k / n];
        sc aL, l0, r0, r1, yk, sL, sR, t, u;
        sc_set_u64(aL, (v >> (k % n)) & 1);
        sc_sub(l0, aL, z);
        sc_sub(t, aL, one); sc_add(t, t, z);  // a_R + z
        rp_ld(yk, b.ypow + (p * N + k) * 8);
        sc_mul(r0, yk, t);
        rp_zz2(t, b, p, k);
        sc_add(r0, r0, t);
        rp_ld(sL, b.vecA + (p * N + k) * 8); rp_ld(sR, b.vecB + (p * N + k) * 8);
        sc_mul(r1, yk, sR);
        sc_mul(t, l0, r0); sc_add(t0, t0, t);
        sc_mul(t, sL, r1); sc_add(t2, t2, t);
        sc_mul(t, l0, r1); sc_mul(u, sL, r0); sc_add(t, t, u); sc_add(t1, t1, t);
    }
}
// P6 (thread per (proof, which)): T_1 = commit(t1, t1_blinding), T_2 = commit(t2, t2_blinding) (sums over parties)
template <int W>
DAPOL_HD_INLINE void rp_p6_body(const RpBatch &b, uint64_t p, int which) {
    sc t, tb;
    rp_ld(t, rp_ch(b, p, which ? CH_T2 : CH_T1));
    rp_ld(tb, rp_ch(b, p, which ? CH_T2B : CH_T1B));
    ge acc;
    ge_identity(acc);
    rp_fixed_mul_acc<W>(acc, b.tabB, 0, t);
    rp_fixed_mul_acc<W>(acc, b.tabBbl, 0, tb);
    rp_store_point(b, p, which, acc);
}
// P7 (thread per proof): T_1, T_2 -> x; t_x, t_x_blinding, e_blinding -> w; ipp domain separator; y^-1 and its doubling
// multipliers; coefficient tables start at 1
DAPOL_HD_INLINE void rp_p7_body(const RpBatch &b, uint64_t p) {
    merlin tr = b.tr[p];
    uint32_t cc[8];
    uint32_t *out = b.proof + p * (b.plen / 4);
    load8(cc, b.ptc + (p * 2 + 0) * 8); store8(out + 16, cc); tr_append_words(tr, "T_1", 3, cc, 32);
    load8(cc, b.ptc + (p * 2 + 1) * 8); store8(out + 24, cc); tr_append_words(tr, "T_2", 3, cc, 32);
    sc x, xx, w, t, u, tx, txb, eb;
    tr_challenge_scalar(tr, "x", 1, x);
    if (sc_iszero(x)) b.status[p] = DAPOL_RP_ERR_CHALLENGE;
    sc_mul(xx, x, x);
    sc t0, t1, t2;
    rp_ld(t0, rp_ch(b, p, CH_T0)); rp_ld(t1, rp_ch(b, p, CH_T1)); rp_ld(t2, rp_ch(b, p, CH_T2));
    sc_mul(t, t1, x); sc_add(tx, t0, t); sc_mul(t, t2, xx); sc_add(tx, tx, t);
    sc_set_u64(txb, 0);
#pragma unroll 1
    for (int j = 0; j < b.m; j++) {
        sc zj, r;
        rp_ld(zj, b.zpow + (p * b.m + j) * 8); rp_ld(r, b.blr + (p * b.m + j) * 8);
        sc_mul(t, zj, r); sc_add(txb, txb, t);
    }
    rp_ld(t, rp_ch(b, p, CH_T1B)); sc_mul(t, t, x); sc_add(txb, txb, t);
    rp_ld(t, rp_ch(b, p, CH_T2B)); sc_mul(t, t, xx); sc_add(txb, txb, t);
    rp_ld(u, rp_ch(b, p, CH_SBL)); sc_mul(u, u, x); rp_ld(t, rp_ch(b, p, CH_ABL)); sc_add(eb, t, u);
    store8(out + 32, tx.v); store8(out + 40, txb.v); store8(out + 48, eb.v);
    tr_append_words(tr, "t_x", 3, tx.v, 32);
    tr_append_words(tr, "t_x_blinding", 12, txb.v, 32);
    tr_append_words(tr, "e_blinding", 10, eb.v, 32);
    tr_challenge_scalar(tr, "w", 1, w);
    tr_ipp_domain_sep(tr, (uint64_t)b.N);
    rp_st(rp_ch(b, p, CH_X), x); rp_st(rp_ch(b, p, CH_W), w);
    sc y, yi;
    rp_ld(y, rp_ch(b, p, CH_Y));
    sc_invert(yi, y);
    rp_st(rp_ch(b, p, CH_YINV), yi);
    t = yi;
#pragma unroll 1
    for (int s = 0; s < b.lg; s++) { rp_st(b.mult + ((p * 3 + 1) * 32 + s) * 8, t); sc_mul(t, t, t); }
    sc one;
    sc_set1(one);
    if (b.N > 1) { rp_st(b.cu[0] + p * (b.N / 2) * 8, one); rp_st(b.cui[0] + p * (b.N / 2) * 8, one); }
    b.tr[p] = tr;
}
// P8 (thread per (proof, k)): l = (a_L - z) + s_L x -> vecA ; r = y^k (a_R + z + s_R x) + z^(2+j) 2^i -> vecB
DAPOL_HD_INLINE void rp_p8_body(const RpBatch &b, uint64_t p, uint32_t k) {
    const uint32_t N = (uint32_t)b.N, n = (uint32_t)b.nbits;
    sc z, x, one, aL, sL, sR, yk, l, r, t;
    rp_ld(z, rp_ch(b, p, CH_Z)); rp_ld(x, rp_ch(b, p, CH_X));
    sc_set1(one);
    uint64_t v = b.values[p * b.m + k / n];
    sc_set_u64(aL, (v >> (k % n)) & 1);
    rp_ld(sL, b.vecA + (p * N + k) * 8); rp_ld(sR, b.vecB + (p * N + k) * 8);
    rp_ld(yk, b.ypow + (p * N + k) * 8);
    sc_mul(t, sL, x); sc_sub(l, aL, z); sc_add(l, l, t);
    sc_mul(t, sR, x); sc_sub(r, aL, one); sc_add(r, r, z); sc_add(r, r, t);
    sc_mul(r, yk, r);
    rp_zz2(t, b, p, k);
    sc_add(r, r, t);
    rp_st(b.vecA + (p * N + k) * 8, l); rp_st(b.vecB + (p * N + k) * 8, r);
}
// IPA round `rnd` (1-based), h = N >> rnd.
// P9 (CTA per proof, partial sums): c_L = <a_L, b_R>, c_R = <a_R, b_L>
DAPOL_HD_INLINE void rp_p9_partial(sc &cl, sc &cr, const RpBatch &b, uint64_t p, int rnd, uint32_t tid, uint32_t T) {
    const uint32_t N = (uint32_t)b.N, h = N >> rnd;
    sc_set_u64(cl, 0); cr = cl;
#pragma unroll 1
    for (uint32_t i = tid; i < h; i += T) {
        sc al, ar, bl, br, t;
        rp_ld(al, b.vecA + (p * N + i) * 8); rp_ld(ar, b.vecA + (p * N + h + i) * 8);
        rp_ld(bl, b.vecB + (p * N + i) * 8); rp_ld(br, b.vecB + (p * N + h + i) * 8);
        sc_mul(t, al, br); sc_add(cl, cl, t);
        sc_mul(t, ar, bl); sc_add(cr, cr, t);
    }
}
// P10 (CTA per (proof, which)): L (which = 0) / R (which = 1) of the round over the ORIGINAL generators.
// With I = original position, pfx = its top rnd-1 bits, bit = its rnd-th bit, i = its low lg-rnd bits:
//   folded G_i = sum cu[pfx] G_I,  cu[pfx] = prod_{r<rnd} u_r^(2 b_r - 1);  folded H_i = sum y^-I cui[pfx] H_I, cui = 1/cu
//   L = <a_lo, G_hi> + <b_hi, H_lo> + c_L w B      R = <a_hi, G_lo> + <b_lo, H_hi> + c_R w B     (Q = w B)
template <int W, bool INL = false>
DAPOL_HD_INLINE void rp_p10_partial(ge &acc, const RpBatch &b, uint64_t p, int rnd, int which, uint32_t tid, uint32_t T) {
    ge_identity(acc);
    const uint32_t N = (uint32_t)b.N, h = N >> rnd, half = N / 2;
    const int cur = (rnd - 1) & 1;
    const uint32_t *cu = b.cu[cur] + p * half * 8, *cui = b.cui[cur] + p * half * 8;
    const int sh = b.lg - rnd;  // log2 h
#pragma unroll 1
    for (uint32_t t = tid; t < N; t += T) {
        sc s, c;
        if (t < half) {         // G terms: positions with bit = 1 - which ... L uses G_hi (bit 1), R uses G_lo (bit 0)
            uint32_t pfx = t >> sh, i = t & (h - 1);
            uint32_t bit = which ? 0u : 1u;
            uint32_t I = (pfx << (sh + 1)) | (bit << sh) | i;
            rp_ld(s, b.vecA + (p * N + (which ? h + i : i)) * 8);
            rp_ld(c, cu + pfx * 8);
            sc_mul(s, s, c);
            rp_fixed_mul_acc<W, INL>(acc, b.tabG, rp_gen_of(b, I), s);
        } else {                // H terms: L uses H_lo (bit 0) with b_hi, R uses H_hi (bit 1) with b_lo
            uint32_t tt = t - half;
            uint32_t pfx = tt >> sh, i = tt & (h - 1);
            uint32_t bit = which ? 1u : 0u;
            uint32_t I = (pfx << (sh + 1)) | (bit << sh) | i;
            rp_ld(s, b.vecB + (p * N + (which ? i : h + i)) * 8);
            rp_ld(c, cui + pfx * 8);
            sc_mul(s, s, c);
            rp_ld(c, b.ypow + (p * N + I) * 8);
            sc_mul(s, s, c);
            rp_fixed_mul_acc<W, INL>(acc, b.tabH, rp_gen_of(b, I), s);
        }
    }
    sc s, c;  // + c_L/R * w * B, one window per thread
    rp_ld(s, rp_ch(b, p, which ? CH_CR : CH_CL));
    rp_ld(c, rp_ch(b, p, CH_W));
    sc_mul(s, s, c);
    rp_fixed_mul_spread<W>(acc, b.tabB, 0, s, tid, T);
}
// P11 (thread per proof): L, R -> u, u^-1
DAPOL_HD_INLINE void rp_p11_body(const RpBatch &b, uint64_t p, int rnd) {
    merlin tr = b.tr[p];
    uint32_t cc[8];
    uint32_t *out = b.proof + p * (b.plen / 4) + 56 + 16 * (rnd - 1);
    load8(cc, b.ptc + (p * 2 + 0) * 8); store8(out, cc); tr_append_words(tr, "L", 1, cc, 32);
    load8(cc, b.ptc + (p * 2 + 1) * 8); store8(out + 8, cc); tr_append_words(tr, "R", 1, cc, 32);
    sc u, ui;
    tr_challenge_scalar(tr, "u", 1, u);
    sc_invert(ui, u);
    rp_st(rp_ch(b, p, CH_U), u); rp_st(rp_ch(b, p, CH_UINV), ui);
    b.tr[p] = tr;
}
// P12 (thread per (proof, i < max(h, 2^(rnd-1)))): fold a, b; grow the coefficient tables; last round writes a, b
DAPOL_HD_INLINE void rp_p12_body(const RpBatch &b, uint64_t p, int rnd, uint32_t i) {
    const uint32_t N = (uint32_t)b.N, h = N >> rnd, half = N / 2;
    sc u, ui;
    rp_ld(u, rp_ch(b, p, CH_U)); rp_ld(ui, rp_ch(b, p, CH_UINV));
    if (i < h) {
        sc lo, hi, t, r;
        rp_ld(lo, b.vecA + (p * N + i) * 8); rp_ld(hi, b.vecA + (p * N + h + i) * 8);
        sc_mul(t, lo, u); sc_mul(r, hi, ui); sc_add(r, r, t);
        rp_st(b.vecA + (p * N + i) * 8, r);
        if (h == 1) store8(b.proof + p * (b.plen / 4) + b.plen / 4 - 16, r.v);
        rp_ld(lo, b.vecB + (p * N + i) * 8); rp_ld(hi, b.vecB + (p * N + h + i) * 8);
        sc_mul(t, lo, ui); sc_mul(r, hi, u); sc_add(r, r, t);
        rp_st(b.vecB + (p * N + i) * 8, r);
        if (h == 1) store8(b.proof + p * (b.plen / 4) + b.plen / 4 - 8, r.v);
    }
    if (rnd < b.lg && i < (1u << (rnd - 1))) {
        const int cur = (rnd - 1) & 1, nxt = rnd & 1;
        sc c, r;
        rp_ld(c, b.cu[cur] + (p * half + i) * 8);
        sc_mul(r, c, ui); rp_st(b.cu[nxt] + (p * half + 2 * i) * 8, r);
        sc_mul(r, c, u); rp_st(b.cu[nxt] + (p * half + 2 * i + 1) * 8, r);
        rp_ld(c, b.cui[cur] + (p * half + i) * 8);
        sc_mul(r, c, u); rp_st(b.cui[nxt] + (p * half + 2 * i) * 8, r);
        sc_mul(r, c, ui); rp_st(b.cui[nxt] + (p * half + 2 * i + 1) * 8, r);
    }
}

// ---- hybrid inner-product argument for the large aggregates (N >= RP_HYBRID_MIN_N) --------------------------------------
// A round over the ORIGINAL generators costs 2N fixed-base multiplications whatever the round; bulletproofs' own rounds
// over the FOLDED generators cost 4 * N / 2^(k-1) variable-base ones (L, R and the folding of G, H).  A variable-base
// multiplication is ~19 fixed-base ones, so the late rounds are cheaper folded: the rounds switch when the vectors are
// RP_FOLD_N = 32 long (round lg - 4).  PM materialises the folded generators of that round -- still fixed-base:
// G^(s)_j = sum_pfx cu[pfx] G_(pfx, j), H^(s)_j = sum_pfx y^-I cui[pfx] H_(pfx, j) -- and the remaining rounds are the
// textbook ones (PV: L, R by variable-base multiplications; PF: G' = u^-1 G_lo + u G_hi, H' = u H_lo + u^-1 H_hi).  The
// group elements L_k, R_k are the same, so the proof bytes are too.  At N = 2048 (m = 32): 16 N fixed-base + 184 variable-
// base multiplications per proof instead of 22 N fixed-base ones.
#define RP_FOLD_N 32
#ifndef RP_HYBRID_MIN_N
#define RP_HYBRID_MIN_N 1024
#endif
// first variable-base round (1-based), or lg + 1 when every round runs over the original generators
DAPOL_HD_INLINE int rp_switch_round(int N, int lg) { return N >= RP_HYBRID_MIN_N ? lg - 4 : lg + 1; }
DAPOL_HD_INLINE uint32_t *rp_gfold(const RpBatch &b, uint64_t p, int which, uint32_t j) { return b.gfold + ((p * 2 + which) * RP_FOLD_N + j) * 32; }
// PM (thread per (proof, G|H, j < RP_FOLD_N)) at the switch round s
template <int W, bool INL = false>
DAPOL_HD_INLINE void rp_pm_body(const RpBatch &b, uint64_t p, int which, uint32_t j, int s) {
    const uint32_t N = (uint32_t)b.N, half = N / 2, npfx = 1u << (s - 1);
    const int cur = (s - 1) & 1, shj = b.lg - s + 1;  // log2 of the folded length
    const uint32_t *co = (which ? b.cui[cur] : b.cu[cur]) + p * half * 8;
    ge acc;
    ge_identity(acc);
#pragma unroll 1
    for (uint32_t pfx = 0; pfx < npfx; pfx++) {
        uint32_t I = (pfx << shj) | j;
        sc c;
        rp_ld(c, co + pfx * 8);
        if (which) {
            sc y;
            rp_ld(y, b.ypow + (p * N + I) * 8);
            sc_mul(c, c, y);
        }
        rp_fixed_mul_acc<W, INL>(acc, which ? b.tabH : b.tabG, rp_gen_of(b, I), c);
    }
    rp_store_ext(rp_gfold(b, p, which, j), acc);
}
// PV (thread per (proof, L|R, t < 2h), h = N >> rnd): term t of
//   L = <a_lo, G_hi> + <b_hi, H_lo> + c_L w B      R = <a_hi, G_lo> + <b_lo, H_hi> + c_R w B
// plus this thread's windows of the B term; the 2h partial points of a (proof, L|R) are summed by the caller
template <int W>
DAPOL_HD_INLINE void rp_pv_partial(ge &acc, const RpBatch &b, uint64_t p, int rnd, int which, uint32_t t, uint32_t g) {
    const uint32_t N = (uint32_t)b.N, h = N >> rnd;
    sc s, c;
    ge pt;
    if (t < h) {
        rp_ld(s, b.vecA + (p * N + (which ? h + t : t)) * 8);
        rp_load_ext(pt, rp_gfold(b, p, 0, which ? t : h + t));
    } else {
        uint32_t tt = t - h;
        rp_ld(s, b.vecB + (p * N + (which ? tt : h + tt)) * 8);
        rp_load_ext(pt, rp_gfold(b, p, 1, which ? h + tt : tt));
    }
    ge_scalarmult_var(acc, s, pt);
    rp_ld(s, rp_ch(b, p, which ? CH_CR : CH_CL));
    rp_ld(c, rp_ch(b, p, CH_W));
    sc_mul(s, s, c);
    rp_fixed_mul_spread<W>(acc, b.tabB, 0, s, t, g);
}
// PF (thread per (proof, G|H, i < h)) after the round's challenge: fold the generators in place
DAPOL_HD_INLINE void rp_pf_body(const RpBatch &b, uint64_t p, int rnd, int which, uint32_t i) {
    const uint32_t h = (uint32_t)b.N >> rnd;
    sc u, ui;
    rp_ld(u, rp_ch(b, p, CH_U)); rp_ld(ui, rp_ch(b, p, CH_UINV));
    ge lo, hi, r;
    rp_load_ext(lo, rp_gfold(b, p, which, i)); rp_load_ext(hi, rp_gfold(b, p, which, h + i));
    if (which) ge_double_scalarmult_var(r, u, lo, ui, hi);
    else ge_double_scalarmult_var(r, ui, lo, u, hi);
    rp_store_ext(rp_gfold(b, p, which, i), r);
}

// ------------------------------------------------------------------------------------------------ verifier passes
// number of variable-base points of a proof: A, S, T_1, T_2, L_k, R_k, V_j
DAPOL_HD_INLINE int rp_nvar(int lg, int m) { return 4 + 2 * lg + m; }
// V0 (thread per proof): format checks (RangeProof::from_bytes), transcript replay, challenges, one shared inversion for
// y and the u_k, scalars of the variable points and of B / B_blinding, multipliers of the s and y^-1 expansions.
DAPOL_HD_INLINE void rp_v0_body(const RpBatch &b, uint64_t p) {
    const uint32_t *pr = b.proof_in + p * (b.plen / 4);
    const int lg = b.lg, m = b.m, nv = rp_nvar(lg, m);
    int ok = 1;
    sc tx, txb, eb, a, bb;
    const uint32_t *ab = pr + b.plen / 4 - 16;
    ok &= sc_is_canonical(pr + 32) & sc_is_canonical(pr + 40) & sc_is_canonical(pr + 48) & sc_is_canonical(ab) & sc_is_canonical(ab + 8);
    rp_ld(tx, pr + 32); rp_ld(txb, pr + 40); rp_ld(eb, pr + 48); rp_ld(a, ab); rp_ld(bb, ab + 8);
    merlin tr;
    tr_init_rangeproof(tr, (uint64_t)b.nbits, (uint64_t)m);
    for (int j = 0; j < m; j++) tr_append_words(tr, "V", 1, b.coms + (p * m + j) * 8, 32);
    // validate_and_append_point: the identity encoding is rejected
    ok &= !words_are_zero(pr) & !words_are_zero(pr + 8) & !words_are_zero(pr + 16) & !words_are_zero(pr + 24);
    tr_append_words(tr, "A", 1, pr, 32);
    tr_append_words(tr, "S", 1, pr + 8, 32);
    sc y, z, zz, x, w, c, one;
    sc_set1(one);
    tr_challenge_scalar(tr, "y", 1, y);
    tr_challenge_scalar(tr, "z", 1, z);
    sc_mul(zz, z, z);
    tr_append_words(tr, "T_1", 3, pr + 16, 32);
    tr_append_words(tr, "T_2", 3, pr + 24, 32);
    tr_challenge_scalar(tr, "x", 1, x);
    tr_append_words(tr, "t_x", 3, pr + 32, 32);
    tr_append_words(tr, "t_x_blinding", 12, pr + 40, 32);
    tr_append_words(tr, "e_blinding", 10, pr + 48, 32);
    tr_challenge_scalar(tr, "w", 1, w);
    tr_ipp_domain_sep(tr, (uint64_t)b.N);
    uint32_t *vs = b.varsc + p * nv * 8;
    // challenges u_k parked in the L_k scalar slots for now; prefix products in the R_k slots
    sc prod = y, t;
    for (int k = 0; k < lg; k++) {
        const uint32_t *L = pr + 56 + 16 * k;
        ok &= !words_are_zero(L) & !words_are_zero(L + 8);
        tr_append_words(tr, "L", 1, L, 32);
        tr_append_words(tr, "R", 1, L + 8, 32);
        sc u;
        tr_challenge_scalar(tr, "u", 1, u);
        rp_st(vs + (4 + 2 * k) * 8, u);
        rp_st(vs + (5 + 2 * k) * 8, prod);  // y u_0 .. u_{k-1}
        sc_mul(prod, prod, u);
    }
    // batching weight c: any value (thread_rng in the reference, bulletproofs verify_multiple_with_rng); here from the transcript
    tr_challenge_scalar(tr, "dapol-b200 batching weight", 26, c);
    int zero_ch = sc_iszero(prod);  // a zero challenge has no inverse (probability ~2^-252); reject
    ok &= !zero_ch;
    sc inv, allinv, yinv;
    if (zero_ch) sc_set1(prod);
    sc_invert(inv, prod);
    sc_set1(allinv);
    for (int k = lg - 1; k >= 0; k--) {
        sc u, pre, ui, usq, uisq;
        rp_ld(u, vs + (4 + 2 * k) * 8); rp_ld(pre, vs + (5 + 2 * k) * 8);
        sc_mul(ui, inv, pre);
        sc_mul(inv, inv, u);
        sc_mul(usq, u, u); sc_mul(uisq, ui, ui);
        sc_mul(allinv, allinv, ui);
        rp_st(vs + (4 + 2 * k) * 8, usq); rp_st(vs + (5 + 2 * k) * 8, uisq);
        rp_st(b.mult + ((p * 3 + 2) * 32 + (lg - 1 - k)) * 8, usq);  // s[i + 2^s] = s[i] * u^2_{lg-1-s}
    }
    yinv = inv;
    t = yinv;
    for (int s = 0; s < lg; s++) { rp_st(b.mult + ((p * 3 + 1) * 32 + s) * 8, t); sc_mul(t, t, t); }
    rp_st(b.svec + p * b.N * 8, allinv);
    rp_st(b.ypow + p * b.N * 8, one);
    // variable-point scalars: A: 1, S: x, T_1: c x, T_2: c x^2, V_j: c z^(2+j)
    sc cx, cxx;
    sc_mul(cx, c, x); sc_mul(cxx, cx, x);
    rp_st(vs, one); rp_st(vs + 8, x); rp_st(vs + 16, cx); rp_st(vs + 24, cxx);
    t = zz;
    for (int j = 0; j < m; j++) {
        sc cz;
        rp_st(b.zpow + (p * m + j) * 8, t);
        sc_mul(cz, c, t);
        rp_st(vs + (4 + 2 * lg + j) * 8, cz);
        sc_mul(t, t, z);
    }
    // delta = (z - z^2) sum_{i<N} y^i - z^3 (2^n - 1) sum_{j<m} z^j ; the sums as products of (1 + q^(2^s)) (N, m powers of two)
    sc sumy, sumz, q, delta, two_n;
    sc_set1(sumy); q = y;
    for (int s = 0; s < lg; s++) { sc_add(t, one, q); sc_mul(sumy, sumy, t); sc_mul(q, q, q); }
    sc_set1(sumz); q = z;
    for (int s = 1; s < m; s <<= 1) { sc_add(t, one, q); sc_mul(sumz, sumz, t); sc_mul(q, q, q); }
    sc_set_u64(two_n, b.nbits == 64 ? ~0ull : ((1ull << b.nbits) - 1));
    sc_sub(t, z, zz); sc_mul(delta, t, sumy);
    sc_mul(t, zz, z); sc_mul(t, t, two_n); sc_mul(t, t, sumz); sc_sub(delta, delta, t);
    // B_blinding: -e_blinding - c t_x_blinding ;  B: w (t_x - a b) + c (delta - t_x)
    sc sB, sBbl, u2;
    sc_mul(t, c, txb); sc_add(t, t, eb); sc_neg(sBbl, t);
    sc_mul(t, a, bb); sc_sub(t, tx, t); sc_mul(t, w, t);
    sc_sub(u2, delta, tx); sc_mul(u2, c, u2); sc_add(sB, t, u2);
    rp_st(rp_ch(b, p, CH_SB), sB); rp_st(rp_ch(b, p, CH_SBBL), sBbl);
    sc mz;
    sc_neg(mz, z);
    rp_st(rp_ch(b, p, CH_Z), z); rp_st(rp_ch(b, p, CH_MZ), mz); rp_st(rp_ch(b, p, CH_A), a); rp_st(rp_ch(b, p, CH_B), bb);
    b.status[p] = ok;
}
// V1 (vgroups threads per proof): sum of s_q * P_q over the variable points q = grp, grp + vgroups, ... of the proof, by
// Straus' interleaving -- ONE chain of 252 doublings shared by the thread's points (signed 4-bit windows over 8 cached
// multiples per point, kept in HBM scratch) instead of one chain per point: at m = 1 (17 points) 4 threads of 4 .. 5 points
// do 1.4 M MAC32 per proof where a thread per point did 3.0 M.  dalek's vartime multiscalar_mul does the same on the CPU
// (Straus below 190 points).  A point that fails to decompress rejects the proof (RangeProof::verify -> Err).
#define RP_V1_PMAX 8
DAPOL_HD_INLINE void rp_store_cached(uint32_t *dst, const ge_cached &c) {
    store8(dst, c.YpX.v); store8(dst + 8, c.YmX.v); store8(dst + 16, c.Z2.v); store8(dst + 24, c.T2d.v);
}
DAPOL_HD_INLINE void rp_load_cached(ge_cached &c, const uint32_t *src) {
    load8(c.YpX.v, src); load8(c.YmX.v, src + 8); load8(c.Z2.v, src + 16); load8(c.T2d.v, src + 24);
}
DAPOL_HD_INLINE void rp_v1_body(const RpBatch &b, uint64_t p, int grp) {
    const int lg = b.lg, m = b.m, nv = rp_nvar(lg, m), G = b.vgroups;
    int8_t dig[RP_V1_PMAX][64];
    int cnt = 0;
#pragma unroll 1
    for (int q = grp; q < nv && cnt < RP_V1_PMAX; q += G, cnt++) {
        const uint32_t *src;
        if (q < 4) src = b.proof_in + p * (b.plen / 4) + 8 * q;
        else if (q < 4 + 2 * lg) src = b.proof_in + p * (b.plen / 4) + 56 + 8 * (q - 4);
        else src = b.coms + (p * m + (q - 4 - 2 * lg)) * 8;
        uint32_t w[8];
        load8(w, src);
        ge cur;
        if (!ge_decompress(cur, w)) {
            b.status[p] = 0;
#pragma unroll 1
            for (int k = 0; k < 64; k++) dig[cnt][k] = 0;
            continue;
        }
        uint32_t *tb = b.vartab + (p * nv + q) * 8 * 32;
        ge_cached c1, c;
        ge_to_cached(c1, cur);
        rp_store_cached(tb, c1);
#pragma unroll 1
        for (int i = 1; i < 8; i++) {
            ge_cadd(cur, cur, c1, 0);
            ge_to_cached(c, cur);
            rp_store_cached(tb + 32 * i, c);
        }
        sc s;
        rp_ld(s, b.varsc + (p * nv + q) * 8);
        int32_t d[64];
        sc_signed_digits<4, 64>(d, s.v, 8);
#pragma unroll 1
        for (int k = 0; k < 64; k++) dig[cnt][k] = (int8_t)d[k];
    }
    ge acc;
    ge_identity(acc);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        if (k != 63) { ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); ge_dbl(acc, acc); }
#pragma unroll 1
        for (int c = 0; c < cnt; c++) {
            int dk = dig[c][k];
            if (dk != 0) {
                int neg = dk < 0;
                ge_cached e;
                rp_load_cached(e, b.vartab + ((p * nv + (uint64_t)(grp + c * G)) * 8 + (uint32_t)((neg ? -dk : dk) - 1)) * 32);
                ge_cadd(acc, acc, e, neg);
            }
        }
    }
    rp_store_ext(b.varpts + (p * nv + grp) * 32, acc);
}
// V2 (CTA per proof, partial sums): the fixed-base part of the verification equation plus the variable partial points
//   G_I: -z - a s_I      H_I: z + y^-I (z^(2+j) 2^i - b s_{N-1-I})      B, B_blinding: scalars from V0
template <int W, bool INL = false>
DAPOL_HD_INLINE void rp_v2_partial(ge &acc, const RpBatch &b, uint64_t p, uint32_t tid, uint32_t T) {
    ge_identity(acc);
    const uint32_t N = (uint32_t)b.N;
    const int nv = rp_nvar(b.lg, b.m);
    sc z, mz, a, bb;
    rp_ld(z, rp_ch(b, p, CH_Z)); rp_ld(mz, rp_ch(b, p, CH_MZ)); rp_ld(a, rp_ch(b, p, CH_A)); rp_ld(bb, rp_ch(b, p, CH_B));
#pragma unroll 1
    for (uint32_t t = tid; t < 2 * N; t += T) {
        sc s, c;
        if (t < N) {
            rp_ld(s, b.svec + (p * N + t) * 8);
            sc_mul(s, a, s); sc_sub(s, mz, s);
            rp_fixed_mul_acc<W, INL>(acc, b.tabG, rp_gen_of(b, t), s);
        } else {
            uint32_t I = t - N;
            rp_ld(s, b.svec + (p * N + (N - 1 - I)) * 8);
            sc_mul(s, bb, s);
            rp_zz2(c, b, p, I);
            sc_sub(s, c, s);
            rp_ld(c, b.ypow + (p * N + I) * 8);
            sc_mul(s, s, c);
            sc_add(s, z, s);
            rp_fixed_mul_acc<W, INL>(acc, b.tabH, rp_gen_of(b, I), s);
        }
    }
    sc s;  // the B and B_blinding terms, one window per thread; the partial sums of the variable points (V1), one per thread
    rp_ld(s, rp_ch(b, p, CH_SB));
    rp_fixed_mul_spread<W>(acc, b.tabB, 0, s, tid, T);
    rp_ld(s, rp_ch(b, p, CH_SBBL));
    rp_fixed_mul_spread<W>(acc, b.tabBbl, 0, s, tid, T);
#pragma unroll 1
    for (uint32_t t = tid; t < (uint32_t)b.vgroups; t += T) {
        ge q;
        rp_load_ext(q, b.varpts + (p * nv + t) * 32);
        ge_add(acc, acc, q);
    }
}

// ================================================================================================ batched verification (bucket method)
// A batch of proofs is accepted as a whole when ONE random linear combination of their verification equations holds:
//   sum_p rho_p * (equation of proof p) == identity,     rho_p = secret random weights, one per proof
// (bulletproofs' own batching idea, verify_multiple_with_rng's weight c, carried across proofs).  The generators are shared,
// so their scalars add up mod l and the fixed-base part costs ONE table multi-scalar multiplication per GROUP of proofs instead
// of one per proof; the variable points (A, S, T_1, T_2, L_k, R_k, V_j of every proof: G * nv of them) form one large
// variable-base MSM, which is where the BUCKET METHOD (Pippenger) pays: each point is added once per c-bit window into the
// bucket of its signed digit -- 253 / c additions per point, no doublings -- and the buckets of a window are folded by
// running sums.  A group that fails is re-verified proof by proof with the Straus path above, so the verdicts are exactly
// those of the per-proof verifier (a bad proof passes a group with probability 2^-252).
struct RpbPlan {
    int G;                 // proofs per group
    // Windows: NW = ceil(254 / c) of them cover the 253 bits of a scalar.  Windows 0 .. NW-2 carry SIGNED digits |d| <= 2^(c-1); the
    // top window is unsigned and c - 1 bits wide, so there is no carry out of it and its digits spread over half of its buckets
    // instead of piling into one or two (a narrow top window would put every point of the group into the same bucket -- one thread
    // adding G * nv points).  The slack bits NW c - 1 - 253 come off the lowest windows, one bit each.
    int c, NW, slack;
    int L;                 // buckets per chunk of the bucket fold (power of two)
    int P;                 // proofs per partial sum of the combined scalars
    uint64_t groups;       // ceil(K / G)
    uint32_t wseed[8];     // ChaCha20 key of the weights
    uint32_t *rho;         // [K][8]
    uint32_t *cached;      // [K * nv][32] variable points in cached form (Y+X, Y-X, 2Z, 2dT)
    uint32_t *keys_in, *keys;    // [K * nv * NW] bucket id of every (point, window) term (zero digit: the id past the last bucket), unsorted / sorted
    uint32_t *vals_in, *vals;    // [K * nv * NW] (point index << 1) | sign
    uint32_t *bucket;      // [groups * NW * 2^(c-1)][32] bucket sums (extended)
    uint32_t *chunk_run, *chunk_tot;  // [groups * NW * nchunks][32]
    uint32_t *window;      // [groups * NW][32]
    uint32_t *gpart;       // [groups][ceil(G / P)][2N + 2][8] partial sums of the combined scalars
    uint32_t *gsc;         // [groups][2N + 2][8] combined scalars of G_i, H_i, B, B_blinding
    uint32_t *gfix;        // [groups][32] fixed-base part of the group's combination (extended)
    int *gbad;             // [groups] set when a proof of the group failed a format / decompression check
    int *gok;              // [groups] 1 = the group's combination is the identity and every proof passed the format checks
};
DAPOL_HD_INLINE uint64_t rpb_buckets_per_window(const RpbPlan &pl) { return 1ull << (pl.c - 1); }
DAPOL_HD_INLINE void rpb_set_windows(RpbPlan &pl, int c) { pl.c = c; pl.NW = (254 + c - 1) / c; pl.slack = pl.NW * c - 1 - 253; }
DAPOL_HD_INLINE uint32_t rpb_window_bits(const RpbPlan &pl, int k) { return (uint32_t)(k == pl.NW - 1 || k < pl.slack ? pl.c - 1 : pl.c); }
DAPOL_HD_INLINE uint32_t rpb_window_offset(const RpbPlan &pl, int k) { return (uint32_t)(k * pl.c - (k < pl.slack ? k : pl.slack)); }
// weight of proof p: draw p of ChaCha20(wseed); never zero in practice (probability 2^-252; a zero weight only weakens the check)
DAPOL_HD_INLINE void rpb_weight_body(const RpbPlan &pl, uint64_t p) {
    uint32_t ks[16];
    chacha20_block(ks, pl.wseed, p, 0x70697070656e6765ull);
    sc r;
    sc_from_wide(r, ks);
    rp_st(pl.rho + p * 8, r);
}
// VB1 (thread per variable point (p, q)): decompress, weight the scalar, emit one (bucket, point) term per window
DAPOL_HD_INLINE void rpb_terms_body(const RpBatch &b, const RpbPlan &pl, uint64_t p, int q) {
    const int lg = b.lg, m = b.m, nv = rp_nvar(lg, m);
    const uint64_t pt = p * nv + q;
    const uint32_t *src;
    if (q < 4) src = b.proof_in + p * (b.plen / 4) + 8 * q;
    else if (q < 4 + 2 * lg) src = b.proof_in + p * (b.plen / 4) + 56 + 8 * (q - 4);
    else src = b.coms + (p * m + (q - 4 - 2 * lg)) * 8;
    uint32_t w[8];
    load8(w, src);
    ge P;
    int good = ge_decompress(P, w);
    if (!good) { b.status[p] = 0; ge_identity(P); }  // RangeProof::verify -> Err for a point that does not decompress
    ge_cached cp;
    ge_to_cached(cp, P);
    rp_store_cached(pl.cached + pt * 32, cp);
    sc s, rho;
    rp_ld(s, b.varsc + pt * 8); rp_ld(rho, pl.rho + p * 8);
    sc_mul(s, s, rho);
    const uint64_t nb = rpb_buckets_per_window(pl), g = p / (uint64_t)pl.G;
    const uint32_t none = (uint32_t)(pl.groups * (uint64_t)pl.NW * nb);
    uint32_t carry = 0;
#pragma unroll 1
    for (int k = 0; k < pl.NW; k++) {
        const uint32_t c = rpb_window_bits(pl, k);
        uint32_t bit = rpb_window_offset(pl, k), wi = bit >> 5, sh = bit & 31, raw = 0;
        if (wi < 8) {
            raw = s.v[wi] >> sh;
            if (sh + c > 32 && wi + 1 < 8) raw |= s.v[wi + 1] << (32 - sh);
        }
        raw = (raw & ((1u << c) - 1u)) + carry;
        carry = (k != pl.NW - 1 && raw > (1u << (c - 1))) ? 1u : 0u;  // the top window keeps its value: at most 2^(c-1) with the carry in
        int32_t d = (int32_t)raw - (int32_t)(carry << c);
        uint32_t neg = d < 0, mag = (uint32_t)(neg ? -d : d);
        const uint64_t e = pt * (uint64_t)pl.NW + k;
        pl.keys_in[e] = (mag && good) ? (uint32_t)((g * pl.NW + k) * nb + (mag - 1)) : none;
        pl.vals_in[e] = ((uint32_t)pt << 1) | neg;
    }
}
// VB2 (thread per bucket): sum of the points whose digit in this window selects the bucket -- a run of the sorted terms
DAPOL_HD_INLINE void rpb_bucket_body(const RpbPlan &pl, uint64_t bk, uint64_t n_terms) {
    uint64_t lo = 0, hi = n_terms;
    const uint32_t want = (uint32_t)bk;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (pl.keys[mid] < want) lo = mid + 1; else hi = mid; }
    ge acc;
    ge_identity(acc);
#pragma unroll 1
    for (uint64_t j = lo; j < n_terms && pl.keys[j] == want; j++) {
        const uint32_t v = pl.vals[j];
        ge_cached cp;
        rp_load_cached(cp, pl.cached + (uint64_t)(v >> 1) * 32);
        ge_cadd(acc, acc, cp, (int)(v & 1u));
    }
    rp_store_ext(pl.bucket + bk * 32, acc);
}
// VB3 (thread per chunk of L buckets of one window): run = sum S_b, tot = sum (i + 1) S_(b0 + i), by running sums from the top
DAPOL_HD_INLINE void rpb_chunk_body(const RpbPlan &pl, uint64_t ch) {
    const uint64_t b0 = ch * (uint64_t)pl.L;
    ge run, tot, s;
    ge_identity(run); ge_identity(tot);
#pragma unroll 1
    for (int i = pl.L - 1; i >= 0; i--) {
        rp_load_ext(s, pl.bucket + (b0 + i) * 32);
        ge_add(run, run, s);
        ge_add(tot, tot, run);
    }
    rp_store_ext(pl.chunk_run + ch * 32, run);
    rp_store_ext(pl.chunk_tot + ch * 32, tot);
}
// VB4 (thread per (group, window)): sum_b (b + 1) S_b = sum_ch tot_ch + L * sum_(ch >= 1) (run_ch + run_(ch+1) + ...)
DAPOL_HD_INLINE void rpb_window_body(const RpbPlan &pl, uint64_t gw) {
    const uint64_t nch = rpb_buckets_per_window(pl) / (uint64_t)pl.L;
    ge suf, hi, lo, s;
    ge_identity(suf); ge_identity(hi); ge_identity(lo);
#pragma unroll 1
    for (int64_t ch = (int64_t)nch - 1; ch >= 0; ch--) {
        rp_load_ext(s, pl.chunk_tot + (gw * nch + ch) * 32);
        ge_add(lo, lo, s);
        if (ch >= 1) {
            rp_load_ext(s, pl.chunk_run + (gw * nch + ch) * 32);
            ge_add(suf, suf, s);
            ge_add(hi, hi, suf);
        }
    }
    for (int L = pl.L; L > 1; L >>= 1) ge_dbl(hi, hi);
    ge_add(lo, lo, hi);
    rp_store_ext(pl.window + gw * 32, lo);
}
// VB5 (thread per group): Horner over the windows, plus the fixed-base part; the group passes iff the sum is the identity
// and no proof of the group failed a format / decompression check
DAPOL_HD_INLINE void rpb_group_body(const RpBatch &b, const RpbPlan &pl, uint64_t g) {
    ge acc, s;
    ge_identity(acc);
#pragma unroll 1
    for (int w = pl.NW - 1; w >= 0; w--) {
        if (w != pl.NW - 1)
            for (uint32_t i = 0; i < rpb_window_bits(pl, w); i++) ge_dbl(acc, acc);
        rp_load_ext(s, pl.window + (g * pl.NW + w) * 32);
        ge_add(acc, acc, s);
    }
    rp_load_ext(s, pl.gfix + g * 32);
    ge_add(acc, acc, s);
    pl.gok[g] = ge_is_identity(acc) && !pl.gbad[g];
}
// (thread per proof) a proof that failed a format / decompression check fails its group
DAPOL_HD_INLINE void rpb_status_body(const RpBatch &b, const RpbPlan &pl, uint64_t p) {
    if (!b.status[p]) pl.gbad[p / (uint64_t)pl.G] = 1;
}
// VBc (thread per (group, part, t)): weighted sum over P proofs of the group of the scalar of generator t --
// t < N: G_t; t < 2N: H_(t-N); 2N: B; 2N + 1: B_blinding (the per-proof scalars are those of rp_v2_partial); then (thread per
// (group, t)) the sum of the parts
DAPOL_HD_INLINE uint64_t rpb_parts(const RpbPlan &pl) { return ((uint64_t)pl.G + pl.P - 1) / pl.P; }
DAPOL_HD_INLINE void rpb_combine_body(const RpBatch &b, const RpbPlan &pl, uint64_t g, uint64_t part, uint32_t t) {
    const uint32_t N = (uint32_t)b.N;
    const uint64_t gend = g * (uint64_t)pl.G + pl.G < b.K ? g * (uint64_t)pl.G + pl.G : b.K;
    const uint64_t p0 = g * (uint64_t)pl.G + part * pl.P, p1 = p0 + pl.P < gend ? p0 + pl.P : gend;
    sc sum;
    sc_set_u64(sum, 0);
#pragma unroll 1
    for (uint64_t p = p0; p < p1; p++) {
        sc s, c, rho;
        if (t < N) {
            sc mz, a;
            rp_ld(mz, rp_ch(b, p, CH_MZ)); rp_ld(a, rp_ch(b, p, CH_A));
            rp_ld(s, b.svec + (p * N + t) * 8);
            sc_mul(s, a, s); sc_sub(s, mz, s);
        } else if (t < 2 * N) {
            const uint32_t I = t - N;
            sc z, bb;
            rp_ld(z, rp_ch(b, p, CH_Z)); rp_ld(bb, rp_ch(b, p, CH_B));
            rp_ld(s, b.svec + (p * N + (N - 1 - I)) * 8);
            sc_mul(s, bb, s);
            rp_zz2(c, b, p, I);
            sc_sub(s, c, s);
            rp_ld(c, b.ypow + (p * N + I) * 8);
            sc_mul(s, s, c);
            sc_add(s, z, s);
        } else rp_ld(s, rp_ch(b, p, t == 2 * N ? CH_SB : CH_SBBL));
        rp_ld(rho, pl.rho + p * 8);
        sc_mul(s, s, rho);
        sc_add(sum, sum, s);
    }
    rp_st(pl.gpart + ((g * rpb_parts(pl) + part) * (2 * N + 2) + t) * 8, sum);
}
DAPOL_HD_INLINE void rpb_combine_sum_body(const RpBatch &b, const RpbPlan &pl, uint64_t g, uint32_t t) {
    const uint64_t per = 2ull * b.N + 2, parts = rpb_parts(pl);
    sc sum, s;
    sc_set_u64(sum, 0);
#pragma unroll 1
    for (uint64_t q = 0; q < parts; q++) {
        rp_ld(s, pl.gpart + ((g * parts + q) * per + t) * 8);
        sc_add(sum, sum, s);
    }
    rp_st(pl.gsc + (g * per + t) * 8, sum);
}
// VBf (CTA per group, partial sums): the group's fixed-base part from the generator tables
template <int W, bool INL = false>
DAPOL_HD_INLINE void rpb_fixed_partial(ge &acc, const RpBatch &b, const RpbPlan &pl, uint64_t g, uint32_t tid, uint32_t T) {
    ge_identity(acc);
    const uint32_t N = (uint32_t)b.N;
#pragma unroll 1
    for (uint32_t t = tid; t < 2 * N + 2; t += T) {
        sc s;
        rp_ld(s, pl.gsc + (g * (2 * N + 2) + t) * 8);
        if (t < N) rp_fixed_mul_acc<W, INL>(acc, b.tabG, rp_gen_of(b, t), s);
        else if (t < 2 * N) rp_fixed_mul_acc<W, INL>(acc, b.tabH, rp_gen_of(b, t - N), s);
        else if (t == 2 * N) rp_fixed_mul_acc<W, INL>(acc, b.tabB, 0, s);
        else rp_fixed_mul_acc<W, INL>(acc, b.tabBbl, 0, s);
    }
}
