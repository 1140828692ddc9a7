// Scalars mod l = 2^252 + 27742317777372353535851937790883648493 on 8 x 32-bit limbs.
// Montgomery reduction with R = 2^256; values are kept canonical (< l) in normal form.
// Replaces curve25519-dalek-ng Scalar for the hot path: Scalar::from_bits / + / * /
// from_bytes_mod_order_wide / invert as reached from /root/reference/src/dapol/mod.rs:385,
// src/dapol/node.rs:75 (v_blinding sum) and bulletproofs' prover/verifier.
#pragma once
#include "fe25519.cuh"

struct sc {
    uint32_t v[8];
};

DAPOL_HD_INLINE uint32_t sc_l_word(int i) {
    const uint32_t L[8] = {SC_L_WORDS};
    return L[i];
}
DAPOL_HD_INLINE sc sc_const_r1() { sc r = {{SC_R1_WORDS}}; return r; }
DAPOL_HD_INLINE sc sc_const_rr() { sc r = {{SC_RR_WORDS}}; return r; }

DAPOL_HD_INLINE void sc_set_u64(sc &r, uint64_t x) {
    r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32);
#pragma unroll
    for (int i = 2; i < 8; i++) r.v[i] = 0;
}
DAPOL_HD_INLINE int sc_iszero(const sc &a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.v[i];
    return x == 0;
}
// x (9 words, < 2l) -> x mod l
DAPOL_HD_INLINE void sc_condsub(sc &r, const uint32_t x[9]) {
    const uint32_t L[8] = {SC_L_WORDS};
    uint32_t d[8];
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)x[i] - (int64_t)L[i];
        d[i] = (uint32_t)bw;
        bw >>= 32;
    }
    bw += (int64_t)x[8];
    uint32_t keep = (uint32_t)0 - (uint32_t)(bw < 0);  // all-ones: x < l, keep x
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (x[i] & keep) | (d[i] & ~keep);
}
// r = a*b/R mod l.  Requires a*b < R*l (e.g. b < l, a any 256-bit value); result canonical.
DAPOL_HD_INLINE void sc_montmul(sc &r, const sc &a, const sc &b) {
    const uint32_t L[8] = {SC_L_WORDS};
    uint32_t T[18];
    mul_wide_8x8(T, a.v, b.v);
    T[16] = 0; T[17] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = T[i] * SC_N0;
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)m * L[j] + T[i + j];
            T[i + j] = (uint32_t)c;
            c >>= 32;
        }
#pragma unroll
        for (int k = i + 8; k < 17; k++) {
            c += T[k];
            T[k] = (uint32_t)c;
            c >>= 32;
        }
    }
    sc_condsub(r, T + 8);
}
DAPOL_HD_INLINE void sc_mul(sc &r, const sc &a, const sc &b) {  // a any 256-bit, b < l
    sc t;
    sc_montmul(t, a, b);
    sc_montmul(r, t, sc_const_rr());
}
DAPOL_HD_INLINE void sc_reduce256(sc &r, const sc &a) { sc_montmul(r, a, sc_const_r1()); }
// r = a / 2 mod l for any 256-bit a (canonical result): one Montgomery product with 2^-1 * R
DAPOL_HD_INLINE void sc_half256(sc &r, const sc &a) {
    sc h = {{SC_INV2R_WORDS}};
    sc_montmul(r, a, h);
}
// Scalar::from_bytes_mod_order_wide on 16 LE words
DAPOL_HD_INLINE void sc_from_wide(sc &r, const uint32_t w[16]) {
    sc lo, hi;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[8 + i]; }
    sc_montmul(lo, lo, sc_const_r1());  // lo mod l
    sc_montmul(hi, hi, sc_const_rr());  // hi * R mod l
    uint32_t x[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)lo.v[i] + hi.v[i]; x[i] = (uint32_t)c; c >>= 32; }
    x[8] = (uint32_t)c;
    sc_condsub(r, x);
}
// Scalar::from_bytes_mod_order_wide on 16 LE words together with its half: h = r / 2 mod l, r = 2 h mod l, both canonical.
// The halving rides on the two Montgomery products of the wide reduction (constants 2^-1 R and 2^-1 R^2) instead of a third.
DAPOL_HD_INLINE void sc_from_wide_with_half(sc &r, sc &h, const uint32_t w[16]) {
    sc lo, hi;
    const sc c1 = {{SC_INV2R_WORDS}}, c2 = {{SC_INV2RR_WORDS}};
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[8 + i]; }
    sc_montmul(lo, lo, c1);  // lo / 2 mod l
    sc_montmul(hi, hi, c2);  // hi * 2^256 / 2 mod l
    uint32_t x[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)lo.v[i] + hi.v[i]; x[i] = (uint32_t)c; c >>= 32; }
    x[8] = (uint32_t)c;
    sc_condsub(h, x);
    c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)h.v[i] + h.v[i]; x[i] = (uint32_t)c; c >>= 32; }
    x[8] = (uint32_t)c;
    sc_condsub(r, x);
}
DAPOL_HD_INLINE void sc_add(sc &r, const sc &a, const sc &b) {  // a, b < l
    uint32_t x[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; x[i] = (uint32_t)c; c >>= 32; }
    x[8] = (uint32_t)c;
    sc_condsub(r, x);
}
DAPOL_HD_INLINE void sc_neg(sc &r, const sc &a) {  // a < l
    const uint32_t L[8] = {SC_L_WORDS};
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= a.v[i];
    uint32_t m = (uint32_t)0 - (uint32_t)(nz != 0);
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)L[i] - (int64_t)a.v[i];
        r.v[i] = (uint32_t)bw & m;
        bw >>= 32;
    }
}
DAPOL_HD_INLINE void sc_sub(sc &r, const sc &a, const sc &b) {
    sc n;
    sc_neg(n, b);
    sc_add(r, a, n);
}
// 1 if the 8 LE words are a canonical scalar (< l)  (Scalar::from_canonical_bytes)
DAPOL_HD_INLINE int sc_is_canonical(const uint32_t w[8]) {
    const uint32_t L[8] = {SC_L_WORDS};
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { bw += (int64_t)w[i] - (int64_t)L[i]; bw >>= 32; }
    return bw < 0;
}
// a^(l-2) mod l, in the Montgomery domain (a < l, a != 0)
DAPOL_HD_INLINE void sc_invert(sc &r, const sc &a) {
    uint32_t E[8] = {SC_L_WORDS};
    E[0] -= 2u;  // l - 2 (no borrow: low word of l is 0x5cf5d3ed)
    sc am, acc;
    sc_montmul(am, a, sc_const_rr());  // a*R
    acc = sc_const_r1();               // 1*R
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        sc_montmul(acc, acc, acc);
        if ((E[i >> 5] >> (i & 31)) & 1u) sc_montmul(acc, acc, am);
    }
    sc one;
    sc_set_u64(one, 1);
    sc_montmul(r, acc, one);  // leave the Montgomery domain
}

// signed radix-2^W digits: a = sum d[k] * 2^(W k), -2^(W-1) <= d[k] <= 2^(W-1), NW = bits/W + 1 windows
template <int W, int BITS>
struct sc_windows {
    static constexpr int NW = BITS / W + 1;
};
template <int W, int NW>
DAPOL_HD_INLINE void sc_signed_digits(int32_t d[NW], const uint32_t *words, int nwords) {
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        int bit = k * W;
        int wi = bit >> 5, sh = bit & 31;
        uint32_t raw = 0;
        if (wi < nwords) {
            raw = words[wi] >> sh;
            if (sh + W > 32 && wi + 1 < nwords) raw |= words[wi + 1] << (32 - sh);
        }
        raw = (raw & ((1u << W) - 1u)) + carry;
        carry = raw > (1u << (W - 1)) ? 1u : 0u;
        d[k] = (int32_t)raw - (int32_t)(carry << W);
    }
}
