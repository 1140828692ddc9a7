/* TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md).  Parity unpinned vs the Rust
 * reference for bytes; pinned vs RFC 9496 / bulletproofs constants via tests/test_oracle_pins.py.
 *
 * GF(2^255-19) with 5x51-bit limbs, scalars mod l with 4x64 Montgomery, ristretto255 group.
 * Restates the published algorithms of curve25519-dalek-ng ^4.1.1 (field.rs, scalar.rs,
 * ristretto.rs, edwards.rs) which the reference calls from src/dapol/node.rs:31,67-76 and
 * src/range/mod.rs:48-119.  Independent of the CUDA product code (different limb layout). */
#ifndef DOR_EC_H
#define DOR_EC_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static inline void fe_0(fe *h) { memset(h, 0, sizeof *h); }
static inline void fe_1(fe *h) { fe_0(h); h->v[0] = 1; }

static inline void fe_carry(fe *h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static inline void fe_add(fe *h, const fe *f, const fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}
static inline void fe_sub(fe *h, const fe *f, const fe *g) {
    /* add 8p so limbs stay non-negative (inputs carried: limbs < 2^52) */
    h->v[0] = f->v[0] + 0x3FFFFFFFFFFF68ULL - g->v[0];
    for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x3FFFFFFFFFFFF8ULL - g->v[i];
    fe_carry(h);
}
static inline void fe_neg(fe *h, const fe *f) { fe z; fe_0(&z); fe_sub(h, &z, f); }

static inline void fe_mul(fe *h, const fe *f, const fe *g) {
    uint64_t f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    uint64_t g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
    uint64_t g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
    u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
    u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
    u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
    u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & M51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & M51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & M51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & M51;
    c = (uint64_t)(r4 >> 51); h->v[4] = (uint64_t)r4 & M51;
    h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static inline void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }
static inline void fe_sqn(fe *h, const fe *f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }

static inline void fe_frombytes(fe *h, const uint8_t s[32]) { /* ignores bit 255 */
    uint64_t w[4];
    memcpy(w, s, 32);
    h->v[0] = w[0] & M51;
    h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    h->v[4] = (w[3] >> 12) & M51;
}
static inline void fe_tobytes(uint8_t s[32], const fe *f) {
    fe t = *f;
    fe_carry(&t); fe_carry(&t);
    /* canonical: add 19, see if it overflows 2^255 */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    uint64_t w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(s, w, 32);
}
static inline int fe_isneg(const fe *f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static inline int fe_iszero(const fe *f) {
    uint8_t s[32]; fe_tobytes(s, f);
    uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i];
    return r == 0;
}
static inline int fe_eq(const fe *a, const fe *b) {
    uint8_t s[32], t[32]; fe_tobytes(s, a); fe_tobytes(t, b);
    return memcmp(s, t, 32) == 0;
}
static inline void fe_cneg(fe *h, int b) { if (b) fe_neg(h, h); }
static inline void fe_abs(fe *h) { fe_cneg(h, fe_isneg(h)); }

/* z^(2^252-3) = z^((p-5)/8)  (dalek field.rs pow_p58) */
static inline void fe_pow22523(fe *out, const fe *z) {
    fe t0, t1, t2;
    fe_sq(&t0, z);
    fe_sqn(&t1, &t0, 2);
    fe_mul(&t1, z, &t1);
    fe_mul(&t0, &t0, &t1);
    fe_sq(&t0, &t0);
    fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 5);
    fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 10);
    fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 20);
    fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 10);
    fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 50);
    fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 100);
    fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 50);
    fe_mul(&t0, &t1, &t0);
    fe_sqn(&t0, &t0, 2);
    fe_mul(out, &t0, z);
}
static inline void fe_invert(fe *out, const fe *z) { /* z^(p-2) = (z^(2^252-3))^8 * z^3 */
    fe t, z3;
    fe_pow22523(&t, z);
    fe_sqn(&t, &t, 3);
    fe_sq(&z3, z); fe_mul(&z3, &z3, z);
    fe_mul(out, &t, &z3);
}

/* constants, filled by ec_init() from their byte encodings */
static fe FE_D, FE_D2, FE_SQRT_M1, FE_SQRT_AD_MINUS_ONE, FE_INVSQRT_A_MINUS_D, FE_ONE_MINUS_D_SQ, FE_D_MINUS_ONE_SQ;

/* RFC 9496 4.2 SQRT_RATIO_M1 */
static inline int fe_sqrt_ratio_m1(fe *r_out, const fe *u, const fe *v) {
    fe v3, v7, r, check, t, neg_u, neg_u_i;
    fe_sq(&v3, v); fe_mul(&v3, &v3, v);
    fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
    fe_mul(&t, u, &v7);
    fe_pow22523(&t, &t);
    fe_mul(&r, u, &v3); fe_mul(&r, &r, &t);
    fe_sq(&check, &r); fe_mul(&check, &check, v);
    fe_neg(&neg_u, u);
    fe_mul(&neg_u_i, &neg_u, &FE_SQRT_M1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &neg_u), flipped_i = fe_eq(&check, &neg_u_i);
    if (flipped || flipped_i) fe_mul(&r, &r, &FE_SQRT_M1);
    fe_abs(&r);
    *r_out = r;
    return correct || flipped;
}

/* ---------------------------------------------------------------- points (extended) */
typedef struct { fe X, Y, Z, T; } ge;

static inline void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
static inline void ge_add(ge *r, const ge *p, const ge *q) {
    fe A, B, C, D, E, F, G, H, t0, t1;
    fe_sub(&t0, &p->Y, &p->X); fe_sub(&t1, &q->Y, &q->X); fe_mul(&A, &t0, &t1);
    fe_add(&t0, &p->Y, &p->X); fe_add(&t1, &q->Y, &q->X); fe_mul(&B, &t0, &t1);
    fe_mul(&C, &p->T, &q->T); fe_mul(&C, &C, &FE_D2);
    fe_mul(&D, &p->Z, &q->Z); fe_add(&D, &D, &D);
    fe_sub(&E, &B, &A); fe_sub(&F, &D, &C); fe_add(&G, &D, &C); fe_add(&H, &B, &A);
    fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->Z, &F, &G); fe_mul(&r->T, &E, &H);
}
static inline void ge_neg(ge *r, const ge *p) { *r = *p; fe_neg(&r->X, &p->X); fe_neg(&r->T, &p->T); }
static inline void ge_sub(ge *r, const ge *p, const ge *q) { ge n; ge_neg(&n, q); ge_add(r, p, &n); }
static inline void ge_dbl(ge *r, const ge *p) {
    fe A, B, C, Dn, E, G, F, H, t;
    fe_sq(&A, &p->X); fe_sq(&B, &p->Y); fe_sq(&C, &p->Z); fe_add(&C, &C, &C);
    fe_neg(&Dn, &A);
    fe_add(&t, &p->X, &p->Y); fe_sq(&E, &t); fe_sub(&E, &E, &A); fe_sub(&E, &E, &B);
    fe_add(&G, &Dn, &B); fe_sub(&F, &G, &C); fe_sub(&H, &Dn, &B);
    fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->Z, &F, &G); fe_mul(&r->T, &E, &H);
}
static inline int ge_is_identity(const ge *p) { return fe_iszero(&p->X) || fe_iszero(&p->Y); }

/* RFC 9496 4.3.2 */
static inline void ge_compress(uint8_t s[32], const ge *p) {
    fe u1, u2, t0, t1, invsqrt, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv, one;
    fe_1(&one);
    fe_add(&t0, &p->Z, &p->Y); fe_sub(&t1, &p->Z, &p->Y); fe_mul(&u1, &t0, &t1);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t0, &u2); fe_mul(&t0, &t0, &u1);
    fe_sqrt_ratio_m1(&invsqrt, &one, &t0);
    fe_mul(&den1, &invsqrt, &u1); fe_mul(&den2, &invsqrt, &u2);
    fe_mul(&z_inv, &den1, &den2); fe_mul(&z_inv, &z_inv, &p->T);
    fe_mul(&ix0, &p->X, &FE_SQRT_M1); fe_mul(&iy0, &p->Y, &FE_SQRT_M1);
    fe_mul(&ench, &den1, &FE_INVSQRT_A_MINUS_D);
    fe_mul(&t0, &p->T, &z_inv);
    if (fe_isneg(&t0)) { x = iy0; y = ix0; den_inv = ench; } else { x = p->X; y = p->Y; den_inv = den2; }
    fe_mul(&t0, &x, &z_inv);
    fe_cneg(&y, fe_isneg(&t0));
    fe_sub(&t0, &p->Z, &y); fe_mul(&t0, &t0, &den_inv);
    fe_abs(&t0);
    fe_tobytes(s, &t0);
}
/* RFC 9496 4.3.1; returns 0 on failure */
static inline int ge_decompress(ge *p, const uint8_t s[32]) {
    fe sf, ss, u1, u2, u2s, v, t0, invsqrt, den_x, den_y, one;
    uint8_t chk[32];
    fe_frombytes(&sf, s);
    fe_tobytes(chk, &sf);
    if (memcmp(chk, s, 32) != 0 || (s[0] & 1)) return 0; /* non-canonical or negative */
    fe_1(&one);
    fe_sq(&ss, &sf); fe_sub(&u1, &one, &ss); fe_add(&u2, &one, &ss); fe_sq(&u2s, &u2);
    fe_sq(&t0, &u1); fe_mul(&t0, &t0, &FE_D); fe_neg(&t0, &t0); fe_sub(&v, &t0, &u2s);
    fe_mul(&t0, &v, &u2s);
    int ok = fe_sqrt_ratio_m1(&invsqrt, &one, &t0);
    fe_mul(&den_x, &invsqrt, &u2);
    fe_mul(&den_y, &invsqrt, &den_x); fe_mul(&den_y, &den_y, &v);
    fe_add(&t0, &sf, &sf); fe_mul(&p->X, &t0, &den_x); fe_abs(&p->X);
    fe_mul(&p->Y, &u1, &den_y);
    fe_1(&p->Z);
    fe_mul(&p->T, &p->X, &p->Y);
    if (!ok || fe_isneg(&p->T) || fe_iszero(&p->Y)) return 0;
    return 1;
}
/* RFC 9496 4.3.4 MAP */
static inline void ge_elligator(ge *p, const fe *t) {
    fe r, u, v, s, s_prime, c, N, w0, w1, w2, w3, one, t0, t1;
    fe_1(&one);
    fe_sq(&r, t); fe_mul(&r, &r, &FE_SQRT_M1);
    fe_add(&u, &r, &one); fe_mul(&u, &u, &FE_ONE_MINUS_D_SQ);
    fe_mul(&t0, &r, &FE_D); fe_neg(&t1, &one); fe_sub(&t0, &t1, &t0); /* -1 - r*D */
    fe_add(&t1, &r, &FE_D); fe_mul(&v, &t0, &t1);
    int ok = fe_sqrt_ratio_m1(&s, &u, &v);
    fe_mul(&s_prime, &s, t); fe_abs(&s_prime); fe_neg(&s_prime, &s_prime);
    if (!ok) s = s_prime;
    if (ok) fe_neg(&c, &one); else c = r;
    fe_sub(&t0, &r, &one); fe_mul(&N, &c, &t0); fe_mul(&N, &N, &FE_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
    fe_mul(&w0, &s, &v); fe_add(&w0, &w0, &w0);
    fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
    fe_sq(&t0, &s); fe_sub(&w2, &one, &t0); fe_add(&w3, &one, &t0);
    fe_mul(&p->X, &w0, &w3); fe_mul(&p->Y, &w2, &w1); fe_mul(&p->Z, &w1, &w3); fe_mul(&p->T, &w0, &w2);
}
static inline void ge_from_uniform(ge *p, const uint8_t b[64]) {
    fe t1, t2; ge p1, p2;
    fe_frombytes(&t1, b); fe_frombytes(&t2, b + 32);
    ge_elligator(&p1, &t1); ge_elligator(&p2, &t2);
    ge_add(p, &p1, &p2);
}

/* ---------------------------------------------------------------- scalars mod l */
typedef struct { uint64_t v[4]; } sc;
static const uint64_t SC_L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
static uint64_t SC_N0;  /* -l^-1 mod 2^64 */
static sc SC_R1, SC_RR; /* R mod l, R^2 mod l (R = 2^256) */

static inline int sc_geq_l5(const uint64_t t[5]) {
    if (t[4]) return 1;
    for (int i = 3; i >= 0; i--) { if (t[i] > SC_L[i]) return 1; if (t[i] < SC_L[i]) return 0; }
    return 1;
}
static inline void sc_sub_l5(uint64_t t[5]) {
    uint64_t b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - SC_L[i] - b; t[i] = (uint64_t)d; b = (uint64_t)(d >> 64) & 1; }
    t[4] -= b;
}
/* a*b/R mod l for any a,b < 2^256 (CIOS, then subtract l until < l) */
static inline void sc_montmul(sc *r, const sc *a, const sc *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * SC_N0;
        c = ((u128)m * SC_L[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * SC_L[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; c >>= 64;
        t[4] = t[5] + (uint64_t)c;
    }
    while (sc_geq_l5(t)) sc_sub_l5(t);
    memcpy(r->v, t, 32);
}
static inline void sc_mul(sc *r, const sc *a, const sc *b) { sc t; sc_montmul(&t, a, b); sc_montmul(r, &t, &SC_RR); }
static inline void sc_reduce(sc *r, const sc *a) { sc_montmul(r, a, &SC_R1); } /* a mod l for a < 2^256 */
static inline void sc_add(sc *r, const sc *a, const sc *b) { /* inputs < l */
    uint64_t t[5]; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; t[i] = (uint64_t)c; c >>= 64; }
    t[4] = (uint64_t)c;
    while (sc_geq_l5(t)) sc_sub_l5(t);
    memcpy(r->v, t, 32);
}
static inline int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline void sc_neg(sc *r, const sc *a) { /* a < l */
    if (sc_iszero(a)) { *r = *a; return; }
    uint64_t b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)SC_L[i] - a->v[i] - b; r->v[i] = (uint64_t)d; b = (uint64_t)(d >> 64) & 1; }
}
static inline void sc_sub(sc *r, const sc *a, const sc *b) { sc n; sc_neg(&n, b); sc_add(r, a, &n); }
static inline void sc_from_u64(sc *r, uint64_t x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }
static inline void sc_frombytes_reduce(sc *r, const uint8_t b[32]) { sc t; memcpy(t.v, b, 32); sc_reduce(r, &t); }
static inline int sc_frombytes_canonical(sc *r, const uint8_t b[32]) {
    uint64_t t[5]; memcpy(t, b, 32); t[4] = 0;
    if (sc_geq_l5(t)) return 0;
    memcpy(r->v, t, 32);
    return 1;
}
static inline void sc_from_wide(sc *r, const uint8_t b[64]) { /* from_bytes_mod_order_wide */
    sc lo, hi, t;
    memcpy(lo.v, b, 32); memcpy(hi.v, b + 32, 32);
    sc_montmul(&lo, &lo, &SC_R1);  /* lo mod l */
    sc_montmul(&t, &hi, &SC_RR);   /* hi * R mod l */
    sc_add(r, &lo, &t);
}
static inline void sc_tobytes(uint8_t b[32], const sc *a) { memcpy(b, a->v, 32); }
static inline void sc_invert(sc *r, const sc *a) { /* a^(l-2), square and multiply */
    uint64_t e[4] = {SC_L[0] - 2, SC_L[1], SC_L[2], SC_L[3]};
    sc acc; sc_from_u64(&acc, 1);
    for (int i = 255; i >= 0; i--) {
        sc_mul(&acc, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) sc_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* ---------------------------------------------------------------- scalar multiplication */
/* signed radix-16 digits, -8 <= d < 8 (dalek Scalar::to_radix_16), scalar < 2^255 */
static inline void sc_radix16(int8_t d[64], const sc *a) {
    const uint8_t *b = (const uint8_t *)a->v;
    for (int i = 0; i < 32; i++) { d[2 * i] = b[i] & 15; d[2 * i + 1] = (b[i] >> 4) & 15; }
    for (int i = 0; i < 63; i++) { int8_t c = (d[i] + 8) >> 4; d[i] -= c << 4; d[i + 1] += c; }
}
static inline void ge_table8(ge tab[8], const ge *p) { /* 1P..8P */
    tab[0] = *p;
    for (int i = 1; i < 8; i++) ge_add(&tab[i], &tab[i - 1], p);
}
static inline void ge_add_digit(ge *acc, const ge tab[8], int d) {
    if (d > 0) ge_add(acc, acc, &tab[d - 1]);
    else if (d < 0) ge_sub(acc, acc, &tab[-d - 1]);
}
/* Straus over k points (dalek straus.rs): shared doublings, 64 radix-16 digits each */
static inline void ge_straus(ge *r, const sc *scalars, const ge (*tabs)[8], int k) {
    int8_t (*dg)[64] = (int8_t(*)[64])__builtin_alloca((size_t)k * 64);
    for (int i = 0; i < k; i++) sc_radix16(dg[i], &scalars[i]);
    ge acc; ge_identity(&acc);
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); }
        for (int i = 0; i < k; i++) ge_add_digit(&acc, tabs[i], dg[i][w]);
    }
    *r = acc;
}
static inline void ge_scalarmult(ge *r, const sc *s, const ge *p) {
    ge tab[1][8]; ge_table8(tab[0], p);
    ge_straus(r, s, (const ge(*)[8])tab, 1);
}

/* Pippenger MSM (dalek pippenger.rs shape): c-bit unsigned windows, vartime */
static inline void ge_msm(ge *r, const sc *scalars, const ge *points, size_t n) {
    if (n == 0) { ge_identity(r); return; }
    if (n < 16) {
        ge acc; ge_identity(&acc);
        for (size_t i = 0; i < n; i++) { ge t; ge_scalarmult(&t, &scalars[i], &points[i]); ge_add(&acc, &acc, &t); }
        *r = acc; return;
    }
    int c = n < 64 ? 4 : n < 512 ? 6 : n < 4096 ? 8 : 10;
    size_t nb = ((size_t)1 << c) - 1;
    ge *buckets = (ge *)__builtin_alloca(nb * sizeof(ge));
    uint8_t *used = (uint8_t *)__builtin_alloca(nb);
    int windows = (253 + c - 1) / c;
    ge total; ge_identity(&total);
    for (int w = windows - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) ge_dbl(&total, &total);
        memset(used, 0, nb);
        for (size_t i = 0; i < n; i++) {
            int bit = w * c; unsigned d = 0;
            for (int k = 0; k < c && bit + k < 256; k++) d |= (unsigned)((scalars[i].v[(bit + k) >> 6] >> ((bit + k) & 63)) & 1) << k;
            if (!d) continue;
            if (used[d - 1]) ge_add(&buckets[d - 1], &buckets[d - 1], &points[i]);
            else { buckets[d - 1] = points[i]; used[d - 1] = 1; }
        }
        ge run, sum; ge_identity(&run); ge_identity(&sum);
        for (size_t b = nb; b-- > 0;) {
            if (used[b]) ge_add(&run, &run, &buckets[b]);
            ge_add(&sum, &sum, &run);
        }
        ge_add(&total, &total, &sum);
    }
    *r = total;
}

/* ---------------------------------------------------------------- init */
static ge GE_B, GE_BBL;           /* Pedersen gens (bulletproofs generators.rs PedersenGens::default) */
static ge TAB_B[2][8];            /* Straus tables for (B, B_blinding) */

static void fe_from_hex_le(fe *h, const char *hex) {
    uint8_t b[32];
    for (int i = 0; i < 32; i++) { unsigned x; sscanf(hex + 2 * i, "%2x", &x); b[i] = (uint8_t)x; }
    fe_frombytes(h, b);
}
#endif
