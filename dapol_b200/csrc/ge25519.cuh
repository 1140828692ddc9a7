// ristretto255 group on top of fe25519.cuh: extended twisted-Edwards points, mixed addition with
// precomputed affine-Niels table entries, doubling, RFC 9496 compress / decompress / one-way map.
// Replaces curve25519-dalek-ng RistrettoPoint::{add, compress, decompress, from_uniform_bytes}
// (/root/reference/src/dapol/node.rs:31,66-76; src/proof/node.rs:58,87-88) and
// PedersenGens::commit's scalar multiplication (fixed-base signed-window comb here).
#pragma once
#include "fe25519.cuh"
#include "sc25519.cuh"

struct ge {  // extended coordinates, x = X/Z, y = Y/Z, T = XY/Z
    fe X, Y, Z, T;
};
// DAPOL_NIELS_PAD (experiment, profiles/r02_variants.txt): table entries in 128-byte slots, so that an entry never straddles two
// 128-byte lines (a 96-byte entry at a 96-byte stride does, two times out of three)
struct
#ifdef DAPOL_NIELS_PAD
    alignas(128)
#endif
    ge_niels {  // affine precomputed: (y+x, y-x, 2d*x*y) -- 96 bytes
    fe ypx, ymx, t2d;
};
struct ge_cached {  // projective precomputed: (Y+X, Y-X, 2Z, 2d*T) -- 128 bytes
    fe YpX, YmX, Z2, T2d;
};

DAPOL_HD_INLINE void ge_identity(ge &p) {
    fe_set0(p.X); fe_set1(p.Y); fe_set1(p.Z); fe_set0(p.T);
}
DAPOL_HD_INLINE void ge_basepoint(ge &p) {
    p.X = fe_const_bx(); p.Y = fe_const_by(); fe_set1(p.Z); p.T = fe_const_bt();
}
DAPOL_HD_INLINE void ge_basepoint_half(ge &p) {  // (1/2 mod l) * B: v * B as a half point needs no scalar halving
    p.X = fe_const_bhx(); p.Y = fe_const_bhy(); fe_set1(p.Z); p.T = fe_const_bht();
}
DAPOL_HD_INLINE void ge_bblinding(ge &p) {  // PedersenGens::default().B_blinding (bulletproofs generators.rs)
    p.X = fe_const_bblx(); p.Y = fe_const_bbly(); fe_set1(p.Z); p.T = fe_const_bblt();
}
DAPOL_HD_INLINE void ge_neg(ge &r, const ge &p) {
    fe_neg(r.X, p.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T);
}

// r = p + q, both extended (add-2008-hwcd-3, a = -1): 9M
DAPOL_HD_INLINE void ge_add(ge &r, const ge &p, const ge &q) {
    fe A, B, C, D, t0, t1;
    fe_sub(t0, p.Y, p.X); fe_sub(t1, q.Y, q.X); fe_mul(A, t0, t1);
    fe_add(t0, p.Y, p.X); fe_add(t1, q.Y, q.X); fe_mul(B, t0, t1);
    fe_mul(C, p.T, q.T); fe_mul(C, C, fe_const_d2());
    fe_mul(D, p.Z, q.Z); fe_dbl(D, D);
    fe_sub(t0, B, A);  // E
    fe_sub(t1, D, C);  // F
    fe_add(D, D, C);   // G
    fe_add(B, B, A);   // H
    fe_mul(r.X, t0, t1); fe_mul(r.Y, D, B); fe_mul(r.Z, t1, D); fe_mul(r.T, t0, B);
}
DAPOL_HD_INLINE void ge_sub(ge &r, const ge &p, const ge &q) {
    ge n;
    ge_neg(n, q);
    ge_add(r, p, n);
}
// r = p + sign*q with q affine-Niels: 7M.  neg != 0 subtracts.
// MUL = fe_mul (a called function on the device, fe25519.cuh) or fe_mul_inl (inlined)
#define DAPOL_GE_MADD_BODY(MUL)                                      \
    fe A, B, C, D, t0, t1, qa, qb;                                   \
    qa = q.ymx; qb = q.ypx;                                          \
    fe_cmov(qa, q.ypx, neg); fe_cmov(qb, q.ymx, neg);                \
    fe_sub(t0, p.Y, p.X); MUL(A, t0, qa);                            \
    fe_add(t0, p.Y, p.X); MUL(B, t0, qb);                            \
    MUL(C, p.T, q.t2d); fe_cneg(C, neg);                             \
    fe_dbl(D, p.Z);                                                  \
    fe_sub(t0, B, A); /* E */                                        \
    fe_sub(t1, D, C); /* F */                                        \
    fe_add(D, D, C);  /* G */                                        \
    fe_add(B, B, A);  /* H */                                        \
    MUL(r.X, t0, t1); MUL(r.Y, D, B); MUL(r.Z, t1, D); MUL(r.T, t0, B);
DAPOL_HD_INLINE void ge_madd(ge &r, const ge &p, const ge_niels &q, int neg) {
    DAPOL_GE_MADD_BODY(fe_mul)
}
// the same with the seven products inlined (19 KB of code): for loops that every warp of the SM runs in step (the MSM kernels
// of the aggregated range proofs: 128-thread CTAs, hundreds of terms per thread), where the instruction cache holds the one
// loop body and the register moves of a call are pure overhead on the multiply pipe
DAPOL_HD_INLINE void ge_madd_inl(ge &r, const ge &p, const ge_niels &q, int neg) {
    DAPOL_GE_MADD_BODY(fe_mul_inl)
}
DAPOL_HD_INLINE void ge_to_cached(ge_cached &c, const ge &p) {
    fe_add(c.YpX, p.Y, p.X); fe_sub(c.YmX, p.Y, p.X); fe_dbl(c.Z2, p.Z); fe_mul(c.T2d, p.T, fe_const_d2());
}
// r = p + sign*q with q projective cached: 8M
DAPOL_HD_INLINE void ge_cadd(ge &r, const ge &p, const ge_cached &q, int neg) {
    fe A, B, C, D, t0, t1, qa, qb;
    qa = q.YmX; qb = q.YpX;
    fe_cmov(qa, q.YpX, neg); fe_cmov(qb, q.YmX, neg);
    fe_sub(t0, p.Y, p.X); fe_mul(A, t0, qa);
    fe_add(t0, p.Y, p.X); fe_mul(B, t0, qb);
    fe_mul(C, p.T, q.T2d); fe_cneg(C, neg);
    fe_mul(D, p.Z, q.Z2);
    fe_sub(t0, B, A); fe_sub(t1, D, C); fe_add(D, D, C); fe_add(B, B, A);
    fe_mul(r.X, t0, t1); fe_mul(r.Y, D, B); fe_mul(r.Z, t1, D); fe_mul(r.T, t0, B);
}
// r = 2p (dbl-2008-hwcd, a = -1): 4S + 4M
DAPOL_HD_INLINE void ge_dbl(ge &r, const ge &p) {
    fe A, B, C, E, G, F, H, t;
    fe_sq(A, p.X); fe_sq(B, p.Y); fe_sq(C, p.Z); fe_dbl(C, C);
    fe_add(t, p.X, p.Y); fe_sq(E, t); fe_sub(E, E, A); fe_sub(E, E, B);
    fe_sub(G, B, A);          // D + B with D = -A
    fe_sub(F, G, C);
    fe_add(H, A, B); fe_neg(H, H);  // D - B = -(A + B)
    fe_mul(r.X, E, F); fe_mul(r.Y, G, H); fe_mul(r.Z, F, G); fe_mul(r.T, E, H);
}
// affine-Niels form of p (one inversion)
DAPOL_HD_INLINE void ge_to_niels(ge_niels &n, const ge &p) {
    fe zi, x, y;
    fe_invert(zi, p.Z);
    fe_mul(x, p.X, zi); fe_mul(y, p.Y, zi);
    fe_add(n.ypx, y, x); fe_sub(n.ymx, y, x);
    fe_mul(n.t2d, x, y); fe_mul(n.t2d, n.t2d, fe_const_d2());
    // store canonical words so table bytes are deterministic
    uint32_t w[8];
    fe_canon(w, n.ypx); fe_fromwords(n.ypx, w);
    fe_canon(w, n.ymx); fe_fromwords(n.ymx, w);
    fe_canon(w, n.t2d); fe_fromwords(n.t2d, w);
}
DAPOL_HD_INLINE int ge_is_identity(const ge &p) {  // RistrettoPoint == identity  <=>  X == 0 or Y == 0
    return fe_iszero(p.X) | fe_iszero(p.Y);
}

// RFC 9496 4.2 SQRT_RATIO_M1(u, v): r = sqrt(u/v) (non-negative) if square, else sqrt(i*u/v)
DAPOL_HD_INLINE int fe_sqrt_ratio_m1(fe &r, const fe &u, const fe &v) {
    fe v3, v7, t, check, nu, nui;
    fe_sq(v3, v); fe_mul(v3, v3, v);
    fe_sq(v7, v3); fe_mul(v7, v7, v);
    fe_mul(t, u, v7);
    fe_pow22523(t, t);
    fe_mul(r, u, v3); fe_mul(r, r, t);
    fe_sq(check, r); fe_mul(check, check, v);
    fe_neg(nu, u);
    fe_mul(nui, nu, fe_const_sqrtm1());
    int correct = fe_eq(check, u), flipped = fe_eq(check, nu), flipped_i = fe_eq(check, nui);
    fe ri;
    fe_mul(ri, r, fe_const_sqrtm1());
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r);
    return correct | flipped;
}
// 1/sqrt(v) for v known to be a non-zero square times {1, -1, i, -i}: the u = 1 case of the above,
// with the three equality tests folded into one canonicalisation (hot: once per tree node).
DAPOL_HD_INLINE int fe_invsqrt(fe &r, const fe &v) {
    fe v3, v7, t, check;
    fe_sq(v3, v); fe_mul(v3, v3, v);
    fe_sq(v7, v3); fe_mul(v7, v7, v);
    fe_pow22523(t, v7);
    fe_mul(r, v3, t);
    fe_sq(check, r); fe_mul(check, check, v);
    uint32_t c[8], m1[8], im[8], one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    fe_canon(c, check);
    fe nu; fe_set1(nu); fe_neg(nu, nu); fe_canon(m1, nu);           // -1
    fe ni; fe_neg(ni, fe_const_sqrtm1()); fe_canon(im, ni);         // -i
    uint32_t d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { d1 |= c[i] ^ one[i]; d2 |= c[i] ^ m1[i]; d3 |= c[i] ^ im[i]; }
    int correct = d1 == 0, flipped = d2 == 0, flipped_i = d3 == 0;
    fe ri;
    fe_mul(ri, r, fe_const_sqrtm1());
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r);
    return correct | flipped;
}

// RFC 9496 4.3.2 Encode -> 8 LE words
DAPOL_HD_INLINE void ge_compress(uint32_t s[8], const ge &p) {
    fe u1, u2, t0, t1, isq, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv;
    fe_add(t0, p.Z, p.Y); fe_sub(t1, p.Z, p.Y); fe_mul(u1, t0, t1);
    fe_mul(u2, p.X, p.Y);
    fe_sq(t0, u2); fe_mul(t0, t0, u1);
    fe_invsqrt(isq, t0);
    fe_mul(den1, isq, u1); fe_mul(den2, isq, u2);
    fe_mul(z_inv, den1, den2); fe_mul(z_inv, z_inv, p.T);
    fe_mul(ix0, p.X, fe_const_sqrtm1()); fe_mul(iy0, p.Y, fe_const_sqrtm1());
    fe_mul(ench, den1, fe_const_invsqrt_a_minus_d());
    fe_mul(t0, p.T, z_inv);
    int rotate = fe_isneg(t0);
    x = p.X; y = p.Y; den_inv = den2;
    fe_cmov(x, iy0, rotate); fe_cmov(y, ix0, rotate); fe_cmov(den_inv, ench, rotate);
    fe_mul(t0, x, z_inv);
    fe_cneg(y, fe_isneg(t0));
    fe_sub(t0, p.Z, y); fe_mul(t0, t0, den_inv);
    fe_abs(t0);
    fe_canon(s, t0);
}
// ---- batched compress(2Q): dalek RistrettoPoint::double_and_compress_batch ------------------------------
// compress() spends ~252 squarings on an inverse square root per point.  For a point known as 2Q the root is
// explicit, so one field INVERSION shared by a whole batch (Montgomery's trick) replaces the per-point chain.
// The tree keeps every commitment as its half point Q (com = 2Q: scalars are halved mod l before the comb,
// parents are Q_L + Q_R), so every node is compressed this way.  Same group element => same canonical bytes.
struct ge_dc_state {
    fe e, f, g, h, x;  // x = e*g*f*h, the value to invert
};
DAPOL_HD_INLINE void ge_dc_prepare(ge_dc_state &st, const ge &p) {
    fe XX, YY, ZZ, dTT, t;
    fe_sq(XX, p.X); fe_sq(YY, p.Y); fe_sq(ZZ, p.Z);
    fe_sq(dTT, p.T); fe_mul(dTT, dTT, fe_const_d());
    fe_dbl(t, p.Y); fe_mul(st.e, p.X, t);   // 2XY
    fe_add(st.f, ZZ, dTT);                  // Z^2 + dT^2
    fe_add(st.g, YY, XX);                   // Y^2 - aX^2
    fe_sub(st.h, ZZ, dTT);                  // Z^2 - dT^2
    fe eg, fh;
    fe_mul(eg, st.e, st.g); fe_mul(fh, st.f, st.h);
    fe_mul(st.x, eg, fh);
}
// xinv = 1/st.x (or 0 when st.x == 0, which only happens for the identity coset: result is then 0 = identity)
DAPOL_HD_INLINE void ge_dc_finish(uint32_t s[8], const ge_dc_state &st, const fe &xinv) {
    fe eg, fh, Zinv, Tinv, t, e, g, h, magic, minus_e, fsq;
    fe_mul(eg, st.e, st.g); fe_mul(fh, st.f, st.h);
    fe_mul(Zinv, eg, xinv); fe_mul(Tinv, fh, xinv);
    fe_mul(t, eg, Zinv);
    int nc1 = fe_isneg(t);
    fe_neg(minus_e, st.e);
    fe_mul(fsq, st.f, fe_const_sqrtm1());
    e = st.e; g = st.g; h = st.h; magic = fe_const_invsqrt_a_minus_d();
    fe_cmov(e, st.g, nc1); fe_cmov(g, minus_e, nc1); fe_cmov(h, fsq, nc1); fe_cmov(magic, fe_const_sqrtm1(), nc1);
    fe_mul(t, h, e); fe_mul(t, t, Zinv);
    fe_cneg(g, fe_isneg(t));
    fe_mul(t, g, Tinv); fe_mul(t, magic, t);
    fe_sub(h, h, g); fe_mul(t, h, t);
    fe_abs(t);
    fe_canon(s, t);
}
// Up to B points per thread; arrays live in local memory (loops are kept rolled).
template <int B>
struct ge_dc_batch {
    ge_dc_state st[B];
    fe pre[B];  // prefix products; after solve(): the compressed words
    fe acc;
    int n;
    DAPOL_HD_MEMBER void init() { fe_set1(acc); n = 0; }
    DAPOL_HD_MEMBER void push(const ge &q) {
        ge_dc_prepare(st[n], q);
        fe x = st[n].x, one;
        fe_set1(one);
        fe_cmov(x, one, fe_iszero(x));
        pre[n] = acc;
        fe_mul(acc, acc, x);
        n++;
    }
    DAPOL_HD_MEMBER void solve() {
        fe inv;
        fe_invert(inv, acc);
#pragma unroll 1
        for (int b = n - 1; b >= 0; b--) {
            fe x = st[b].x, one, xinv;
            fe_set1(one);
            int z = fe_iszero(x);
            fe_cmov(x, one, z);
            fe_mul(xinv, inv, pre[b]);
            fe_mul(inv, inv, x);
            fe zero;
            fe_set0(zero);
            fe_cmov(xinv, zero, z);
            uint32_t s[8];
            ge_dc_finish(s, st[b], xinv);
            fe_setwords_raw(pre[b], s);
        }
    }
    DAPOL_HD_MEMBER void get(int b, uint32_t s[8]) const { fe_getwords_raw(s, pre[b]); }
};

// RFC 9496 4.3.1 Decode from 8 LE words; returns 1 on success
DAPOL_HD_INLINE int ge_decompress(ge &p, const uint32_t s[8]) {
    fe sf, ss, u1, u2, u2s, v, t0, isq, den_x, den_y, one;
    uint32_t chk[8];
    fe_fromwords(sf, s);
    fe_canon(chk, sf);
    uint32_t diff = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) diff |= chk[i] ^ s[i];
    int bad = (diff != 0) | (int)(s[0] & 1u);  // non-canonical or negative
    fe_set1(one);
    fe_sq(ss, sf); fe_sub(u1, one, ss); fe_add(u2, one, ss); fe_sq(u2s, u2);
    fe_sq(t0, u1); fe_mul(t0, t0, fe_const_d()); fe_neg(t0, t0); fe_sub(v, t0, u2s);
    fe_mul(t0, v, u2s);
    int ok = fe_sqrt_ratio_m1(isq, one, t0);
    fe_mul(den_x, isq, u2);
    fe_mul(den_y, isq, den_x); fe_mul(den_y, den_y, v);
    fe_dbl(t0, sf); fe_mul(p.X, t0, den_x); fe_abs(p.X);
    fe_mul(p.Y, u1, den_y);
    fe_set1(p.Z);
    fe_mul(p.T, p.X, p.Y);
    return (!bad) & ok & !fe_isneg(p.T) & !fe_iszero(p.Y);
}
// RFC 9496 4.3.4 MAP (dalek elligator_ristretto_flavor)
DAPOL_HD_INLINE void ge_elligator(ge &p, const fe &t) {
    fe r, u, v, s, sp, c, N, w0, w1, w2, w3, one, t0, t1;
    fe_set1(one);
    fe_sq(r, t); fe_mul(r, r, fe_const_sqrtm1());
    fe_add(u, r, one); fe_mul(u, u, fe_const_one_minus_d_sq());
    fe_mul(t0, r, fe_const_d()); fe_neg(t1, one); fe_sub(t0, t1, t0);
    fe_add(t1, r, fe_const_d()); fe_mul(v, t0, t1);
    int ok = fe_sqrt_ratio_m1(s, u, v);
    fe_mul(sp, s, t); fe_abs(sp); fe_neg(sp, sp);
    fe_cmov(s, sp, !ok);
    fe_neg(c, one); fe_cmov(c, r, !ok);
    fe_sub(t0, r, one); fe_mul(N, c, t0); fe_mul(N, N, fe_const_d_minus_one_sq()); fe_sub(N, N, v);
    fe_mul(w0, s, v); fe_dbl(w0, w0);
    fe_mul(w1, N, fe_const_sqrt_ad_minus_one());
    fe_sq(t0, s); fe_sub(w2, one, t0); fe_add(w3, one, t0);
    fe_mul(p.X, w0, w3); fe_mul(p.Y, w2, w1); fe_mul(p.Z, w1, w3); fe_mul(p.T, w0, w2);
}
// RistrettoPoint::from_uniform_bytes on 16 LE words
DAPOL_HD_INLINE void ge_from_uniform(ge &p, const uint32_t w[16]) {
    fe t1, t2;
    ge p1, p2;
    fe_fromwords(t1, w); fe_fromwords(t2, w + 8);
    ge_elligator(p1, t1); ge_elligator(p2, t2);
    ge_add(p, p1, p2);
}

DAPOL_HD_INLINE void load_niels(ge_niels &q, const ge_niels *src) {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    load8(q.ypx.v, s); load8(q.ymx.v, s + 8); load8(q.t2d.v, s + 16);
}


// acc = sign * q (q affine Niels): X = 2x, Y = 2y, Z = 2, T = 2xy = t2d / d -- one multiplication instead of the seven of a
// mixed addition into the identity (the first window of every comb)
DAPOL_HD_INLINE void ge_from_niels(ge &r, const ge_niels &q, int neg) {
    fe_sub(r.X, q.ypx, q.ymx);
    fe_add(r.Y, q.ypx, q.ymx);
    fe_set_u32(r.Z, 2);
    fe_mul(r.T, q.t2d, fe_const_dinv());
    fe_cneg(r.X, neg); fe_cneg(r.T, neg);
}

// Window of the table of B/2 (values are 64-bit) that goes with window W of the B_blinding table.  Up to 16 bits
// both tables use W (L2-resident tables); the wide HBM-resident windows pair with 22 bits for the value table
// (3 windows, 0.6 GB: a wider one would cost HBM without saving an addition).
template <int W>
struct comb_value_window {
    static constexpr int value = W <= 16 ? W : 22;
};

struct NodeStore {
    uint64_t *idx;     // [T] tree index (path bits) of the node inside its level
    uint64_t *v;       // [T] liability sum
    uint32_t *r;       // [T][8] blinding factor words (leaves: as given, possibly unreduced; else canonical)
    uint32_t *comc;    // [T][8] compress(com)
    uint32_t *hash;    // [T][8] node hash
    uint32_t *ext;     // [T][32] com in extended coordinates X,Y,Z,T (build-time only)
    uint8_t *is_pad;   // [T]
    uint32_t *hash_hi; // [T][8] upper half of a 64-byte node hash (D = Blake2b, DAPOL_HASH_BLAKE2B); nullptr for 32-byte digests
};

DAPOL_HD_INLINE void store_ge(uint32_t *dst, const ge &p) {
    store8(dst, p.X.v); store8(dst + 8, p.Y.v); store8(dst + 16, p.Z.v); store8(dst + 24, p.T.v);
}
DAPOL_HD_INLINE void load_ge(ge &p, const uint32_t *src) {
    load8(p.X.v, src); load8(p.Y.v, src + 8); load8(p.Z.v, src + 16); load8(p.T.v, src + 24);
}

// ---- fixed-base signed-window comb: table[k][e] = (e+1) * 2^(W k) * P as affine Niels ------------
// acc += sum_k d[k] * 2^(W k) * P, digits from sc_signed_digits<W, NW>.
// FRESH: acc is known to be the identity on entry, so the first non-zero window initialises it (ge_from_niels).
//
// Table look-ups are 96-byte reads at data-dependent addresses (L2 hits for the 27 MB tables of window 15, HBM reads for the
// wide windows): the entry of window k + 1 is requested before the addition of window k starts.  Measured and rejected
// (profiles/r01d_variants.txt): prefetch.global.L2 of all entries of a scalar up front (+8 % time: 60 MB of prefetched lines
// do not survive in L2 until used), and staging through shared memory with a cp.async ring of 4 / 6 entries per thread
// (+6 .. 14 %: the wait / ld.shared pair per window costs more than the latency it hides).
template <int W, int NW>
DAPOL_HD_INLINE const ge_niels *comb_entry(const ge_niels *table, const int32_t d[NW], int k) {
    int32_t dk = d[k];
    uint32_t e = dk ? (uint32_t)(dk < 0 ? -dk : dk) - 1u : 0u;
    return table + ((size_t)k * (1u << (W - 1)) + e);
}
template <int W, int NW, bool FRESH = false, bool INL = false>
DAPOL_HD_INLINE void ge_comb_accumulate(ge &acc, const ge_niels *__restrict__ table, const int32_t d[NW]) {
    int fresh = FRESH;
    ge_niels q, qn;
    load_niels(qn, comb_entry<W, NW>(table, d, 0));
#pragma unroll 1
    for (int k = 0; k < NW; k++) {
        int32_t dk = d[k];
        q = qn;
        if (k + 1 < NW) load_niels(qn, comb_entry<W, NW>(table, d, k + 1));
        if (dk != 0) {
            if (fresh) { ge_from_niels(acc, q, dk < 0); fresh = 0; }
            else if (INL) ge_madd_inl(acc, acc, q, dk < 0);
            else ge_madd(acc, acc, q, dk < 0);
        }
    }
}
