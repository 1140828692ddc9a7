// GF(2^255-19) on 8 x 32-bit saturated limbs for the B200 integer pipe (IMAD / IMAD.WIDE).
//
// Representation: any 256-bit value x stands for x mod p (p = 2^255-19); 2^256 = 38 (mod p).
// mul/sq = 512-bit product via generated carry chains (fe_mulsqr_gen.inc) + fold by 38.
// Values are only made canonical where the algorithm needs bytes or a sign (fe_tobytes).
//
// Replaces (for the hot path) curve25519-dalek-ng's u64 backend FieldElement51, which the
// reference reaches through PedersenGens::commit / RistrettoPoint::{add,compress}
// (/root/reference/src/dapol/node.rs:31,66-76).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DAPOL_HD __host__ __device__
#define DAPOL_HD_INLINE __host__ __device__ __forceinline__
#define DAPOL_HD_MEMBER __host__ __device__ __forceinline__
#else
#define DAPOL_HD
#define DAPOL_HD_INLINE static inline
#define DAPOL_HD_MEMBER inline
#endif

#ifndef __CUDA_ARCH__
// Host emulation of the PTX carry-chain instructions (tests/host_emu only; cf_ is block-local).
#define EMU_LO_(x, y) ((uint64_t)(uint32_t)((uint64_t)(x) * (uint64_t)(y)))
#define EMU_HI_(x, y) ((uint64_t)(uint32_t)(((uint64_t)(x) * (uint64_t)(y)) >> 32))
#define EMU_SET_(d, t) do { uint64_t t__ = (t); (d) = (uint32_t)t__; cf_ = (uint32_t)(t__ >> 32); } while (0)
#define EMU_SETNC_(d, t) do { uint64_t t__ = (t); (d) = (uint32_t)t__; } while (0)
#define EMU_MUL_LO(d, x, y) (d) = (uint32_t)EMU_LO_(x, y)
#define EMU_MUL_HI(d, x, y) (d) = (uint32_t)EMU_HI_(x, y)
#define EMU_MAD_LO_CC(d, x, y, c) EMU_SET_(d, EMU_LO_(x, y) + (uint64_t)(c))
#define EMU_MAD_HI_CC(d, x, y, c) EMU_SET_(d, EMU_HI_(x, y) + (uint64_t)(c))
#define EMU_MADC_LO_CC(d, x, y, c) EMU_SET_(d, EMU_LO_(x, y) + (uint64_t)(c) + cf_)
#define EMU_MADC_HI_CC(d, x, y, c) EMU_SET_(d, EMU_HI_(x, y) + (uint64_t)(c) + cf_)
#define EMU_MADC_LO(d, x, y, c) EMU_SETNC_(d, EMU_LO_(x, y) + (uint64_t)(c) + cf_)
#define EMU_MADC_HI(d, x, y, c) EMU_SETNC_(d, EMU_HI_(x, y) + (uint64_t)(c) + cf_)
#define EMU_ADD_CC(d, x, y) EMU_SET_(d, (uint64_t)(x) + (uint64_t)(y))
#define EMU_ADDC_CC(d, x, y) EMU_SET_(d, (uint64_t)(x) + (uint64_t)(y) + cf_)
#define EMU_ADDC(d, x, y) EMU_SETNC_(d, (uint64_t)(x) + (uint64_t)(y) + cf_)
#define EMU_ADD(d, x, y) EMU_SETNC_(d, (uint64_t)(x) + (uint64_t)(y))
// sub.cc / subc: CC.CF holds the borrow
#define EMU_SETB_(d, t) do { uint64_t t__ = (t); (d) = (uint32_t)t__; cf_ = (uint32_t)(t__ >> 32) & 1u; } while (0)
#define EMU_SUB_CC(d, x, y) EMU_SETB_(d, (uint64_t)(uint32_t)(x) - (uint64_t)(uint32_t)(y))
#define EMU_SUBC_CC(d, x, y) EMU_SETB_(d, (uint64_t)(uint32_t)(x) - (uint64_t)(uint32_t)(y) - cf_)
#define EMU_SUBC(d, x, y) EMU_SETNC_(d, (uint64_t)(uint32_t)(x) - (uint64_t)(uint32_t)(y) - cf_)
#endif

#ifdef DAPOL_FE_GEN_INC
#include DAPOL_FE_GEN_INC
#else
#include "fe_mulsqr_gen.inc"
#endif

struct fe {
    uint32_t v[8];
};

// 32-byte strings live in device memory as 8 little-endian words; 128-bit vector accesses
DAPOL_HD_INLINE void store8(uint32_t *dst, const uint32_t w[8]) {
#if defined(__CUDA_ARCH__)
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
#else
    for (int i = 0; i < 8; i++) dst[i] = w[i];
#endif
}
DAPOL_HD_INLINE void load8(uint32_t w[8], const uint32_t *src) {
#if defined(__CUDA_ARCH__)
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 a = s[0], b = s[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
#else
    for (int i = 0; i < 8; i++) w[i] = src[i];
#endif
}

DAPOL_HD_INLINE void fe_set0(fe &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
DAPOL_HD_INLINE void fe_set1(fe &r) { fe_set0(r); r.v[0] = 1; }
DAPOL_HD_INLINE void fe_set_u32(fe &r, uint32_t x) { fe_set0(r); r.v[0] = x; }

DAPOL_HD_INLINE void fe_mul_inl(fe &r, const fe &a, const fe &b) {
    uint32_t R[16];
    mul_wide_8x8(R, a.v, b.v);
    fold38(r.v, R);
}
DAPOL_HD_INLINE void fe_sq_inl(fe &r, const fe &a) {
    uint32_t R[16];
    sqr_wide_8(R, a.v);
    fold38(r.v, R);
}
// On the device the multiplication and the squaring are CALLED, not inlined: a product is ~135 instructions (2 KB) and a
// node kernel contains a few hundred of them, so the fully inlined kernels were 150 .. 215 KB of straight-line code against a
// 32 KB instruction cache per SM -- ncu showed 46 % of all warp stalls of k_pad as no_instruction.  Arguments and result
// travel in registers (structs of 8 words by value: no stack traffic), and every warp of the SM now runs the same 3.5 KB.
#ifndef DAPOL_FE_CALL
#define DAPOL_FE_CALL 1
#endif
#if defined(__CUDA_ARCH__) && DAPOL_FE_CALL
static __device__ __noinline__ fe fe_mul_call(fe a, fe b) {
    fe r;
    fe_mul_inl(r, a, b);
    return r;
}
static __device__ __noinline__ fe fe_sq_call(fe a) {
    fe r;
    fe_sq_inl(r, a);
    return r;
}
__device__ __forceinline__ void fe_mul(fe &r, const fe &a, const fe &b) { r = fe_mul_call(a, b); }
__device__ __forceinline__ void fe_sq(fe &r, const fe &a) { r = fe_sq_call(a); }
#else
DAPOL_HD_INLINE void fe_mul(fe &r, const fe &a, const fe &b) { fe_mul_inl(r, a, b); }
DAPOL_HD_INLINE void fe_sq(fe &r, const fe &a) { fe_sq_inl(r, a); }
#endif
// r = a^(2^n), n >= 1 (loop kept rolled: the body is ~110 instructions)
DAPOL_HD_INLINE void fe_sqn(fe &r, const fe &a, int n) {
    fe_sq(r, a);
#pragma unroll 1
    for (int i = 1; i < n; i++) fe_sq(r, r);
}

// r = a + b.  Carry out of 2^256 folds back as +38 (twice: the second can only fire on a tiny value).
DAPOL_HD_INLINE void fe_add(fe &r, const fe &a, const fe &b) {
    uint32_t c;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t m = c * 38u;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(c)
        : "r"(m));
    r.v[0] += c * 38u;
#else
    uint64_t t = 0;
    uint32_t o[8];
    for (int i = 0; i < 8; i++) { t += (uint64_t)a.v[i] + b.v[i]; o[i] = (uint32_t)t; t >>= 32; }
    c = (uint32_t)t;
    t = (uint64_t)c * 38u;
    for (int i = 0; i < 8; i++) { t += o[i]; o[i] = (uint32_t)t; t >>= 32; }
    o[0] += (uint32_t)t * 38u;
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
#endif
}

// r = a - b.  Borrow out of 2^256 folds back as -38.
DAPOL_HD_INLINE void fe_sub(fe &r, const fe &a, const fe &b) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t m = bw & 38u;  // bw is 0 or 0xffffffff
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "=r"(bw)
        : "r"(m));
    r.v[0] -= bw & 38u;
#else
    int64_t t = 0;
    uint32_t o[8];
    for (int i = 0; i < 8; i++) { t += (int64_t)a.v[i] - (int64_t)b.v[i]; o[i] = (uint32_t)t; t >>= 32; }
    uint32_t m = t ? 38u : 0u;
    t = -(int64_t)m;
    for (int i = 0; i < 8; i++) { t += (int64_t)o[i]; o[i] = (uint32_t)t; t >>= 32; }
    o[0] -= t ? 38u : 0u;
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
#endif
}

DAPOL_HD_INLINE void fe_neg(fe &r, const fe &a) {
    fe z;
    fe_set0(z);
    fe_sub(r, z, a);
}
DAPOL_HD_INLINE void fe_dbl(fe &r, const fe &a) { fe_add(r, a, a); }

// canonical little-endian words of a mod p
DAPOL_HD_INLINE void fe_canon(uint32_t o[8], const fe &a) {
    uint32_t t[8];
    // fold bit 255: x = (x mod 2^255) + 19*(x >> 255)
    uint64_t c = 19ull * (a.v[7] >> 31);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (i == 7) ? (uint64_t)(a.v[7] & 0x7fffffffu) : (uint64_t)a.v[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    // now t < 2^255 + 19.  u = t + 19; if u >= 2^255 then t - p = u - 2^255
    uint32_t u[8];
    c = 19;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += t[i];
        u[i] = (uint32_t)c;
        c >>= 32;
    }
    uint32_t ge = (uint32_t)0 - (u[7] >> 31);  // all-ones if t >= p
    u[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = (u[i] & ge) | (t[i] & ~ge);
}
DAPOL_HD_INLINE void fe_tobytes(uint8_t s[32], const fe &a) {
    uint32_t o[8];
    fe_canon(o, a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)o[i];
        s[4 * i + 1] = (uint8_t)(o[i] >> 8);
        s[4 * i + 2] = (uint8_t)(o[i] >> 16);
        s[4 * i + 3] = (uint8_t)(o[i] >> 24);
    }
}
// from 8 LE words; bit 255 ignored (dalek FieldElement::from_bytes)
DAPOL_HD_INLINE void fe_fromwords(fe &r, const uint32_t w[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    r.v[7] &= 0x7fffffffu;
}
// raw 8-word storage inside an fe (used to park already-canonical byte strings in fe-typed scratch)
DAPOL_HD_INLINE void fe_setwords_raw(fe &r, const uint32_t w[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
}
DAPOL_HD_INLINE void fe_getwords_raw(uint32_t w[8], const fe &a) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a.v[i];
}
DAPOL_HD_INLINE int fe_isneg(const fe &a) {
    uint32_t o[8];
    fe_canon(o, a);
    return (int)(o[0] & 1u);
}
DAPOL_HD_INLINE int fe_iszero(const fe &a) {
    uint32_t o[8];
    fe_canon(o, a);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= o[i];
    return x == 0;
}
DAPOL_HD_INLINE int fe_eq(const fe &a, const fe &b) {
    fe d;
    fe_sub(d, a, b);
    return fe_iszero(d);
}
DAPOL_HD_INLINE void fe_cmov(fe &r, const fe &a, int flag) {
    uint32_t m = (uint32_t)0 - (uint32_t)(flag != 0);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (r.v[i] & ~m);
}
DAPOL_HD_INLINE void fe_cneg(fe &r, int flag) {
    fe n;
    fe_neg(n, r);
    fe_cmov(r, n, flag);
}
DAPOL_HD_INLINE void fe_abs(fe &r) { fe_cneg(r, fe_isneg(r)); }

// z^((p-5)/8) = z^(2^252-3)   (dalek field.rs pow_p58): 251 squarings + 12 multiplications
DAPOL_HD_INLINE void fe_pow22523(fe &out, const fe &z) {
    fe t0, t1, t2;
    fe_sq(t0, z);
    fe_sqn(t1, t0, 2);
    fe_mul(t1, z, t1);
    fe_mul(t0, t0, t1);
    fe_sq(t0, t0);
    fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 5);
    fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 10);
    fe_mul(t1, t1, t0);
    fe_sqn(t2, t1, 20);
    fe_mul(t1, t2, t1);
    fe_sqn(t1, t1, 10);
    fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 50);
    fe_mul(t1, t1, t0);
    fe_sqn(t2, t1, 100);
    fe_mul(t1, t2, t1);
    fe_sqn(t1, t1, 50);
    fe_mul(t0, t1, t0);
    fe_sqn(t0, t0, 2);
    fe_mul(out, t0, z);
}
// z^(p-2)
DAPOL_HD_INLINE void fe_invert(fe &out, const fe &z) {
    fe t, z3;
    fe_pow22523(t, z);
    fe_sqn(t, t, 3);
    fe_sq(z3, z);
    fe_mul(z3, z3, z);
    fe_mul(out, t, z3);
}

// field / curve / scalar constants (canonical LE words), generated by tools/gen_consts.py
#define FE_CONST(name, w0, w1, w2, w3, w4, w5, w6, w7) \
    DAPOL_HD_INLINE fe name() { fe r = {{w0, w1, w2, w3, w4, w5, w6, w7}}; return r; }
#include "consts_gen.inc"
