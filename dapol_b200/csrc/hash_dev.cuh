// Device hashes for the DAPOL+ hot path (all 32-bit ALU-pipe work, word-oriented, little-endian):
//   BLAKE3 (single chunk, <= 1024 B) and BLAKE2s-256 as the node digest D
//     (reference: D::new/update/finalize in /root/reference/src/dapol/node.rs:33-36,66-70 and
//      src/dapol/mod.rs:347-384,418-424; blake3 is the bench's D, Blake2s the KAT's),
//   ChaCha20 block for the seeded padding / prover RNG contract (replaces thread_rng, node.rs:87),
//   Keccak-f[1600] for SHAKE256 / SHA3-512 / STROBE-128 (bulletproofs generators + merlin).
#pragma once
#include <stdint.h>
#include "fe25519.cuh"

#define DAPOL_HASH_BLAKE3 0
#define DAPOL_HASH_BLAKE2S 1

DAPOL_HD_INLINE uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
DAPOL_HD_INLINE uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
DAPOL_HD_INLINE uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> ((64 - n) & 63)); }

#define BLAKE_IV0 0x6A09E667u
#define BLAKE_IV1 0xBB67AE85u
#define BLAKE_IV2 0x3C6EF372u
#define BLAKE_IV3 0xA54FF53Au
#define BLAKE_IV4 0x510E527Fu
#define BLAKE_IV5 0x9B05688Cu
#define BLAKE_IV6 0x1F83D9ABu
#define BLAKE_IV7 0x5BE0CD19u

#define BLAKE_G(a, b, c, d, x, y)                 \
    do {                                          \
        a = a + b + (x); d = rotr32(d ^ a, 16);   \
        c = c + d;       b = rotr32(b ^ c, 12);   \
        a = a + b + (y); d = rotr32(d ^ a, 8);    \
        c = c + d;       b = rotr32(b ^ c, 7);    \
    } while (0)

// ---------------------------------------------------------------- BLAKE3
#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_ROOT 8u
#define B3_PARENT 4u
#define B3_MAX_DEPTH 16  // chaining-value stack of the tree mode: inputs of up to 2^16 chunks (64 MB)

#define B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    BLAKE_G(s0, s4, s8, s12, m[m0], m[m1]);                                            \
    BLAKE_G(s1, s5, s9, s13, m[m2], m[m3]);                                            \
    BLAKE_G(s2, s6, s10, s14, m[m4], m[m5]);                                           \
    BLAKE_G(s3, s7, s11, s15, m[m6], m[m7]);                                           \
    BLAKE_G(s0, s5, s10, s15, m[m8], m[m9]);                                           \
    BLAKE_G(s1, s6, s11, s12, m[m10], m[m11]);                                         \
    BLAKE_G(s2, s7, s8, s13, m[m12], m[m13]);                                          \
    BLAKE_G(s3, s4, s9, s14, m[m14], m[m15]);

// cv <- compress(cv, m, counter, block_len, flags); message schedule fully unrolled (indices static).  The counter is the
// chunk number (0 for every input of at most one chunk and for parent nodes).
DAPOL_HD_INLINE void blake3_compress_t(uint32_t cv[8], const uint32_t m[16], uint32_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = BLAKE_IV0, s9 = BLAKE_IV1, s10 = BLAKE_IV2, s11 = BLAKE_IV3, s12 = counter, s13 = 0, s14 = block_len, s15 = flags;
    B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8; cv[1] = s1 ^ s9; cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}
DAPOL_HD_INLINE void blake3_compress(uint32_t cv[8], const uint32_t m[16], uint32_t block_len, uint32_t flags) {
    blake3_compress_t(cv, m, 0u, block_len, flags);
}
DAPOL_HD_INLINE void blake3_iv(uint32_t cv[8]) {
    cv[0] = BLAKE_IV0; cv[1] = BLAKE_IV1; cv[2] = BLAKE_IV2; cv[3] = BLAKE_IV3;
    cv[4] = BLAKE_IV4; cv[5] = BLAKE_IV5; cv[6] = BLAKE_IV6; cv[7] = BLAKE_IV7;
}

// ---------------------------------------------------------------- BLAKE2s-256
#define B2S_ROUND(a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15) \
    BLAKE_G(v0, v4, v8, v12, m[a0], m[a1]);                                             \
    BLAKE_G(v1, v5, v9, v13, m[a2], m[a3]);                                             \
    BLAKE_G(v2, v6, v10, v14, m[a4], m[a5]);                                            \
    BLAKE_G(v3, v7, v11, v15, m[a6], m[a7]);                                            \
    BLAKE_G(v0, v5, v10, v15, m[a8], m[a9]);                                            \
    BLAKE_G(v1, v6, v11, v12, m[a10], m[a11]);                                          \
    BLAKE_G(v2, v7, v8, v13, m[a12], m[a13]);                                           \
    BLAKE_G(v3, v4, v9, v14, m[a14], m[a15]);

DAPOL_HD_INLINE void blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t, int last) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = BLAKE_IV0, v9 = BLAKE_IV1, v10 = BLAKE_IV2, v11 = BLAKE_IV3;
    uint32_t v12 = BLAKE_IV4 ^ t, v13 = BLAKE_IV5, v14 = last ? ~BLAKE_IV6 : BLAKE_IV6, v15 = BLAKE_IV7;
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}
DAPOL_HD_INLINE void blake2s_iv(uint32_t h[8]) {
    blake3_iv(h);
    h[0] ^= 0x01010020u;
}

// ---------------------------------------------------------------- BLAKE2b-512 (64-byte digests: blake2::Blake2b, src/tests.rs:104-105)
// The reference reaches it through new_blank + build only (Dapol::new insists on 32-byte digests, mod.rs:101-103).  A node's
// 64-byte hash is kept as two 32-byte halves (NodeStore::hash / hash_hi), so every 32-byte code path is untouched.
#ifndef DAPOL_HASH_BLAKE2B
#define DAPOL_HASH_BLAKE2B 2
#endif
#define B2B_G(a, b, c, d, x, y)                        \
    a = a + b + (x); d = rotl64(d ^ a, 32);            \
    c = c + d;       b = rotl64(b ^ c, 40);            \
    a = a + b + (y); d = rotl64(d ^ a, 48);            \
    c = c + d;       b = rotl64(b ^ c, 1);
#define B2B_ROUND(a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15) \
    B2B_G(v0, v4, v8, v12, m[a0], m[a1]);                                               \
    B2B_G(v1, v5, v9, v13, m[a2], m[a3]);                                               \
    B2B_G(v2, v6, v10, v14, m[a4], m[a5]);                                              \
    B2B_G(v3, v7, v11, v15, m[a6], m[a7]);                                              \
    B2B_G(v0, v5, v10, v15, m[a8], m[a9]);                                              \
    B2B_G(v1, v6, v11, v12, m[a10], m[a11]);                                            \
    B2B_G(v2, v7, v8, v13, m[a12], m[a13]);                                             \
    B2B_G(v3, v4, v9, v14, m[a14], m[a15]);
DAPOL_HD_INLINE void blake2b_compress(uint64_t h[8], const uint64_t m[16], uint64_t t, int last) {
    uint64_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint64_t v8 = 0x6a09e667f3bcc908ull, v9 = 0xbb67ae8584caa73bull, v10 = 0x3c6ef372fe94f82bull, v11 = 0xa54ff53a5f1d36f1ull;
    uint64_t v12 = 0x510e527fade682d1ull ^ t, v13 = 0x9b05688c2b3e6c1full, v14 = 0x1f83d9abfb41bd6bull, v15 = 0x5be0cd19137e2179ull;
    if (last) v14 = ~v14;
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2B_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2B_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2B_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2B_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2B_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2B_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2B_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2B_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}
DAPOL_HD_INLINE void blake2b_iv512(uint64_t h[8]) {
    h[0] = 0x6a09e667f3bcc908ull ^ 0x01010040ull; h[1] = 0xbb67ae8584caa73bull; h[2] = 0x3c6ef372fe94f82bull; h[3] = 0xa54ff53a5f1d36f1ull;
    h[4] = 0x510e527fade682d1ull; h[5] = 0x9b05688c2b3e6c1full; h[6] = 0x1f83d9abfb41bd6bull; h[7] = 0x5be0cd19137e2179ull;
}
DAPOL_HD_INLINE uint64_t b2b_pair(const uint32_t *w) { return (uint64_t)w[0] | ((uint64_t)w[1] << 32); }
DAPOL_HD_INLINE void b2b_out(uint32_t lo[8], uint32_t hi[8], const uint64_t h[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        lo[2 * i] = (uint32_t)h[i]; lo[2 * i + 1] = (uint32_t)(h[i] >> 32);
        hi[2 * i] = (uint32_t)h[4 + i]; hi[2 * i + 1] = (uint32_t)(h[4 + i] >> 32);
    }
}
// Blake2b(32-byte string) -> 64 bytes as two halves        -- DapolNode::new: D(compress(com))   (node.rs:33-36)
DAPOL_HD_INLINE void dapol_b2b_hash32(uint32_t lo[8], uint32_t hi[8], const uint32_t in[8]) {
    uint64_t h[8], m[16];
    blake2b_iv512(h);
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = b2b_pair(in + 2 * i);
#pragma unroll
    for (int i = 4; i < 16; i++) m[i] = 0;
    blake2b_compress(h, m, 32, 1);
    b2b_out(lo, hi, h);
}
// Blake2b(C(L) || C(R) || H(L) || H(R)), 32 + 32 + 64 + 64 = 192 bytes   -- Mergeable::merge (node.rs:66-70)
DAPOL_HD_INLINE void dapol_b2b_hash192(uint32_t lo[8], uint32_t hi[8], const uint32_t cl[8], const uint32_t cr[8], const uint32_t hl_lo[8],
                                       const uint32_t hl_hi[8], const uint32_t hr_lo[8], const uint32_t hr_hi[8]) {
    uint64_t h[8], m[16];
    blake2b_iv512(h);
#pragma unroll
    for (int i = 0; i < 4; i++) { m[i] = b2b_pair(cl + 2 * i); m[4 + i] = b2b_pair(cr + 2 * i); m[8 + i] = b2b_pair(hl_lo + 2 * i); m[12 + i] = b2b_pair(hl_hi + 2 * i); }
    blake2b_compress(h, m, 128, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) { m[i] = b2b_pair(hr_lo + 2 * i); m[4 + i] = b2b_pair(hr_hi + 2 * i); }
#pragma unroll
    for (int i = 8; i < 16; i++) m[i] = 0;
    blake2b_compress(h, m, 192, 1);
    b2b_out(lo, hi, h);
}

// ---------------------------------------------------------------- D over fixed-size node inputs
// hash = D(32-byte string)          -- DapolNode::new: D(compress(com))        (node.rs:33-36)
DAPOL_HD_INLINE void dapol_hash32(int hash_id, uint32_t out[8], const uint32_t in[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = in[i]; m[8 + i] = 0; }
    if (hash_id == DAPOL_HASH_BLAKE3) {
        blake3_iv(out);
        blake3_compress(out, m, 32, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT);
    } else {
        blake2s_iv(out);
        blake2s_compress(out, m, 32, 1);
    }
}
// hash = D(C(L) || C(R) || H(L) || H(R)), 128 bytes   -- Mergeable::merge (node.rs:66-70)
DAPOL_HD_INLINE void dapol_hash128(int hash_id, uint32_t out[8], const uint32_t cl[8], const uint32_t cr[8],
                                   const uint32_t hl[8], const uint32_t hr[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = cl[i]; m[8 + i] = cr[i]; }
    if (hash_id == DAPOL_HASH_BLAKE3) {
        blake3_iv(out);
        blake3_compress(out, m, 64, B3_CHUNK_START);
#pragma unroll
        for (int i = 0; i < 8; i++) { m[i] = hl[i]; m[8 + i] = hr[i]; }
        blake3_compress(out, m, 64, B3_CHUNK_END | B3_ROOT);
    } else {
        blake2s_iv(out);
        blake2s_compress(out, m, 64, 0);
#pragma unroll
        for (int i = 0; i < 8; i++) { m[i] = hl[i]; m[8 + i] = hr[i]; }
        blake2s_compress(out, m, 128, 1);
    }
}

// Incremental D over a byte stream assembled from several parts (leaf derivation, mod.rs:347-384): ids of any length hash as
// the reference's D does.  BLAKE3 inputs longer than one 1024-byte chunk use the tree mode of the spec (chunk chaining values
// on a small stack, left subtree = largest power of two of chunks); returns 0 on success.
struct dapol_hasher {
    uint32_t h[8];
    uint32_t m[16];  // current block, little-endian packed
    uint32_t fill;   // bytes in m
    uint32_t total;  // bytes compressed so far (before m); BLAKE3: inside the current chunk
    int hash_id;
    uint32_t chunk;    // BLAKE3: number of the current chunk
    uint32_t stack_n;  // BLAKE3: chaining values of completed subtrees
    uint32_t stack[B3_MAX_DEPTH][8];
};
// parent node of the BLAKE3 tree: cv <- compress(IV, left || right, PARENT [| ROOT])
DAPOL_HD_INLINE void blake3_parent(uint32_t cv[8], const uint32_t left[8], const uint32_t right[8], uint32_t root) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = left[i]; m[8 + i] = right[i]; }
    blake3_iv(cv);
    blake3_compress(cv, m, 64, B3_PARENT | root);
}
DAPOL_HD_INLINE void hasher_init(dapol_hasher &s, int hash_id) {
    s.hash_id = hash_id; s.fill = 0; s.total = 0; s.chunk = 0; s.stack_n = 0;
    if (hash_id == DAPOL_HASH_BLAKE3) blake3_iv(s.h); else blake2s_iv(s.h);
#pragma unroll
    for (int i = 0; i < 16; i++) s.m[i] = 0;
}
DAPOL_HD_INLINE void hasher_flush_full(dapol_hasher &s) {  // compress a full block that more input follows
    if (s.hash_id == DAPOL_HASH_BLAKE3) {
        if (s.total == 960) {  // last block of a chunk that is not the last chunk: its chaining value goes on the stack
            blake3_compress_t(s.h, s.m, s.chunk, 64, B3_CHUNK_END);
            uint32_t cv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) cv[i] = s.h[i];
            for (uint32_t done = s.chunk + 1; (done & 1u) == 0 && s.stack_n > 0; done >>= 1) {  // merge completed subtrees
                uint32_t l[8], p[8];
                s.stack_n--;
                for (int i = 0; i < 8; i++) l[i] = s.stack[s.stack_n][i];
                blake3_parent(p, l, cv, 0u);
                for (int i = 0; i < 8; i++) cv[i] = p[i];
            }
            if (s.stack_n < B3_MAX_DEPTH) { for (int i = 0; i < 8; i++) s.stack[s.stack_n][i] = cv[i]; }
            s.stack_n++;
            blake3_iv(s.h);
            s.chunk++; s.fill = 0; s.total = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) s.m[i] = 0;
            return;
        }
        blake3_compress_t(s.h, s.m, s.chunk, 64, s.total == 0 ? B3_CHUNK_START : 0u);
    } else blake2s_compress(s.h, s.m, s.total + 64, 0);
    s.total += 64; s.fill = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s.m[i] = 0;
}
DAPOL_HD_INLINE void hasher_update(dapol_hasher &s, const uint8_t *p, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        if (s.fill == 64) hasher_flush_full(s);
        uint32_t w = s.fill >> 2, sh = (s.fill & 3) * 8;
        // dynamic word index: keep it a select chain so m[] stays in registers
#pragma unroll
        for (int k = 0; k < 16; k++) if ((uint32_t)k == w) s.m[k] |= (uint32_t)p[i] << sh;
        s.fill++;
    }
}
DAPOL_HD_INLINE void hasher_update_words(dapol_hasher &s, const uint32_t *w, int nwords) {
    for (int i = 0; i < nwords; i++) {
        uint8_t b[4] = {(uint8_t)w[i], (uint8_t)(w[i] >> 8), (uint8_t)(w[i] >> 16), (uint8_t)(w[i] >> 24)};
        hasher_update(s, b, 4);
    }
}
DAPOL_HD_INLINE int hasher_final(dapol_hasher &s, uint32_t out[8]) {
    if (s.hash_id == DAPOL_HASH_BLAKE3) {
        if (s.stack_n > B3_MAX_DEPTH) return -1;  // longer than 2^16 chunks
        blake3_compress_t(s.h, s.m, s.chunk, s.fill, (s.total == 0 ? B3_CHUNK_START : 0u) | B3_CHUNK_END | (s.stack_n == 0 ? B3_ROOT : 0u));
        while (s.stack_n > 0) {  // fold the stack: the last parent is the root
            uint32_t l[8], r[8];
            s.stack_n--;
            for (int i = 0; i < 8; i++) { l[i] = s.stack[s.stack_n][i]; r[i] = s.h[i]; }
            blake3_parent(s.h, l, r, s.stack_n == 0 ? B3_ROOT : 0u);
        }
    } else {
        blake2s_compress(s.h, s.m, s.total + s.fill, 1);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = s.h[i];
    return 0;
}

// ---------------------------------------------------------------- ChaCha20 block (rand_chacha ChaCha20Rng layout)
#define CHACHA_QR(a, b, c, d)                                   \
    a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); \
    a += b; d = rotl32(d ^ a, 8);  c += d; b = rotl32(b ^ c, 7);
DAPOL_HD_INLINE void chacha20_block(uint32_t out[16], const uint32_t key[8], uint64_t counter, uint64_t stream) {
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = (uint32_t)stream, x15 = (uint32_t)(stream >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        CHACHA_QR(x0, x4, x8, x12) CHACHA_QR(x1, x5, x9, x13) CHACHA_QR(x2, x6, x10, x14) CHACHA_QR(x3, x7, x11, x15)
        CHACHA_QR(x0, x5, x10, x15) CHACHA_QR(x1, x6, x11, x12) CHACHA_QR(x2, x7, x8, x13) CHACHA_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + 0x61707865u; out[1] = x1 + 0x3320646eu; out[2] = x2 + 0x79622d32u; out[3] = x3 + 0x6b206574u;
    out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
    out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
    out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32);
    out[14] = x14 + (uint32_t)stream; out[15] = x15 + (uint32_t)(stream >> 32);
}

// ---------------------------------------------------------------- Keccak-f[1600]
DAPOL_HD_INLINE void keccak_f1600(uint64_t st[25]) {
    const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c0 = st[0] ^ st[5] ^ st[10] ^ st[15] ^ st[20], c1 = st[1] ^ st[6] ^ st[11] ^ st[16] ^ st[21];
        uint64_t c2 = st[2] ^ st[7] ^ st[12] ^ st[17] ^ st[22], c3 = st[3] ^ st[8] ^ st[13] ^ st[18] ^ st[23];
        uint64_t c4 = st[4] ^ st[9] ^ st[14] ^ st[19] ^ st[24];
        uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
        for (int j = 0; j < 25; j += 5) { st[j] ^= d0; st[j + 1] ^= d1; st[j + 2] ^= d2; st[j + 3] ^= d3; st[j + 4] ^= d4; }
        // rho + pi
        uint64_t b[25];
        b[0] = st[0];
        b[10] = rotl64(st[1], 1);  b[20] = rotl64(st[2], 62); b[5] = rotl64(st[3], 28);  b[15] = rotl64(st[4], 27);
        b[16] = rotl64(st[5], 36); b[1] = rotl64(st[6], 44);  b[11] = rotl64(st[7], 6);  b[21] = rotl64(st[8], 55);
        b[6] = rotl64(st[9], 20);  b[7] = rotl64(st[10], 3);  b[17] = rotl64(st[11], 10); b[2] = rotl64(st[12], 43);
        b[12] = rotl64(st[13], 25); b[22] = rotl64(st[14], 39); b[23] = rotl64(st[15], 41); b[8] = rotl64(st[16], 45);
        b[18] = rotl64(st[17], 15); b[3] = rotl64(st[18], 21); b[13] = rotl64(st[19], 8);  b[14] = rotl64(st[20], 18);
        b[24] = rotl64(st[21], 2);  b[9] = rotl64(st[22], 61); b[19] = rotl64(st[23], 56); b[4] = rotl64(st[24], 14);
#pragma unroll
        for (int j = 0; j < 25; j += 5) {
            st[j] = b[j] ^ (~b[j + 1] & b[j + 2]);
            st[j + 1] = b[j + 1] ^ (~b[j + 2] & b[j + 3]);
            st[j + 2] = b[j + 2] ^ (~b[j + 3] & b[j + 4]);
            st[j + 3] = b[j + 3] ^ (~b[j + 4] & b[j]);
            st[j + 4] = b[j + 4] ^ (~b[j] & b[j + 1]);
        }
        st[0] ^= RC[r];
    }
}
