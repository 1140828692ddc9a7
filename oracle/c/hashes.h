/* TEST INFRASTRUCTURE ONLY -- CPU oracle hashes: BLAKE3 (single chunk, <= 1024 B), BLAKE2s-256,
 * Keccak-f[1600] (SHA3-512, SHAKE256, STROBE-128/merlin), ChaCha20 block.
 * These restate the published algorithms of the crates the reference uses as `D`
 * (blake3 ^0.3.8: benches/dapol.rs:38; blake2 ^0.9: src/dapol/tests.rs:13) and that
 * bulletproofs/merlin use internally (sha3, keccak). */
#ifndef DOR_HASHES_H
#define DOR_HASHES_H
#include <stdint.h>
#include <string.h>
#include <stddef.h>

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static const uint32_t BLAKE_IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};

#define BLAKE_G(a, b, c, d, x, y)                         \
    do {                                                  \
        a = a + b + (x); d = rotr32(d ^ a, 16);           \
        c = c + d;       b = rotr32(b ^ c, 12);           \
        a = a + b + (y); d = rotr32(d ^ a, 8);            \
        c = c + d;       b = rotr32(b ^ c, 7);            \
    } while (0)

/* ---------------------------------------------------------------- BLAKE3 */
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};

static inline void blake3_compress_t(uint32_t cv[8], const uint8_t block[64], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t m[16], s[16], t[16];
    memcpy(m, block, 64);
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = BLAKE_IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    for (int r = 0; r < 7; r++) {
        BLAKE_G(s[0], s[4], s[8], s[12], m[0], m[1]);
        BLAKE_G(s[1], s[5], s[9], s[13], m[2], m[3]);
        BLAKE_G(s[2], s[6], s[10], s[14], m[4], m[5]);
        BLAKE_G(s[3], s[7], s[11], s[15], m[6], m[7]);
        BLAKE_G(s[0], s[5], s[10], s[15], m[8], m[9]);
        BLAKE_G(s[1], s[6], s[11], s[12], m[10], m[11]);
        BLAKE_G(s[2], s[7], s[8], s[13], m[12], m[13]);
        BLAKE_G(s[3], s[4], s[9], s[14], m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) cv[i] = s[i] ^ s[i + 8];
}
static inline void blake3_compress(uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint32_t flags) {
    blake3_compress_t(cv, block, 0, block_len, flags);
}
/* chaining value of chunk #counter (<= 1024 bytes); root = 1 only for a single-chunk input */
static inline void blake3_chunk_cv(uint32_t cv[8], const uint8_t *in, size_t len, uint64_t counter, int root) {
    memcpy(cv, BLAKE_IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64] = {0};
        size_t off = b * 64, n = len - off < 64 ? len - off : 64;
        memcpy(block, in + off, n);
        uint32_t flags = (b == 0 ? 1u : 0u) | (b == nblocks - 1 ? (2u | (root ? 8u : 0u)) : 0u);
        blake3_compress_t(cv, block, counter, (uint32_t)n, flags);
    }
}
static inline void blake3_parent(uint32_t out[8], const uint32_t l[8], const uint32_t r[8], int root) {
    uint8_t block[64];
    memcpy(block, l, 32); memcpy(block + 32, r, 32);
    memcpy(out, BLAKE_IV, 32);
    blake3_compress_t(out, block, 0, 64, 4u | (root ? 8u : 0u));
}
/* BLAKE3 hash of any length (the tree mode of the spec: 1024-byte chunks, left subtree = the largest power of two of
 * chunks that leaves at least one for the right) -- ids of any length hash as blake3::Hasher does (mod.rs:347-353) */
static inline int blake3_hash(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t stack[64][8], cv[8];
    int n = 0;
    size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
    for (size_t c = 0; c + 1 < nchunks; c++) {  /* every chunk but the last: push, merging completed subtrees */
        blake3_chunk_cv(cv, in + 1024 * c, 1024, c, 0);
        for (size_t total = c + 1; (total & 1) == 0; total >>= 1) { n--; blake3_parent(cv, stack[n], cv, 0); }
        memcpy(stack[n++], cv, 32);
    }
    blake3_chunk_cv(cv, in + 1024 * (nchunks - 1), len - 1024 * (nchunks - 1), nchunks - 1, n == 0);
    while (n > 0) { n--; blake3_parent(cv, stack[n], cv, n == 0); }
    memcpy(out, cv, 32);
    return 0;
}

/* ---------------------------------------------------------------- BLAKE2s-256 */
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline void blake2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last) {
    uint32_t m[16], v[16];
    memcpy(m, block, 64);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = BLAKE_IV[i]; }
    v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 10; r++) {
        const uint8_t *s = B2S_SIGMA[r];
        BLAKE_G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        BLAKE_G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        BLAKE_G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        BLAKE_G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        BLAKE_G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        BLAKE_G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        BLAKE_G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        BLAKE_G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
static inline int blake2s_hash(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t h[8];
    memcpy(h, BLAKE_IV, 32);
    h[0] ^= 0x01010020;
    size_t off = 0;
    while (len - off > 64) { blake2s_compress(h, in + off, off + 64, 0); off += 64; }
    uint8_t block[64] = {0};
    memcpy(block, in + off, len - off);
    blake2s_compress(h, block, len, 1);
    memcpy(out, h, 32);
    return 0;
}

/* ---------------------------------------------------------------- BLAKE2b-512 (blake2::Blake2b, src/tests.rs:104-105: 64-byte digests) */
static const uint64_t B2B_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static inline uint64_t b2b_rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
#define B2B_G(a, b, c, d, x, y)                      \
    do {                                             \
        a = a + b + (x); d = b2b_rotr(d ^ a, 32);    \
        c = c + d;       b = b2b_rotr(b ^ c, 24);    \
        a = a + b + (y); d = b2b_rotr(d ^ a, 16);    \
        c = c + d;       b = b2b_rotr(b ^ c, 63);    \
    } while (0)
static inline void blake2b_compress(uint64_t h[8], const uint8_t block[128], uint64_t t, int last) {
    uint64_t m[16], v[16];
    memcpy(m, block, 128);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2B_IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2S_SIGMA[r % 10];
        B2B_G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        B2B_G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        B2B_G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        B2B_G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        B2B_G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        B2B_G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        B2B_G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        B2B_G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
static inline int blake2b_hash(const uint8_t *in, size_t len, uint8_t out[64]) {
    uint64_t h[8];
    memcpy(h, B2B_IV, 64);
    h[0] ^= 0x01010040ULL;
    size_t off = 0;
    while (len - off > 128) { blake2b_compress(h, in + off, off + 128, 0); off += 128; }
    uint8_t block[128] = {0};
    memcpy(block, in + off, len - off);
    blake2b_compress(h, block, len, 1);
    memcpy(out, h, 64);
    return 0;
}

/* ---------------------------------------------------------------- Keccak-f[1600] */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline void keccak_f(uint64_t st[25]) {
    uint64_t bc[5], t;
    for (int r = 0; r < 24; r++) {
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        t = st[1];
        for (int i = 0; i < 24; i++) { int j = KECCAK_PILN[i]; bc[0] = st[j]; st[j] = rotl64(t, KECCAK_ROTC[i]); t = bc[0]; }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KECCAK_RC[r];
    }
}
/* sponge: absorb whole message with pad byte `ds`, squeeze outlen */
static inline void keccak_sponge(const uint8_t *in, size_t len, size_t rate, uint8_t ds, uint8_t *out, size_t outlen) {
    uint64_t st[25] = {0};
    uint8_t *sb = (uint8_t *)st;
    size_t pos = 0;
    for (size_t i = 0; i < len; i++) { sb[pos++] ^= in[i]; if (pos == rate) { keccak_f(st); pos = 0; } }
    sb[pos] ^= ds; sb[rate - 1] ^= 0x80;
    keccak_f(st);
    pos = 0;
    for (size_t i = 0; i < outlen; i++) { if (pos == rate) { keccak_f(st); pos = 0; } out[i] = sb[pos++]; }
}
static inline void sha3_512(const uint8_t *in, size_t len, uint8_t out[64]) { keccak_sponge(in, len, 72, 0x06, out, 64); }
static inline void shake256(const uint8_t *in, size_t len, uint8_t *out, size_t outlen) { keccak_sponge(in, len, 136, 0x1F, out, outlen); }

/* ---------------------------------------------------------------- STROBE-128 / merlin (merlin strobe.rs, transcript.rs) */
typedef struct { uint64_t st[25]; uint8_t pos, pos_begin, cur_flags; } strobe;
#define STROBE_R 166
enum { SF_I = 1, SF_A = 2, SF_C = 4, SF_T = 8, SF_M = 16, SF_K = 32 };

static inline void strobe_run_f(strobe *s) {
    uint8_t *b = (uint8_t *)s->st;
    b[s->pos] ^= s->pos_begin; b[s->pos + 1] ^= 0x04; b[STROBE_R + 1] ^= 0x80;
    keccak_f(s->st);
    s->pos = 0; s->pos_begin = 0;
}
static inline void strobe_absorb(strobe *s, const uint8_t *d, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_squeeze(strobe *s, uint8_t *d, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { d[i] = b[s->pos]; b[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_begin_op(strobe *s, uint8_t flags, int more) {
    if (more) return;
    uint8_t hdr[2] = {s->pos_begin, flags};
    s->pos_begin = s->pos + 1; s->cur_flags = flags;
    strobe_absorb(s, hdr, 2);
    if ((flags & (SF_C | SF_K)) && s->pos != 0) strobe_run_f(s);
}
static inline void strobe_meta_ad(strobe *s, const void *d, size_t n, int more) { strobe_begin_op(s, SF_M | SF_A, more); strobe_absorb(s, (const uint8_t *)d, n); }
static inline void strobe_ad(strobe *s, const void *d, size_t n, int more) { strobe_begin_op(s, SF_A, more); strobe_absorb(s, (const uint8_t *)d, n); }
static inline void strobe_prf(strobe *s, uint8_t *d, size_t n) { strobe_begin_op(s, SF_I | SF_A | SF_C, 0); strobe_squeeze(s, d, n); }
static inline void strobe_init(strobe *s, const char *label) {
    memset(s, 0, sizeof *s);
    uint8_t *b = (uint8_t *)s->st;
    b[0] = 1; b[1] = STROBE_R + 2; b[2] = 1; b[3] = 0; b[4] = 1; b[5] = 96;
    memcpy(b + 6, "STROBEv1.0.2", 12);
    keccak_f(s->st);
    strobe_meta_ad(s, label, strlen(label), 0);
}
typedef strobe transcript;
static inline void tr_append(transcript *t, const char *label, const void *msg, uint32_t len) {
    strobe_meta_ad(t, label, strlen(label), 0);
    strobe_meta_ad(t, &len, 4, 1);
    strobe_ad(t, msg, len, 0);
}
static inline void tr_append_u64(transcript *t, const char *label, uint64_t x) { tr_append(t, label, &x, 8); }
static inline void tr_challenge(transcript *t, const char *label, uint8_t *out, uint32_t n) {
    strobe_meta_ad(t, label, strlen(label), 0);
    strobe_meta_ad(t, &n, 4, 1);
    strobe_prf(t, out, n);
}
static inline void tr_init(transcript *t, const void *label, uint32_t len) { /* Transcript::new(label) */
    strobe_init(t, "Merlin v1.0");
    tr_append(t, "dom-sep", label, len);
}

/* ---------------------------------------------------------------- ChaCha20 block (rand_chacha ChaCha20Rng layout) */
static inline void chacha20_block(const uint8_t key[32], uint64_t counter, uint64_t stream, uint8_t out[64]) {
    uint32_t st[16], w[16];
    st[0] = 0x61707865; st[1] = 0x3320646e; st[2] = 0x79622d32; st[3] = 0x6b206574;
    memcpy(st + 4, key, 32);
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32);
    st[14] = (uint32_t)stream; st[15] = (uint32_t)(stream >> 32);
    memcpy(w, st, 64);
#define CHACHA_QR(a, b, c, d)                                        \
    w[a] += w[b]; w[d] = rotl32(w[d] ^ w[a], 16); w[c] += w[d]; w[b] = rotl32(w[b] ^ w[c], 12); \
    w[a] += w[b]; w[d] = rotl32(w[d] ^ w[a], 8);  w[c] += w[d]; w[b] = rotl32(w[b] ^ w[c], 7);
    for (int i = 0; i < 10; i++) {
        CHACHA_QR(0, 4, 8, 12) CHACHA_QR(1, 5, 9, 13) CHACHA_QR(2, 6, 10, 14) CHACHA_QR(3, 7, 11, 15)
        CHACHA_QR(0, 5, 10, 15) CHACHA_QR(1, 6, 11, 12) CHACHA_QR(2, 7, 8, 13) CHACHA_QR(3, 4, 9, 14)
    }
    for (int i = 0; i < 16; i++) w[i] += st[i];
    memcpy(out, w, 64);
}
#endif
