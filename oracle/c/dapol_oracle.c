/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the DAPOL+ hot path, plain C restatement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (dapol_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" against the Rust reference for commitment/hash/root/proof
 * bytes (no golden vectors exist there and it cannot be built here -- SURVEY.md F2-F5); pinned
 * against the reference's id->index KAT, root value 26, 672-byte single proof, and against
 * RFC 9496 / bulletproofs / merlin published vectors (tests/test_oracle_pins.py), and
 * cross-checked against the independent big-int restatement oracle/pyref.py.
 *
 * Reference lines followed (under /root/reference):
 *   src/dapol/mod.rs:323-441   leaf derivation, shuffle_index
 *   src/dapol/node.rs:29-89    DapolNode::new / merge / padding
 *   src/dapol/mod.rs:172-190   generate_proof_batch (single leaf)
 *   src/proof/mod.rs:41-95, src/proof/node.rs:51-102   verification + node wire format
 *   src/range/mod.rs:16-161, padding.rs:40-197, splitting.rs:38-211
 * Upstream algorithms restated: smtree build, bulletproofs range_proof + inner_product_proof,
 * merlin, dalek. */
#include <stdio.h>
#include <stdlib.h>
#include "ec.h"
#include "hashes.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

enum { HASH_BLAKE3 = 0, HASH_BLAKE2S = 1, HASH_BLAKE2B = 2 };
/* digest length of D: 32, or 64 for blake2::Blake2b (src/tests.rs:104-105: the 64-byte-digest half of the reference's test matrix,
 * reachable through new_blank + build only, since Dapol::new insists on 32 bytes, mod.rs:101-103) */
#define DLEN(hash_id) ((hash_id) == HASH_BLAKE2B ? 64 : 32)
enum { ERR_OK = 0, ERR_TREE_HEIGHT_TOO_BIG = 1, ERR_SPARSITY_TOO_SMALL = 2, ERR_INVALID_DIGEST_SIZE = 3,
       ERR_DUPLICATED_INTERNAL_ID = 4, ERR_FAILED_TO_MAP_INDEX = 5, ERR_BAD_ARG = 16, ERR_NOT_FOUND = 17, ERR_BUFFER = 18 };

static int g_init = 0;

EXPORT void dor_init(void) {
    if (g_init) return;
    fe_from_hex_le(&FE_D, "a3785913ca4deb75abd841414d0a700098e879777940c78c73fe6f2bee6c0352");
    fe_add(&FE_D2, &FE_D, &FE_D);
    fe_from_hex_le(&FE_SQRT_M1, "b0a00e4a271beec478e42fad0618432fa7d7fb3d99004d2b0bdfc14f8024832b");
    /* remaining constants derived, then checked in tests against RFC 9496 values */
    fe one, t, t2;
    fe_1(&one);
    fe_sq(&t, &FE_D); fe_sub(&FE_ONE_MINUS_D_SQ, &one, &t);
    fe_sub(&t, &FE_D, &one); fe_sq(&FE_D_MINUS_ONE_SQ, &t);
    /* a - d = -1 - d ; invsqrt */
    fe_neg(&t, &one); fe_sub(&t, &t, &FE_D);
    fe_sqrt_ratio_m1(&FE_INVSQRT_A_MINUS_D, &one, &t);
    /* sqrt(a*d - 1) = sqrt(-d - 1): RFC representative is the ODD root */
    fe_sqrt_ratio_m1(&t2, &t, &one);
    if (!fe_isneg(&t2)) fe_neg(&t2, &t2);
    FE_SQRT_AD_MINUS_ONE = t2;
    /* scalar Montgomery constants */
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - SC_L[0] * inv;
    SC_N0 = (uint64_t)0 - inv;
    sc x; sc_from_u64(&x, 1);
    for (int i = 0; i < 256; i++) sc_add(&x, &x, &x);
    SC_R1 = x;
    for (int i = 0; i < 256; i++) sc_add(&x, &x, &x);
    SC_RR = x;
    /* basepoint: decompress the canonical encoding, then fix to the standard representative */
    static const uint8_t BC[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                   0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    ge_decompress(&GE_B, BC);
    uint8_t h[64];
    sha3_512(BC, 32, h);
    ge_from_uniform(&GE_BBL, h);
    ge_table8(TAB_B[0], &GE_B);
    ge_table8(TAB_B[1], &GE_BBL);
    g_init = 1;
}

static int hash_buf(int hash_id, const uint8_t *in, size_t len, uint8_t *out /* DLEN(hash_id) bytes */) {
    if (hash_id == HASH_BLAKE3) return blake3_hash(in, len, out);
    if (hash_id == HASH_BLAKE2S) return blake2s_hash(in, len, out);
    if (hash_id == HASH_BLAKE2B) return blake2b_hash(in, len, out);
    return -1;
}
EXPORT int dor_hash(int hash_id, const uint8_t *in, size_t len, uint8_t *out) { return hash_buf(hash_id, in, len, out); }

/* PedersenGens::commit via 2-base Straus, as dalek's constant-time multiscalar_mul does */
static void commit_ge(ge *out, uint64_t v, const sc *r) {
    sc s[2];
    sc_from_u64(&s[0], v); s[1] = *r;
    ge_straus(out, s, (const ge(*)[8])TAB_B, 2);
}
EXPORT void dor_commit(uint64_t v, const uint8_t r[32], uint8_t comc[32]) {
    sc rs; ge p;
    sc_frombytes_reduce(&rs, r);
    commit_ge(&p, v, &rs);
    ge_compress(comc, &p);
}
EXPORT void dor_scalarmult_base(const uint8_t s[32], uint8_t out[32]) {
    sc k; ge p; sc_frombytes_reduce(&k, s);
    ge_scalarmult(&p, &k, &GE_B);
    ge_compress(out, &p);
}
EXPORT int dor_decompress_recompress(const uint8_t s[32], uint8_t out[32]) {
    ge p; if (!ge_decompress(&p, s)) return 0;
    ge_compress(out, &p); return 1;
}
EXPORT void dor_from_uniform(const uint8_t b[64], uint8_t out[32]) { ge p; ge_from_uniform(&p, b); ge_compress(out, &p); }
EXPORT void dor_point_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    ge p, q; ge_decompress(&p, a); ge_decompress(&q, b); ge_add(&p, &p, &q); ge_compress(out, &p);
}
EXPORT void dor_get_constant(int which, uint8_t out[32]) {
    const fe *c[] = {&FE_D, &FE_SQRT_M1, &FE_SQRT_AD_MINUS_ONE, &FE_INVSQRT_A_MINUS_D, &FE_ONE_MINUS_D_SQ, &FE_D_MINUS_ONE_SQ};
    if (which == 100) { ge_compress(out, &GE_BBL); return; }
    fe_tobytes(out, c[which]);
}
static void rng_scalar(sc *out, const uint8_t seed[32], uint64_t k, uint64_t stream) {
    uint8_t b[64]; chacha20_block(seed, k, stream, b); sc_from_wide(out, b);
}
EXPORT void dor_rng_scalar(const uint8_t seed[32], uint64_t k, uint64_t stream, uint8_t out[32]) {
    sc s; rng_scalar(&s, seed, k, stream); sc_tobytes(out, &s);
}
EXPORT void dor_sc_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc_mul(&x, &x, &y); sc_tobytes(out, &x);
}
EXPORT void dor_sc_invert(const uint8_t a[32], uint8_t out[32]) { sc x; sc_frombytes_reduce(&x, a); sc_invert(&x, &x); sc_tobytes(out, &x); }
EXPORT void dor_merlin_test(const uint8_t *label, uint32_t llen, const char *mlabel, const uint8_t *msg, uint32_t mlen,
                            const char *clabel, uint8_t *out, uint32_t outlen) {
    transcript t; tr_init(&t, label, llen); tr_append(&t, mlabel, msg, mlen); tr_challenge(&t, clabel, out, outlen);
}

/* ------------------------------------------------------------------ Bulletproof generators */
static ge *g_G = NULL, *g_H = NULL; /* [party][64] */
static int g_gens_m = 0;
static void ensure_gens(int m) { /* BulletproofGens::new(64, m): SHAKE256("GeneratorsChain"||tag||le32(j)) */
    if (m <= g_gens_m) return;
#pragma omp critical(dor_gens)
    if (m > g_gens_m) {
        ge *G = (ge *)malloc(sizeof(ge) * 64 * (size_t)m), *H = (ge *)malloc(sizeof(ge) * 64 * (size_t)m);
        if (g_gens_m) { memcpy(G, g_G, sizeof(ge) * 64 * g_gens_m); memcpy(H, g_H, sizeof(ge) * 64 * g_gens_m); }
        for (int j = g_gens_m; j < m; j++)
            for (int tag = 0; tag < 2; tag++) {
                uint8_t label[20], stream[64 * 64];
                memcpy(label, "GeneratorsChain", 15);
                label[15] = tag ? 'H' : 'G';
                uint32_t jj = (uint32_t)j; memcpy(label + 16, &jj, 4);
                shake256(label, 20, stream, sizeof stream);
                for (int i = 0; i < 64; i++) ge_from_uniform(&(tag ? H : G)[j * 64 + i], stream + 64 * i);
            }
        ge *oG = g_G, *oH = g_H;
        g_G = G; g_H = H; g_gens_m = m;
        (void)oG; (void)oH; /* old tables intentionally leaked: other threads may still read them */
    }
}
EXPORT void dor_bp_gen(int is_h, uint32_t party, uint32_t i, uint8_t out[32]) {
    ensure_gens((int)party + 1);
    ge_compress(out, &(is_h ? g_H : g_G)[party * 64 + i]);
}

/* ------------------------------------------------------------------ leaf derivation (mod.rs:323-441) */
typedef struct { uint64_t key; uint64_t val; } kv;
static int kv_cmp(const void *a, const void *b) { uint64_t x = ((const kv *)a)->key, y = ((const kv *)b)->key; return x < y ? -1 : x > y; }

/* open-addressing set of u64 (tree_index_set) */
typedef struct { uint64_t *slot; uint8_t *used; uint64_t mask; } u64set;
static void set_init(u64set *s, uint64_t n) { uint64_t c = 16; while (c < 2 * n + 8) c <<= 1; s->slot = calloc(c, 8); s->used = calloc(c, 1); s->mask = c - 1; }
static int set_insert(u64set *s, uint64_t k) {
    uint64_t h = (k * 0x9E3779B97F4A7C15ULL) >> 7;
    for (;; h++) { h &= s->mask; if (!s->used[h]) { s->used[h] = 1; s->slot[h] = k; return 1; } if (s->slot[h] == k) return 0; }
}
static void set_free(u64set *s) { free(s->slot); free(s->used); }

EXPORT int dor_derive_leaves(int hash_id, uint64_t n, const uint8_t *iid_blob, const uint64_t *iid_off,
                             const uint8_t *eid_blob, const uint64_t *eid_off, const uint8_t *audit_seed, uint64_t seed_len,
                             int height, uint64_t *out_idx, uint8_t *out_blind, uint64_t *err_pos) {
    if (DLEN(hash_id) != 32) return ERR_INVALID_DIGEST_SIZE; /* mod.rs:101-103 */
    if (height > 64) return ERR_TREE_HEIGHT_TOO_BIG;
    if (height < 64 && ((uint64_t)1 << height) < n * 2) return ERR_SPARSITY_TOO_SMALL;
    if (height == 0 && n) return ERR_BAD_ARG;
    /* duplicate internal ids: detect via audit_id equality (same id <=> same audit_id, mod 2^-256) */
    uint8_t *audit = (uint8_t *)malloc(32 * (n ? n : 1));
    /* ids of any length (the reference hashes whatever it is given, mod.rs:347-384): one buffer sized for the longest input */
    uint64_t longest = 64;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t il = iid_off[i + 1] - iid_off[i], el = eid_off[i + 1] - eid_off[i];
        if (seed_len + il > longest) longest = seed_len + il;
        if (42 + el > longest) longest = 42 + el;
    }
    uint8_t *buf = (uint8_t *)malloc(longest);
    int rc = ERR_OK;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t il = iid_off[i + 1] - iid_off[i];
        memcpy(buf, audit_seed, seed_len); memcpy(buf + seed_len, iid_blob + iid_off[i], il);
        hash_buf(hash_id, buf, seed_len + il, audit + 32 * i);
    }
    kv *dup = (kv *)malloc(sizeof(kv) * (n ? n : 1));
    for (uint64_t i = 0; i < n; i++) { memcpy(&dup[i].key, audit + 32 * i, 8); dup[i].val = i; }
    qsort(dup, n, sizeof(kv), kv_cmp);
    uint64_t first_dup = UINT64_MAX;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && dup[j].key == dup[i].key) j++;
        for (uint64_t a = i; a < j; a++)
            for (uint64_t b = a + 1; b < j; b++)
                if (memcmp(audit + 32 * dup[a].val, audit + 32 * dup[b].val, 32) == 0) {
                    uint64_t later = dup[a].val > dup[b].val ? dup[a].val : dup[b].val;
                    if (later < first_dup) first_dup = later;
                }
        i = j;
    }
    free(dup);
    u64set used; set_init(&used, n);
    for (uint64_t i = 0; i < n && rc == ERR_OK; i++) {
        if (i == first_dup) { rc = ERR_DUPLICATED_INTERNAL_ID; if (err_pos) *err_pos = i; break; }
        uint64_t el = eid_off[i + 1] - eid_off[i];
        uint8_t seed[32];
        memcpy(buf, audit + 32 * i, 32); memcpy(buf + 32, "index_seed", 10); memcpy(buf + 42, eid_blob + eid_off[i], el);
        hash_buf(hash_id, buf, 42 + el, seed);
        int found = 0;
        for (int t = 0; t < 128; t++) {
            hash_buf(hash_id, seed, 32, seed);
            uint64_t be = 0;
            for (int k = 0; k < 8; k++) be = (be << 8) | seed[k];
            uint64_t cand = be >> (64 - height);
            if (set_insert(&used, cand)) { out_idx[i] = cand; found = 1; break; }
        }
        if (!found) { rc = ERR_FAILED_TO_MAP_INDEX; if (err_pos) *err_pos = i; break; }
        memcpy(buf + 32, "blind_seed", 10);
        hash_buf(hash_id, buf, 42 + el, out_blind + 32 * i);
        out_blind[32 * i + 31] &= 0x7f; /* Scalar::from_bits */
    }
    set_free(&used);
    free(audit); free(buf);
    return rc;
}

/* ------------------------------------------------------------------ tree (smtree build restated) */
typedef struct {
    uint64_t idx, v;
    uint8_t r[32];    /* blinding bytes as the reference holds them (leaf: from_bits, maybe unreduced) */
    uint8_t comc[32]; /* compress(com) */
    uint8_t hash[64]; /* DLEN(hash_id) bytes used */
    uint8_t is_pad;
    ge ext;
} node;
typedef struct { node *nodes; uint64_t count; } level;
typedef struct dor_tree { int hash_id, height; level *lv; /* lv[h], h = 0 root .. height leaves */ uint64_t n_pads; } dor_tree;

static void node_finish(int hash_id, node *nd) { /* DapolNode::new hash part: hash = D(compress(com)) (node.rs:33-36) */
    ge_compress(nd->comc, &nd->ext);
    hash_buf(hash_id, nd->comc, 32, nd->hash);
}
static void node_new(int hash_id, node *nd, uint64_t idx, uint64_t v, const uint8_t r[32], int is_pad) {
    sc rs;
    nd->idx = idx; nd->v = v; nd->is_pad = (uint8_t)is_pad;
    memcpy(nd->r, r, 32);
    sc_frombytes_reduce(&rs, r);
    commit_ge(&nd->ext, v, &rs);
    node_finish(hash_id, nd);
}
static void node_merge(int hash_id, node *p, const node *l, const node *r) { /* node.rs:64-80 */
    uint8_t buf[192];
    const int dl = DLEN(hash_id);
    memcpy(buf, l->comc, 32); memcpy(buf + 32, r->comc, 32); memcpy(buf + 64, l->hash, dl); memcpy(buf + 64 + dl, r->hash, dl);
    hash_buf(hash_id, buf, 64 + 2 * dl, p->hash);
    p->idx = l->idx >> 1; p->v = l->v + r->v; p->is_pad = 0;
    sc a, b; sc_frombytes_reduce(&a, l->r); sc_frombytes_reduce(&b, r->r); sc_add(&a, &a, &b); sc_tobytes(p->r, &a);
    ge_add(&p->ext, &l->ext, &r->ext);
    ge_compress(p->comc, &p->ext);
}

EXPORT void dor_tree_free(dor_tree *t) {
    if (!t) return;
    for (int h = 0; h <= t->height; h++) free(t->lv[h].nodes);
    free(t->lv); free(t);
}

/* Dapol::new_blank + build (mod.rs:196-208) with padding draws from ChaCha20(pad_seed): draw order =
 * creation order, level H..1, left to right, starting at block pad_base.
 * Level layout: lv[h].nodes[2j], [2j+1] are the left/right children of lv[h-1] parent j. */
#define NO_DRAW UINT64_MAX
static int is_pair(const node *cur, uint64_t cnt, uint64_t i) {
    return !(cur[i].idx & 1) && i + 1 < cnt && cur[i + 1].idx == cur[i].idx + 1;
}
/* pad_mode 0: the reference's behaviour under the seeded-RNG contract (above).  pad_mode 1 (opt-in): the
 * padding node at (level h, index i) draws block i of stream h of ChaCha20(pad_seed) -- Paddable::padding(idx, secret) as a
 * function of its arguments (src/dapol/node.rs:85-88 leaves that as a TODO); pad_base is ignored. */
static int tree_build_mode(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                           const uint8_t *blindings, const uint8_t pad_seed[32], uint64_t pad_base, int nthreads, int pad_mode,
                           const uint64_t *level_base, dor_tree **out);
/* leaf hashes given by the caller instead of D(compress(com)): the opt-in id / salt leaf hash (include/dapol_b200.h,
 * DAPOL_LEAF_HASH_ID_SALT); NULL = the reference's rule (node.rs:33-36) */
static const uint8_t *g_leaf_hash_override = NULL;
EXPORT int dor_tree_build_leaf_hashes(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                                      const uint8_t *blindings, const uint8_t *leaf_hashes, const uint8_t pad_seed[32], uint64_t pad_base,
                                      int nthreads, dor_tree **out) {
    g_leaf_hash_override = leaf_hashes;  /* test infrastructure: single-threaded callers */
    int rc = tree_build_mode(hash_id, height, n, idx_sorted, values, blindings, pad_seed, pad_base, nthreads, 0, NULL, out);
    g_leaf_hash_override = NULL;
    return rc;
}
/* salt = D(audit_id || "salt_seed" || external_id), leaf hash = D("leaf" || external_id || salt), audit_id = D(audit_seed || internal_id) */
EXPORT void dor_leaf_id_hashes(int hash_id, uint64_t n, const uint8_t *iid_blob, const uint64_t *iid_off, const uint8_t *eid_blob,
                               const uint64_t *eid_off, const uint8_t *audit_seed, uint64_t seed_len, uint8_t *out) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t il = iid_off[i + 1] - iid_off[i], el = eid_off[i + 1] - eid_off[i];
        uint8_t *buf = (uint8_t *)malloc(64 + seed_len + il + el), audit[32], salt[32];
        memcpy(buf, audit_seed, seed_len); memcpy(buf + seed_len, iid_blob + iid_off[i], il);
        hash_buf(hash_id, buf, seed_len + il, audit);
        memcpy(buf, audit, 32); memcpy(buf + 32, "salt_seed", 9); memcpy(buf + 41, eid_blob + eid_off[i], el);
        hash_buf(hash_id, buf, 41 + el, salt);
        memcpy(buf, "leaf", 4); memcpy(buf + 4, eid_blob + eid_off[i], el); memcpy(buf + 4 + el, salt, 32);
        hash_buf(hash_id, buf, 36 + el, out + 32 * i);
        free(buf);
    }
}
EXPORT int dor_tree_build(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                          const uint8_t *blindings, const uint8_t pad_seed[32], uint64_t pad_base, int nthreads, dor_tree **out) {
    return tree_build_mode(hash_id, height, n, idx_sorted, values, blindings, pad_seed, pad_base, nthreads, 0, NULL, out);
}
/* One shard of a tree split by leaf-index prefix: the subtree under one node of level k draws its padding
 * blindings from the blocks the single-tree creation order (level H..1, left to right) gives them: the r-th padding node of
 * the subtree's level h draws block level_base[h] + r of stream 0. */
EXPORT int dor_tree_build_shard(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                                const uint8_t *blindings, const uint8_t pad_seed[32], const uint64_t *level_base /*[height+1]*/,
                                int nthreads, dor_tree **out) {
    return tree_build_mode(hash_id, height, n, idx_sorted, values, blindings, pad_seed, 0, nthreads, 0, level_base, out);
}
EXPORT int dor_tree_build_positional(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                                     const uint8_t *blindings, const uint8_t pad_key[32], int nthreads, dor_tree **out) {
    return tree_build_mode(hash_id, height, n, idx_sorted, values, blindings, pad_key, 0, nthreads, 1, NULL, out);
}
static int tree_build_mode(int hash_id, int height, uint64_t n, const uint64_t *idx_sorted, const uint64_t *values,
                           const uint8_t *blindings, const uint8_t pad_seed[32], uint64_t pad_base, int nthreads, int pad_mode,
                           const uint64_t *level_base, dor_tree **out) {
    if (height > 64 || height < 0 || n == 0 || (height == 0 && n != 1)) return ERR_BAD_ARG;
    for (uint64_t i = 0; i < n; i++) {
        if (i && idx_sorted[i] <= idx_sorted[i - 1]) return ERR_BAD_ARG;
        if (height < 64 && (idx_sorted[i] >> height)) return ERR_BAD_ARG;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    dor_tree *t = (dor_tree *)calloc(1, sizeof *t);
    t->hash_id = hash_id; t->height = height;
    t->lv = (level *)calloc((size_t)height + 1, sizeof(level));
    node *cur = (node *)calloc(n, sizeof(node));
    uint64_t cnt = n, draw = pad_base, n_pads = 0;
#pragma omp parallel for schedule(dynamic, 16)
    for (uint64_t i = 0; i < n; i++) {
        node_new(hash_id, &cur[i], idx_sorted[i], values[i], blindings + 32 * i, 0);
        if (g_leaf_hash_override) memcpy(cur[i].hash, g_leaf_hash_override + DLEN(hash_id) * i, DLEN(hash_id));
    }
    for (int h = height; h >= 1; h--) {
        uint64_t np = 0;
        for (uint64_t i = 0; i < cnt; np++) i += is_pair(cur, cnt, i) ? 2 : 1;
        node *full = (node *)calloc(2 * np, sizeof(node));
        uint64_t *pad_draw = (uint64_t *)malloc(8 * np);
        uint64_t j = 0;
        if (level_base) draw = level_base[h];
        for (uint64_t i = 0; i < cnt; j++) {
            if (is_pair(cur, cnt, i)) {
                full[2 * j] = cur[i]; full[2 * j + 1] = cur[i + 1]; pad_draw[j] = NO_DRAW; i += 2;
            } else {
                uint64_t side = cur[i].idx & 1;
                full[2 * j + side] = cur[i];
                full[2 * j + (1 - side)].idx = cur[i].idx ^ 1;
                full[2 * j + (1 - side)].is_pad = 1;
                pad_draw[j] = draw++; n_pads++;
                i += 1;
            }
        }
        free(cur);
        node *par = (node *)calloc(np, sizeof(node));
#pragma omp parallel for schedule(dynamic, 16)
        for (uint64_t k = 0; k < np; k++) {
            if (pad_draw[k] != NO_DRAW) { /* DapolNode::padding (node.rs:86-88): new(0, Scalar::random(rng)) */
                node *pd = full[2 * k].is_pad ? &full[2 * k] : &full[2 * k + 1];
                sc r; uint8_t rb[32];
                if (pad_mode == 1) rng_scalar(&r, pad_seed, pd->idx, (uint64_t)h);
                else rng_scalar(&r, pad_seed, pad_draw[k], 0);
                sc_tobytes(rb, &r);
                node_new(hash_id, pd, pd->idx, 0, rb, 1);
            }
            node_merge(hash_id, &par[k], &full[2 * k], &full[2 * k + 1]);
        }
        free(pad_draw);
        t->lv[h].nodes = full; t->lv[h].count = 2 * np;
        cur = par; cnt = np;
    }
    t->lv[0].nodes = cur; t->lv[0].count = cnt;
    t->n_pads = n_pads;
    *out = t;
    return ERR_OK;
}
EXPORT uint64_t dor_tree_level_size(const dor_tree *t, int h) { return t->lv[h].count; }
EXPORT uint64_t dor_tree_num_pads(const dor_tree *t) { return t->n_pads; }
EXPORT void dor_tree_level_copy(const dor_tree *t, int h, uint64_t *idx, uint64_t *v, uint8_t *r, uint8_t *comc, uint8_t *hash, uint8_t *is_pad) {
    const level *l = &t->lv[h];
    for (uint64_t i = 0; i < l->count; i++) {
        const node *nd = &l->nodes[i];
        if (idx) idx[i] = nd->idx;
        if (v) v[i] = nd->v;
        if (r) memcpy(r + 32 * i, nd->r, 32);
        if (comc) memcpy(comc + 32 * i, nd->comc, 32);
        if (hash) memcpy(hash + DLEN(t->hash_id) * i, nd->hash, DLEN(t->hash_id));
        if (is_pad) is_pad[i] = nd->is_pad;
    }
}
static const node *tree_find(const dor_tree *t, int h, uint64_t idx) {
    const level *l = &t->lv[h];
    uint64_t lo = 0, hi = l->count;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (l->nodes[mid].idx < idx) lo = mid + 1; else hi = mid; }
    return (lo < l->count && l->nodes[lo].idx == idx) ? &l->nodes[lo] : NULL;
}
/* siblings of a leaf's path, leaf level first, root child last (smtree get_merkle_path_ref_batch, single leaf) */
static int tree_path(const dor_tree *t, uint64_t leaf_idx, const node **sib /*[height]*/, const node **leaf) {
    const node *lf = tree_find(t, t->height, leaf_idx);
    if (!lf || lf->is_pad) return ERR_NOT_FOUND;
    if (leaf) *leaf = lf;
    for (int h = t->height, k = 0; h >= 1; h--, k++) {
        uint64_t at = t->height - h >= 64 ? 0 : leaf_idx >> (t->height - h);
        sib[k] = tree_find(t, h, at ^ 1);
        if (!sib[k]) return ERR_NOT_FOUND;
    }
    return ERR_OK;
}
EXPORT int dor_tree_path(const dor_tree *t, uint64_t leaf_idx, uint64_t *v, uint8_t *r, uint8_t *comc, uint8_t *hash) {
    const node *sib[64];
    int rc = tree_path(t, leaf_idx, sib, NULL);
    if (rc) return rc;
    for (int k = 0; k < t->height; k++) {
        v[k] = sib[k]->v; memcpy(r + 32 * k, sib[k]->r, 32); memcpy(comc + 32 * k, sib[k]->comc, 32); memcpy(hash + DLEN(t->hash_id) * k, sib[k]->hash, DLEN(t->hash_id));
    }
    return ERR_OK;
}
EXPORT int dor_tree_get_node(const dor_tree *t, int h, uint64_t idx, uint64_t *v, uint8_t r[32], uint8_t comc[32], uint8_t *hash, uint8_t *is_pad) {
    const node *nd = tree_find(t, h, idx);
    if (!nd) return ERR_NOT_FOUND;
    *v = nd->v; memcpy(r, nd->r, 32); memcpy(comc, nd->comc, 32); memcpy(hash, nd->hash, DLEN(t->hash_id)); *is_pad = nd->is_pad;
    return ERR_OK;
}

/* ------------------------------------------------------------------ Bulletproofs (bulletproofs ^4.0.0) */
static void tr_append_sc(transcript *t, const char *l, const sc *s) { uint8_t b[32]; sc_tobytes(b, s); tr_append(t, l, b, 32); }
static void tr_challenge_sc(transcript *t, const char *l, sc *out) { uint8_t b[64]; tr_challenge(t, l, b, 64); sc_from_wide(out, b); }
static int is_zero32(const uint8_t *b) { uint8_t r = 0; for (int i = 0; i < 32; i++) r |= b[i]; return r == 0; }
static void sc_inner(sc *out, const sc *a, const sc *b, size_t n) {
    sc acc, t; sc_from_u64(&acc, 0);
    for (size_t i = 0; i < n; i++) { sc_mul(&t, &a[i], &b[i]); sc_add(&acc, &acc, &t); }
    *out = acc;
}
static void sc_pow(sc *out, const sc *b, uint64_t e) {
    sc acc, x = *b; sc_from_u64(&acc, 1);
    while (e) { if (e & 1) sc_mul(&acc, &acc, &x); sc_mul(&x, &x, &x); e >>= 1; }
    *out = acc;
}

/* RangeProof::prove_multiple_with_rng over Transcript::new(&[]) (src/range/mod.rs:48-78).
 * RNG: the k-th Scalar::random = wide-reduce(ChaCha20(seed, stream) block base+k). */
EXPORT int dor_rp_prove(int nbits, int m, const uint64_t *values, const uint8_t *blindings, const uint8_t seed[32],
                        uint64_t stream, uint64_t base_block, uint8_t *out, uint64_t *out_len) {
    if (!(nbits == 8 || nbits == 16 || nbits == 32 || nbits == 64) || m <= 0 || (m & (m - 1))) return ERR_BAD_ARG;
    ensure_gens(m);
    const ge *G = g_G, *H = g_H; /* party j generator i at [j*64 + i] */
    size_t n = (size_t)nbits, N = n * (size_t)m;
    uint64_t draw = base_block;
    transcript tr; tr_init(&tr, "", 0);
    tr_append(&tr, "dom-sep", "rangeproof v1", 13);
    tr_append_u64(&tr, "n", (uint64_t)n); tr_append_u64(&tr, "m", (uint64_t)m);
    sc *bl = malloc(sizeof(sc) * m), *a_bl = malloc(sizeof(sc) * m), *s_bl = malloc(sizeof(sc) * m);
    sc *sL = malloc(sizeof(sc) * N), *sR = malloc(sizeof(sc) * N);
    sc *l0 = malloc(sizeof(sc) * N), *r0 = malloc(sizeof(sc) * N), *r1 = malloc(sizeof(sc) * N);
    sc *t1b = malloc(sizeof(sc) * m), *t2b = malloc(sizeof(sc) * m);
    ge *Gv = malloc(sizeof(ge) * N), *Hv = malloc(sizeof(ge) * N);
    ge A, S, T1, T2; ge_identity(&A); ge_identity(&S); ge_identity(&T1); ge_identity(&T2);
    uint8_t buf[32];
    for (int j = 0; j < m; j++) { /* Party::assign_position_with_rng */
        sc_frombytes_reduce(&bl[j], blindings + 32 * j);
        ge V; commit_ge(&V, values[j], &bl[j]); ge_compress(buf, &V);
        tr_append(&tr, "V", buf, 32);
        rng_scalar(&a_bl[j], seed, draw++, stream);
        ge Aj; ge_scalarmult(&Aj, &a_bl[j], &GE_BBL);
        for (size_t i = 0; i < n; i++) {
            if ((values[j] >> i) & 1) ge_add(&Aj, &Aj, &G[j * 64 + i]); else ge_sub(&Aj, &Aj, &H[j * 64 + i]);
            Gv[j * n + i] = G[j * 64 + i]; Hv[j * n + i] = H[j * 64 + i];
        }
        rng_scalar(&s_bl[j], seed, draw++, stream);
        for (size_t i = 0; i < n; i++) rng_scalar(&sL[j * n + i], seed, draw++, stream);
        for (size_t i = 0; i < n; i++) rng_scalar(&sR[j * n + i], seed, draw++, stream);
        ge Sj, tmp; ge_scalarmult(&Sj, &s_bl[j], &GE_BBL);
        ge_msm(&tmp, &sL[j * n], &Gv[j * n], n); ge_add(&Sj, &Sj, &tmp);
        ge_msm(&tmp, &sR[j * n], &Hv[j * n], n); ge_add(&Sj, &Sj, &tmp);
        ge_add(&A, &A, &Aj); ge_add(&S, &S, &Sj);
    }
    uint8_t *o = out;
    ge_compress(o, &A); tr_append(&tr, "A", o, 32); o += 32;
    ge_compress(o, &S); tr_append(&tr, "S", o, 32); o += 32;
    sc y, z, zz, x, w, one, t0s, t1s, t2s;
    sc_from_u64(&one, 1);
    tr_challenge_sc(&tr, "y", &y); tr_challenge_sc(&tr, "z", &z);
    sc_mul(&zz, &z, &z);
    sc_from_u64(&t0s, 0); t1s = t0s; t2s = t0s;
    for (int j = 0; j < m; j++) { /* Party::apply_challenge_with_rng */
        sc offset_zz, exp_y, exp_2, zj;
        sc_pow(&zj, &z, (uint64_t)j); sc_mul(&offset_zz, &zz, &zj);
        sc_pow(&exp_y, &y, (uint64_t)j * n);
        exp_2 = one;
        sc t0, t1, t2, acc1, tmp, tmp2;
        sc_from_u64(&t0, 0); t2 = t0; acc1 = t0;
        for (size_t i = 0; i < n; i++) {
            size_t k = j * n + i;
            sc aL, aR;
            sc_from_u64(&aL, (values[j] >> i) & 1);
            sc_sub(&aR, &aL, &one);
            sc_sub(&l0[k], &aL, &z);                                   /* l0 = a_L - z ; l1 = s_L */
            sc_add(&tmp, &aR, &z); sc_mul(&tmp, &exp_y, &tmp);
            sc_mul(&tmp2, &offset_zz, &exp_2); sc_add(&r0[k], &tmp, &tmp2); /* r0 = y^k (a_R + z) + z^(2+j) 2^i */
            sc_mul(&r1[k], &exp_y, &sR[k]);                            /* r1 = y^k s_R */
            sc_mul(&tmp, &l0[k], &r0[k]); sc_add(&t0, &t0, &tmp);
            sc_mul(&tmp, &sL[k], &r1[k]); sc_add(&t2, &t2, &tmp);
            sc_add(&tmp, &l0[k], &sL[k]); sc_add(&tmp2, &r0[k], &r1[k]); sc_mul(&tmp, &tmp, &tmp2); sc_add(&acc1, &acc1, &tmp);
            sc_mul(&exp_y, &exp_y, &y); sc_add(&exp_2, &exp_2, &exp_2);
        }
        sc_sub(&t1, &acc1, &t0); sc_sub(&t1, &t1, &t2);
        rng_scalar(&t1b[j], seed, draw++, stream);
        rng_scalar(&t2b[j], seed, draw++, stream);
        /* pc_gens.commit(t1, t1_blinding): scalars are full-width here -> 2-base Straus with Scalar v */
        sc ss[2]; ge c;
        ss[0] = t1; ss[1] = t1b[j]; ge_straus(&c, ss, (const ge(*)[8])TAB_B, 2); ge_add(&T1, &T1, &c);
        ss[0] = t2; ss[1] = t2b[j]; ge_straus(&c, ss, (const ge(*)[8])TAB_B, 2); ge_add(&T2, &T2, &c);
        sc_add(&t0s, &t0s, &t0); sc_add(&t1s, &t1s, &t1); sc_add(&t2s, &t2s, &t2);
    }
    ge_compress(o, &T1); tr_append(&tr, "T_1", o, 32); o += 32;
    ge_compress(o, &T2); tr_append(&tr, "T_2", o, 32); o += 32;
    tr_challenge_sc(&tr, "x", &x);
    if (sc_iszero(&x)) return ERR_BAD_ARG;
    sc t_x, t_x_bl, e_bl, tmp, tmp2, xx;
    sc_mul(&xx, &x, &x);
    sc_mul(&tmp, &t1s, &x); sc_add(&t_x, &t0s, &tmp); sc_mul(&tmp, &t2s, &xx); sc_add(&t_x, &t_x, &tmp);
    sc_from_u64(&t_x_bl, 0); e_bl = t_x_bl;
    sc *lv = l0, *rv = r0; /* l(x) = l0 + l1 x ; r(x) = r0 + r1 x, in place */
    for (int j = 0; j < m; j++) {
        sc offset_zz, zj;
        sc_pow(&zj, &z, (uint64_t)j); sc_mul(&offset_zz, &zz, &zj);
        sc_mul(&tmp, &offset_zz, &bl[j]); sc_add(&t_x_bl, &t_x_bl, &tmp);
        sc_mul(&tmp, &t2b[j], &x); sc_add(&tmp, &tmp, &t1b[j]); sc_mul(&tmp, &tmp, &x); sc_add(&t_x_bl, &t_x_bl, &tmp);
        sc_mul(&tmp, &s_bl[j], &x); sc_add(&tmp, &tmp, &a_bl[j]); sc_add(&e_bl, &e_bl, &tmp);
        for (size_t i = 0; i < n; i++) {
            size_t k = j * n + i;
            sc_mul(&tmp, &sL[k], &x); sc_add(&lv[k], &l0[k], &tmp);
            sc_mul(&tmp, &r1[k], &x); sc_add(&rv[k], &r0[k], &tmp);
        }
    }
    sc_tobytes(o, &t_x); tr_append(&tr, "t_x", o, 32); o += 32;
    sc_tobytes(o, &t_x_bl); tr_append(&tr, "t_x_blinding", o, 32); o += 32;
    sc_tobytes(o, &e_bl); tr_append(&tr, "e_blinding", o, 32); o += 32;
    tr_challenge_sc(&tr, "w", &w);
    ge Q; ge_scalarmult(&Q, &w, &GE_B);
    /* InnerProductProof::create, G_factors = 1, H_factors = y^-i applied in the first round */
    tr_append(&tr, "dom-sep", "ipp v1", 6); tr_append_u64(&tr, "n", (uint64_t)N);
    sc y_inv; sc_invert(&y_inv, &y);
    sc *hf = malloc(sizeof(sc) * N), *sa = malloc(sizeof(sc) * N), *sb = malloc(sizeof(sc) * N);
    hf[0] = one; for (size_t i = 1; i < N; i++) sc_mul(&hf[i], &hf[i - 1], &y_inv);
    int first = 1;
    size_t len = N;
    while (len > 1) {
        size_t h = len / 2;
        sc cL, cR, u, ui;
        sc_inner(&cL, lv, rv + h, h); sc_inner(&cR, lv + h, rv, h);
        ge Lp, Rp, t;
        /* L = <a_L, G_R> + <b_R, H_L'> + c_L Q ; R = <a_R, G_L> + <b_L, H_R'> + c_R Q */
        for (size_t i = 0; i < h; i++) { if (first) { sc_mul(&sa[i], &rv[h + i], &hf[i]); sc_mul(&sb[i], &rv[i], &hf[h + i]); } else { sa[i] = rv[h + i]; sb[i] = rv[i]; } }
        ge_msm(&Lp, lv, Gv + h, h); ge_msm(&t, sa, Hv, h); ge_add(&Lp, &Lp, &t); ge_scalarmult(&t, &cL, &Q); ge_add(&Lp, &Lp, &t);
        ge_msm(&Rp, lv + h, Gv, h); ge_msm(&t, sb, Hv + h, h); ge_add(&Rp, &Rp, &t); ge_scalarmult(&t, &cR, &Q); ge_add(&Rp, &Rp, &t);
        ge_compress(o, &Lp); tr_append(&tr, "L", o, 32); o += 32;
        ge_compress(o, &Rp); tr_append(&tr, "R", o, 32); o += 32;
        tr_challenge_sc(&tr, "u", &u); sc_invert(&ui, &u);
        for (size_t i = 0; i < h; i++) {
            sc_mul(&tmp, &lv[i], &u); sc_mul(&tmp2, &lv[h + i], &ui); sc_add(&lv[i], &tmp, &tmp2);
            sc_mul(&tmp, &rv[i], &ui); sc_mul(&tmp2, &rv[h + i], &u); sc_add(&rv[i], &tmp, &tmp2);
            sc s2[2]; ge tb[2][8], r;
            s2[0] = ui; s2[1] = u;
            ge_table8(tb[0], &Gv[i]); ge_table8(tb[1], &Gv[h + i]); ge_straus(&r, s2, (const ge(*)[8])tb, 2); Gv[i] = r;
            if (first) { sc_mul(&s2[0], &u, &hf[i]); sc_mul(&s2[1], &ui, &hf[h + i]); } else { s2[0] = u; s2[1] = ui; }
            ge_table8(tb[0], &Hv[i]); ge_table8(tb[1], &Hv[h + i]); ge_straus(&r, s2, (const ge(*)[8])tb, 2); Hv[i] = r;
        }
        first = 0; len = h;
    }
    sc_tobytes(o, &lv[0]); o += 32;
    sc_tobytes(o, &rv[0]); o += 32;
    *out_len = (uint64_t)(o - out);
    free(bl); free(a_bl); free(s_bl); free(sL); free(sR); free(l0); free(r0); free(r1); free(t1b); free(t2b);
    free(Gv); free(Hv); free(hf); free(sa); free(sb);
    return ERR_OK;
}

/* RangeProof::verify_multiple over Transcript::new(&[]) (src/range/mod.rs:83-119); 1 = accept */
EXPORT int dor_rp_verify(int nbits, int m, const uint8_t *proof, uint64_t len, const uint8_t *commitments) {
    if (!(nbits == 8 || nbits == 16 || nbits == 32 || nbits == 64) || m <= 0 || (m & (m - 1))) return 0;
    /* RangeProof::from_bytes / InnerProductProof::from_bytes */
    if (len % 32 || len < 7 * 32) return 0;
    uint64_t ne = len / 32 - 7;
    if (ne < 2 || (ne - 2) % 2) return 0;
    uint64_t lg = (ne - 2) / 2;
    if (lg >= 32) return 0;
    sc t_x, t_x_bl, e_bl, a, b;
    if (!sc_frombytes_canonical(&t_x, proof + 128) || !sc_frombytes_canonical(&t_x_bl, proof + 160) ||
        !sc_frombytes_canonical(&e_bl, proof + 192) || !sc_frombytes_canonical(&a, proof + len - 64) ||
        !sc_frombytes_canonical(&b, proof + len - 32)) return 0;
    size_t n = (size_t)nbits, N = n * (size_t)m;
    ensure_gens(m);
    transcript tr; tr_init(&tr, "", 0);
    tr_append(&tr, "dom-sep", "rangeproof v1", 13);
    tr_append_u64(&tr, "n", (uint64_t)n); tr_append_u64(&tr, "m", (uint64_t)m);
    for (int j = 0; j < m; j++) tr_append(&tr, "V", commitments + 32 * j, 32);
    if (is_zero32(proof) || is_zero32(proof + 32)) return 0;
    tr_append(&tr, "A", proof, 32); tr_append(&tr, "S", proof + 32, 32);
    sc y, z, zz, x, w, one;
    sc_from_u64(&one, 1);
    tr_challenge_sc(&tr, "y", &y); tr_challenge_sc(&tr, "z", &z); sc_mul(&zz, &z, &z);
    if (is_zero32(proof + 64) || is_zero32(proof + 96)) return 0;
    tr_append(&tr, "T_1", proof + 64, 32); tr_append(&tr, "T_2", proof + 96, 32);
    tr_challenge_sc(&tr, "x", &x);
    tr_append(&tr, "t_x", proof + 128, 32); tr_append(&tr, "t_x_blinding", proof + 160, 32); tr_append(&tr, "e_blinding", proof + 192, 32);
    tr_challenge_sc(&tr, "w", &w);
    if (((uint64_t)1 << lg) != N) return 0;
    tr_append(&tr, "dom-sep", "ipp v1", 6); tr_append_u64(&tr, "n", (uint64_t)N);
    sc us[32], u_sq[32], u_inv[32], u_inv_sq[32];
    const uint8_t *lr = proof + 224;
    for (uint64_t k = 0; k < lg; k++) {
        if (is_zero32(lr + 64 * k) || is_zero32(lr + 64 * k + 32)) return 0;
        tr_append(&tr, "L", lr + 64 * k, 32); tr_append(&tr, "R", lr + 64 * k + 32, 32);
        tr_challenge_sc(&tr, "u", &us[k]);
    }
    /* batching weight c: any value (thread_rng in the reference); derived from the transcript here */
    sc c; tr_challenge_sc(&tr, "dapol-b200 batching weight", &c);
    size_t npts = 2 * N + 2 * lg + (size_t)m + 6;
    sc *ss = malloc(sizeof(sc) * npts); ge *pp = malloc(sizeof(ge) * npts);
    size_t q = 0;
    int ok = 1;
    sc tmp, tmp2;
    ok &= ge_decompress(&pp[q], proof); ss[q++] = one;                                   /* A */
    ok &= ge_decompress(&pp[q], proof + 32); ss[q++] = x;                                /* x S */
    ok &= ge_decompress(&pp[q], proof + 64); sc_mul(&ss[q], &c, &x); q++;                /* c x T1 */
    ok &= ge_decompress(&pp[q], proof + 96); sc_mul(&tmp, &x, &x); sc_mul(&ss[q], &c, &tmp); q++; /* c x^2 T2 */
    sc allinv; sc_from_u64(&allinv, 1);
    for (uint64_t k = 0; k < lg; k++) {
        sc_mul(&u_sq[k], &us[k], &us[k]); sc_invert(&u_inv[k], &us[k]); sc_mul(&u_inv_sq[k], &u_inv[k], &u_inv[k]);
        sc_mul(&allinv, &allinv, &u_inv[k]);
        ok &= ge_decompress(&pp[q], lr + 64 * k); ss[q++] = u_sq[k];
        ok &= ge_decompress(&pp[q], lr + 64 * k + 32); ss[q++] = u_inv_sq[k];
    }
    for (int j = 0; j < m; j++) {                                                         /* c z^(2+j) V_j */
        sc zj; sc_pow(&zj, &z, (uint64_t)j); sc_mul(&tmp, &zz, &zj); sc_mul(&ss[q], &c, &tmp);
        ok &= ge_decompress(&pp[q], commitments + 32 * j); q++;
    }
    if (!ok) { free(ss); free(pp); return 0; }
    sc *s = malloc(sizeof(sc) * N);
    s[0] = allinv;
    for (size_t i = 1; i < N; i++) { int lgi = 63 - __builtin_clzll((unsigned long long)i); sc_mul(&s[i], &s[i - ((size_t)1 << lgi)], &u_sq[lg - 1 - lgi]); }
    sc y_inv, sum_y, sum_z, delta, yp, zp;
    sc_invert(&y_inv, &y);
    sc_from_u64(&sum_y, 0); yp = one; for (size_t i = 0; i < N; i++) { sc_add(&sum_y, &sum_y, &yp); sc_mul(&yp, &yp, &y); }
    sc_from_u64(&sum_z, 0); zp = one; for (int j = 0; j < m; j++) { sc_add(&sum_z, &sum_z, &zp); sc_mul(&zp, &zp, &z); }
    /* delta = (z - z^2) sum_y - z^3 (2^n - 1) sum_z */
    sc two_n; if (n == 64) { two_n.v[0] = UINT64_MAX; two_n.v[1] = two_n.v[2] = two_n.v[3] = 0; } else sc_from_u64(&two_n, ((uint64_t)1 << n) - 1);
    sc_sub(&tmp, &z, &zz); sc_mul(&delta, &tmp, &sum_y);
    sc_mul(&tmp, &zz, &z); sc_mul(&tmp, &tmp, &two_n); sc_mul(&tmp, &tmp, &sum_z); sc_sub(&delta, &delta, &tmp);
    /* B_blinding: -e_bl - c t_x_bl ; B: w (t_x - a b) + c (delta - t_x) */
    pp[q] = GE_BBL; sc_mul(&tmp, &c, &t_x_bl); sc_add(&tmp, &tmp, &e_bl); sc_neg(&ss[q], &tmp); q++;
    pp[q] = GE_B; sc_mul(&tmp, &a, &b); sc_sub(&tmp, &t_x, &tmp); sc_mul(&tmp, &w, &tmp);
    sc_sub(&tmp2, &delta, &t_x); sc_mul(&tmp2, &c, &tmp2); sc_add(&ss[q], &tmp, &tmp2); q++;
    sc yip = one, minus_z; sc_neg(&minus_z, &z);
    for (size_t i = 0; i < N; i++) {
        size_t j = i / n, ii = i % n;
        sc zj, two_i;
        /* G_i: -z - a s_i */
        sc_mul(&tmp, &a, &s[i]); sc_sub(&ss[q], &minus_z, &tmp); pp[q] = g_G[j * 64 + ii]; q++;
        /* H_i: z + y^-i (z^2 z^j 2^ii - b s_{N-1-i}) */
        sc_pow(&zj, &z, (uint64_t)j); sc_mul(&zj, &zj, &zz);
        sc_from_u64(&two_i, (uint64_t)1 << ii); sc_mul(&zj, &zj, &two_i);
        sc_mul(&tmp, &b, &s[N - 1 - i]); sc_sub(&tmp, &zj, &tmp); sc_mul(&tmp, &tmp, &yip); sc_add(&ss[q], &z, &tmp);
        pp[q] = g_H[j * 64 + ii]; q++;
        sc_mul(&yip, &yip, &y_inv);
    }
    ge mega; ge_msm(&mega, ss, pp, q);
    int res = ge_is_identity(&mega);
    free(ss); free(pp); free(s);
    return res;
}

/* ------------------------------------------------------------------ aggregation policies + wire formats */
enum { POLICY_PADDING = 0, POLICY_SPLITTING = 1 };
static uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static void put_be(uint8_t *o, uint64_t x, int k) { for (int i = 0; i < k; i++) o[i] = (uint8_t)(x >> (8 * (k - 1 - i))); }
static uint64_t get_be(const uint8_t *o, int k) { uint64_t x = 0; for (int i = 0; i < k; i++) x = (x << 8) | o[i]; return x; }
static uint64_t rp_size(uint64_t m) { uint64_t lg = 0; while (((uint64_t)1 << lg) < 64 * m) lg++; return 32 * (9 + 2 * lg); }

/* plan: groups[(start,count,m)] then singles from `pos` (padding.rs:88-118, splitting.rs:100-129) */
typedef struct { uint64_t start, count, m; } agg_group;
static int policy_plan(uint64_t nsib, uint64_t agg, int policy, agg_group *g, int *ng, uint64_t *single_from) {
    if (agg > nsib) return ERR_BAD_ARG; /* reference: slice out-of-bounds panic */
    *ng = 0;
    if (policy == POLICY_PADDING) { g[0].start = 0; g[0].count = agg; g[0].m = next_pow2(agg); *ng = 1; *single_from = agg; }
    else {
        uint64_t base = next_pow2(agg), pos = 0;
        while (pos < agg) { if (agg & base) { g[*ng].start = pos; g[*ng].count = base; g[*ng].m = base; (*ng)++; pos += base; } base >>= 1; }
        *single_from = pos;
    }
    return ERR_OK;
}
EXPORT uint64_t dor_inclusion_proof_size(int height, uint64_t agg, int policy) {
    agg_group g[64]; int ng; uint64_t sf;
    if (policy_plan((uint64_t)height, agg, policy, g, &ng, &sf)) return 0;
    uint64_t sz = policy == POLICY_SPLITTING ? 2 : 0;
    for (int i = 0; i < ng; i++) sz += 8 + rp_size(g[i].m);
    sz += 8 + 672 * ((uint64_t)height - sf);
    sz += 2 + 8 + ((uint64_t)height + 7) / 8 + 8 + 64 * (uint64_t)height;
    return sz;
}
/* the same for a digest of dlen bytes: a sibling is com (32) || hash (dlen), src/proof/node.rs:74-79 */
EXPORT uint64_t dor_inclusion_proof_size_d(int height, uint64_t agg, int policy, int dlen) {
    uint64_t sz = dor_inclusion_proof_size(height, agg, policy);
    return sz ? sz + (uint64_t)(dlen - 32) * (uint64_t)height : 0;
}
/* ChaCha20 key of the prover's nonce streams of one tree (RNG contract, include/dapol_b200.h): BLAKE3(label || seed || root
 * commitment || root hash || le64 policy || le64 aggregation factor || le64 height) -- the seed bound to the tree (its root
 * commits to every witness), the policy and the factor, so a re-used seed never re-uses a nonce with another witness. */
static void prover_nonce_key_d(const uint8_t seed[32], const uint8_t root_com[32], const uint8_t *root_hash, int dl, int policy, uint64_t agg,
                               int height, uint8_t key[32]) {
    static const char label[] = "dapol-b200 prover nonce key v1";
    uint8_t buf[30 + 64 + 64 + 24];
    uint64_t w[3] = {(uint64_t)policy, agg, (uint64_t)height};
    memcpy(buf, label, 30); memcpy(buf + 30, seed, 32); memcpy(buf + 62, root_com, 32); memcpy(buf + 94, root_hash, dl);
    for (int i = 0; i < 3; i++) for (int k = 0; k < 8; k++) buf[94 + dl + 8 * i + k] = (uint8_t)(w[i] >> (8 * k));
    hash_buf(0, buf, 94 + dl + 24, key);
}
EXPORT void dor_prover_nonce_key(const uint8_t seed[32], const uint8_t root_com[32], const uint8_t root_hash[32], int policy, uint64_t agg,
                                 int height, uint8_t key[32]) {
    prover_nonce_key_d(seed, root_com, root_hash, 32, policy, agg, height, key);
}
/* Dapol::generate_proof (mod.rs:167-190) + DapolProof::serialize (proof/mod.rs:68-73).
 * RNG contract: range proof #q of this DapolProof (aggregated first, then singles) draws from
 * ChaCha20(dor_prover_nonce_key(seed, root, policy, agg, height), stream = leaf_idx) starting at block q << 32. */
EXPORT int dor_prove_inclusion(const dor_tree *t, uint64_t leaf_idx, uint64_t agg, int policy, const uint8_t seed[32],
                               uint8_t *out, uint64_t cap, uint64_t *out_len) {
    const node *sib[64];
    int rc = tree_path(t, leaf_idx, sib, NULL);
    if (rc) return rc;
    const int dl = DLEN(t->hash_id);
    uint64_t H = (uint64_t)t->height, need = dor_inclusion_proof_size_d(t->height, agg, policy, dl);
    if (!need) return ERR_BAD_ARG;
    if (cap < need) return ERR_BUFFER;
    agg_group g[64]; int ng; uint64_t sf;
    policy_plan(H, agg, policy, g, &ng, &sf);
    uint8_t key[32];
    prover_nonce_key_d(seed, t->lv[0].nodes[0].comc, t->lv[0].nodes[0].hash, DLEN(t->hash_id), policy, agg, t->height, key);
    seed = key;
    uint8_t *o = out; uint64_t q = 0, plen;
    if (policy == POLICY_SPLITTING) { put_be(o, (uint64_t)ng, 2); o += 2; }
    for (int i = 0; i < ng; i++) {
        uint64_t vals[64]; uint8_t bls[64 * 32];
        memset(vals, 0, sizeof vals); memset(bls, 0, sizeof bls);
        for (uint64_t k = 0; k < g[i].m; k++) {
            if (k < g[i].count) { vals[k] = sib[g[i].start + k]->v; memcpy(bls + 32 * k, sib[g[i].start + k]->r, 32); }
            else bls[32 * k] = 1; /* (0, Scalar::one()) padding party (padding.rs:98-101) */
        }
        rc = dor_rp_prove(64, (int)g[i].m, vals, bls, seed, leaf_idx, q << 32, o + 8, &plen);
        if (rc) return rc;
        put_be(o, plen, 8); o += 8 + plen; q++;
    }
    put_be(o, H - sf, 8); o += 8;
    for (uint64_t k = sf; k < H; k++) {
        rc = dor_rp_prove(64, 1, &sib[k]->v, sib[k]->r, seed, leaf_idx, q << 32, o, &plen);
        if (rc) return rc;
        o += plen; q++;
    }
    /* MerkleProof::serialize (smtree, UPSTREAM-RECALL: widths unverified) */
    put_be(o, H, 2); o += 2; put_be(o, 1, 8); o += 8;
    uint64_t nb = (H + 7) / 8;
    if (nb) { put_be(o, H == 64 ? leaf_idx : leaf_idx << (8 * nb - H), (int)nb); o += nb; }
    put_be(o, H, 8); o += 8;
    for (uint64_t k = 0; k < H; k++) { memcpy(o, sib[k]->comc, 32); memcpy(o + 32, sib[k]->hash, dl); o += 32 + dl; }
    *out_len = (uint64_t)(o - out);
    return ERR_OK;
}
/* DapolProof::deserialize + verify (proof/mod.rs:41-47,76-95); 1 = accept */
EXPORT int dor_verify_inclusion(int hash_id, int policy, const uint8_t *p, uint64_t len, const uint8_t root_com[32],
                                const uint8_t *root_hash, const uint8_t leaf_com[32], const uint8_t *leaf_hash) {
    const int dl = DLEN(hash_id);
    const uint64_t ss = 32 + (uint64_t)dl; /* bytes of one sibling */
    uint64_t pos = 0, nagg = 1;
    const uint8_t *aggp[64]; uint64_t aggl[64];
#define NEED(k) do { if (len - pos < (uint64_t)(k)) return 0; } while (0)
    if (policy == POLICY_SPLITTING) { NEED(2); nagg = get_be(p, 2); pos += 2; if (nagg > 64) return 0; }
    for (uint64_t i = 0; i < nagg; i++) {
        NEED(8); uint64_t sz = get_be(p + pos, 8); pos += 8;
        if (sz > len - pos) return 0;
        aggp[i] = p + pos; aggl[i] = sz; pos += sz;
    }
    NEED(8); uint64_t nind = get_be(p + pos, 8); pos += 8;
    if (nind > 64 || (len - pos) / 672 < nind) return 0;
    const uint8_t *ind = p + pos; pos += 672 * nind;
    NEED(10); uint64_t H = get_be(p + pos, 2); pos += 2;
    if (get_be(p + pos, 8) != 1 || H > 64) return 0;
    pos += 8;
    uint64_t nb = (H + 7) / 8; NEED(nb + 8);
    uint64_t idx = nb ? get_be(p + pos, (int)nb) : 0;
    if (nb && H != 64) idx >>= (8 * nb - H);
    pos += nb;
    uint64_t nsib = get_be(p + pos, 8); pos += 8;
    if (nsib != H || (len - pos) / ss < nsib) return 0;
    const uint8_t *sib = p + pos;
    /* MerkleProof::verify: fold upward with DapolProofNode::merge (proof/node.rs:56-69) */
    ge cur, s; uint8_t curc[32], curh[64], buf[192];
    if (!ge_decompress(&cur, leaf_com)) return 0;
    memcpy(curc, leaf_com, 32); memcpy(curh, leaf_hash, dl);
    for (uint64_t k = 0; k < H; k++) {
        const uint8_t *sc_ = sib + ss * k, *sh = sc_ + 32;
        if (!ge_decompress(&s, sc_)) return 0;
        int cur_is_right = (int)((idx >> k) & 1);
        memcpy(buf, cur_is_right ? sc_ : curc, 32); memcpy(buf + 32, cur_is_right ? curc : sc_, 32);
        memcpy(buf + 64, cur_is_right ? sh : curh, dl); memcpy(buf + 64 + dl, cur_is_right ? curh : sh, dl);
        hash_buf(hash_id, buf, 64 + 2 * dl, curh);
        ge_add(&cur, &cur, &s); ge_compress(curc, &cur);
    }
    if (memcmp(curc, root_com, 32) || memcmp(curh, root_hash, dl)) return 0;
    /* R::verify on the siblings' commitments (padding.rs:168-197 / splitting.rs:180-211) */
    if (nind > nsib) return 0;
    uint64_t n_agg_coms = nsib - nind;
    uint8_t coms[64 * 32];
    if (policy == POLICY_PADDING) {
        uint64_t m = next_pow2(n_agg_coms);
        uint8_t bbl[32]; ge_compress(bbl, &GE_BBL);
        for (uint64_t k = 0; k < m; k++) memcpy(coms + 32 * k, k < n_agg_coms ? sib + ss * k : bbl, 32);
        if (!dor_rp_verify(64, (int)m, aggp[0], aggl[0], coms)) return 0;
    } else {
        uint64_t base = next_pow2(n_agg_coms), at = 0, i = 0;
        while (at < n_agg_coms) {
            if (n_agg_coms & base) {
                if (i >= nagg) return 0;
                for (uint64_t k = 0; k < base; k++) memcpy(coms + 32 * k, sib + ss * (at + k), 32);
                if (!dor_rp_verify(64, (int)base, aggp[i], aggl[i], coms)) return 0;
                i++; at += base;
            }
            base >>= 1;
        }
    }
    for (uint64_t k = 0; k < nind; k++)
        if (!dor_rp_verify(64, 1, ind + 672 * k, 672, sib + ss * (n_agg_coms + k))) return 0;
    return 1;
}

/* ------------------------------------------------------------------ batch proofs: ONE DapolProof for several leaves
 * Dapol::generate_proof_batch (src/dapol/mod.rs:172-190), DapolProof::verify_batch (src/proof/mod.rs:49-54), test shape
 * src/proof/tests.rs:6-35.  smtree get_merkle_path_ref_batch / MerkleProof::verify_batch (UPSTREAM-RECALL):
 * level by level from the leaves up, left to right; a node's sibling is listed only if it is not itself on the way up. */
typedef struct { int h; uint64_t idx; } sib_ref;
/* siblings a batch proof carries, in proof order; idx strictly increasing.  Returns the count (plan may be NULL). */
static uint64_t batch_sibling_plan(int height, uint64_t k, const uint64_t *idx, sib_ref *plan) {
    uint64_t *cur = malloc(8 * k), n = k, out = 0;
    memcpy(cur, idx, 8 * k);
    for (int h = height; h >= 1; h--) {
        for (uint64_t i = 0; i < n; i++) {
            uint64_t s = cur[i] ^ 1;
            int have = (i > 0 && cur[i - 1] == s) || (i + 1 < n && cur[i + 1] == s);
            if (!have) { if (plan) { plan[out].h = h; plan[out].idx = s; } out++; }
        }
        uint64_t m = 0;
        for (uint64_t i = 0; i < n; i++) if (!m || cur[m - 1] != cur[i] >> 1) cur[m++] = cur[i] >> 1;
        n = m;
    }
    free(cur);
    return out;
}
/* nonce key of a batch of more than one leaf: the tree's prover key chained over the leaf indexes, 64 per link (stream 0) */
static void batch_nonce_key(uint8_t key[32], uint64_t k, const uint64_t *idx) {
    static const char label[] = "dapol-b200 batch proof nonce key v1";
    uint8_t buf[35 + 32 + 8 + 512];
    memcpy(buf, label, 35); memcpy(buf + 35, key, 32);
    for (int b = 0; b < 8; b++) buf[67 + b] = (uint8_t)(k >> (8 * b));
    hash_buf(0, buf, 75, key);
    for (uint64_t i = 0; i < k; i += 64) {
        uint64_t c = k - i < 64 ? k - i : 64;
        memcpy(buf, key, 32);
        for (uint64_t j = 0; j < c; j++) for (int b = 0; b < 8; b++) buf[32 + 8 * j + b] = (uint8_t)(idx[i + j] >> (8 * b));
        hash_buf(0, buf, 32 + 8 * c, key);
    }
}
static uint64_t range_part_size(uint64_t nsib, uint64_t agg, int policy) {
    agg_group g[64]; int ng; uint64_t sf;
    if (policy_plan(nsib, agg, policy, g, &ng, &sf)) return 0;
    uint64_t sz = policy == POLICY_SPLITTING ? 2 : 0;
    for (int i = 0; i < ng; i++) { if (g[i].m > 64) return 0; sz += 8 + rp_size(g[i].m); }
    return sz + 8 + 672 * (nsib - sf);
}
EXPORT uint64_t dor_batch_proof_size(int height, uint64_t k, const uint64_t *idx, uint64_t agg, int policy) {
    if (k == 0 || height < 0 || height > 64) return 0;
    for (uint64_t i = 1; i < k; i++) if (idx[i] <= idx[i - 1]) return 0;
    uint64_t nsib = batch_sibling_plan(height, k, idx, NULL), rs = range_part_size(nsib, agg, policy);
    if (!rs) return 0;
    return rs + 2 + 8 + k * (((uint64_t)height + 7) / 8) + 8 + 64 * nsib;
}
EXPORT uint64_t dor_batch_proof_size_d(int height, uint64_t k, const uint64_t *idx, uint64_t agg, int policy, int dlen) {
    uint64_t sz = dor_batch_proof_size(height, k, idx, agg, policy);
    return sz ? sz + (uint64_t)(dlen - 32) * batch_sibling_plan(height, k, idx, NULL) : 0;
}
EXPORT int dor_prove_inclusion_batch(const dor_tree *t, uint64_t k, const uint64_t *idx, uint64_t agg, int policy, const uint8_t seed[32],
                                     uint8_t *out, uint64_t cap, uint64_t *out_len) {
    if (k == 1) return dor_prove_inclusion(t, idx[0], agg, policy, seed, out, cap, out_len);
    const int dl = DLEN(t->hash_id);
    uint64_t need = dor_batch_proof_size_d(t->height, k, idx, agg, policy, dl);
    if (!need) return ERR_BAD_ARG;
    if (cap < need) return ERR_BUFFER;
    for (uint64_t i = 0; i < k; i++) { const node *lf = tree_find(t, t->height, idx[i]); if (!lf || lf->is_pad) return ERR_NOT_FOUND; }
    uint64_t nsib = batch_sibling_plan(t->height, k, idx, NULL);
    sib_ref *plan = malloc(sizeof(sib_ref) * (nsib + 1));
    const node **sib = malloc(sizeof(node *) * (nsib + 1));
    batch_sibling_plan(t->height, k, idx, plan);
    for (uint64_t i = 0; i < nsib; i++) sib[i] = tree_find(t, plan[i].h, plan[i].idx);
    uint8_t key[32];
    prover_nonce_key_d(seed, t->lv[0].nodes[0].comc, t->lv[0].nodes[0].hash, DLEN(t->hash_id), policy, agg, t->height, key);
    batch_nonce_key(key, k, idx);
    agg_group g[64]; int ng; uint64_t sf, q = 0, plen; int rc = 0;
    policy_plan(nsib, agg, policy, g, &ng, &sf);
    uint8_t *o = out;
    if (policy == POLICY_SPLITTING) { put_be(o, (uint64_t)ng, 2); o += 2; }
    for (int i = 0; i < ng && !rc; i++) {
        uint64_t vals[64]; uint8_t bls[64 * 32];
        memset(vals, 0, sizeof vals); memset(bls, 0, sizeof bls);
        for (uint64_t j = 0; j < g[i].m; j++) {
            if (j < g[i].count) { vals[j] = sib[g[i].start + j]->v; memcpy(bls + 32 * j, sib[g[i].start + j]->r, 32); }
            else bls[32 * j] = 1;
        }
        rc = dor_rp_prove(64, (int)g[i].m, vals, bls, key, 0, q << 32, o + 8, &plen);
        put_be(o, plen, 8); o += 8 + plen; q++;
    }
    put_be(o, nsib - sf, 8); o += 8;
    for (uint64_t j = sf; j < nsib && !rc; j++) { rc = dor_rp_prove(64, 1, &sib[j]->v, sib[j]->r, key, 0, q << 32, o, &plen); o += plen; q++; }
    uint64_t H = (uint64_t)t->height, nb = (H + 7) / 8;
    put_be(o, H, 2); o += 2; put_be(o, k, 8); o += 8;
    for (uint64_t i = 0; i < k; i++) if (nb) { put_be(o, H == 64 ? idx[i] : idx[i] << (8 * nb - H), (int)nb); o += nb; }
    put_be(o, nsib, 8); o += 8;
    for (uint64_t j = 0; j < nsib; j++) { memcpy(o, sib[j]->comc, 32); memcpy(o + 32, sib[j]->hash, dl); o += 32 + dl; }
    *out_len = (uint64_t)(o - out);
    free(plan); free(sib);
    return rc;
}
typedef struct { uint64_t idx; ge p; uint8_t c[32], h[64]; } pnode;
/* DapolProof::deserialize + verify_batch(root, leaves) (proof/mod.rs:49-54,76-95); leaves in index order; 1 = accept */
EXPORT int dor_verify_inclusion_batch(int hash_id, int policy, const uint8_t *p, uint64_t len, const uint8_t root_com[32], const uint8_t *root_hash,
                                      uint64_t k, const uint8_t *leaf_coms, const uint8_t *leaf_hashes) {
    const int dl = DLEN(hash_id);
    const uint64_t ss = 32 + (uint64_t)dl;
    uint64_t pos = 0, nagg = 1;
    const uint8_t *aggp[64]; uint64_t aggl[64];
#define NEEDB(n) do { if (len - pos < (uint64_t)(n)) return 0; } while (0)
    if (policy == POLICY_SPLITTING) { NEEDB(2); nagg = get_be(p, 2); pos += 2; if (nagg > 64) return 0; }
    for (uint64_t i = 0; i < nagg; i++) {
        NEEDB(8); uint64_t sz = get_be(p + pos, 8); pos += 8;
        if (sz > len - pos) return 0;
        aggp[i] = p + pos; aggl[i] = sz; pos += sz;
    }
    NEEDB(8); uint64_t nind = get_be(p + pos, 8); pos += 8;
    if ((len - pos) / 672 < nind) return 0;
    const uint8_t *ind = p + pos; pos += 672 * nind;
    NEEDB(10); uint64_t H = get_be(p + pos, 2); pos += 2;
    uint64_t nl = get_be(p + pos, 8); pos += 8;
    if (H > 64 || nl != k || k == 0) return 0;
    uint64_t nb = (H + 7) / 8;
    if ((len - pos) / (nb ? nb : 1) < k && nb) return 0;
    uint64_t *idx = malloc(8 * k);
    for (uint64_t i = 0; i < k; i++) { uint64_t x = nb ? get_be(p + pos, (int)nb) : 0; if (nb && H != 64) x >>= (8 * nb - H); idx[i] = x; pos += nb; }
    int ok = 1;
    for (uint64_t i = 1; i < k; i++) if (idx[i] <= idx[i - 1]) ok = 0;
    uint64_t nsib = 0;
    if (ok && len - pos >= 8) { nsib = get_be(p + pos, 8); pos += 8; } else ok = 0;
    if (ok && ((len - pos) / ss < nsib || nsib != batch_sibling_plan((int)H, k, idx, NULL) || nind > nsib)) ok = 0;
    if (!ok) { free(idx); return 0; }
    const uint8_t *sib = p + pos;
    sib_ref *plan = malloc(sizeof(sib_ref) * (nsib + 1));
    batch_sibling_plan((int)H, k, idx, plan);
    pnode *cur = malloc(sizeof(pnode) * k), *nxt = malloc(sizeof(pnode) * k);
    uint64_t n = k, used = 0;
    for (uint64_t i = 0; i < k && ok; i++) {
        cur[i].idx = idx[i]; memcpy(cur[i].c, leaf_coms + 32 * i, 32); memcpy(cur[i].h, leaf_hashes + (uint64_t)dl * i, dl);
        ok = ge_decompress(&cur[i].p, cur[i].c);
    }
    /* siblings are consumed in plan order: all of a level (left to right) before the next level up */
    for (int h = (int)H; h >= 1 && ok; h--) {
        uint64_t m = 0, at = used;   /* `at` walks this level's siblings in the order the plan lists them */
        /* first pass: how many siblings does this level consume (so the next level starts after them) */
        uint64_t lvl_cnt = 0;
        for (uint64_t j = used; j < nsib && plan[j].h == h; j++) lvl_cnt++;
        for (uint64_t i = 0; i < n && ok; ) {
            pnode other; const pnode *l, *r; uint64_t step = 1;
            if (i + 1 < n && cur[i + 1].idx == (cur[i].idx ^ 1)) { other = cur[i + 1]; step = 2; }
            else {
                if (at >= used + lvl_cnt || plan[at].idx != (cur[i].idx ^ 1)) { ok = 0; break; }
                other.idx = plan[at].idx; memcpy(other.c, sib + ss * at, 32); memcpy(other.h, sib + ss * at + 32, dl);
                ok = ge_decompress(&other.p, other.c); at++;
                if (!ok) break;
            }
            if (cur[i].idx & 1) { l = &other; r = &cur[i]; } else { l = &cur[i]; r = &other; }
            uint8_t buf[192];
            memcpy(buf, l->c, 32); memcpy(buf + 32, r->c, 32); memcpy(buf + 64, l->h, dl); memcpy(buf + 64 + dl, r->h, dl);
            nxt[m].idx = cur[i].idx >> 1;
            hash_buf(hash_id, buf, 64 + 2 * dl, nxt[m].h);
            ge_add(&nxt[m].p, &l->p, &r->p); ge_compress(nxt[m].c, &nxt[m].p);
            m++; i += step;
        }
        if (ok && at != used + lvl_cnt) ok = 0;
        used += lvl_cnt;
        pnode *t_ = cur; cur = nxt; nxt = t_; n = m;
    }
    if (ok && (n != 1 || memcmp(cur[0].c, root_com, 32) || memcmp(cur[0].h, root_hash, dl))) ok = 0;
    free(cur); free(nxt); free(plan); free(idx);
    if (!ok) return 0;
    /* R::verify on the siblings' commitments, in proof order (padding.rs:168-197 / splitting.rs:180-211) */
    uint64_t n_agg_coms = nsib - nind;
    uint8_t coms[64 * 32];
    if (n_agg_coms > 64) return 0;
    if (policy == POLICY_PADDING) {
        uint64_t m = next_pow2(n_agg_coms);
        uint8_t bbl[32]; ge_compress(bbl, &GE_BBL);
        for (uint64_t j = 0; j < m; j++) memcpy(coms + 32 * j, j < n_agg_coms ? sib + ss * j : bbl, 32);
        if (nagg != 1 || !dor_rp_verify(64, (int)m, aggp[0], aggl[0], coms)) return 0;
    } else {
        uint64_t base = next_pow2(n_agg_coms), at = 0, i = 0;
        while (at < n_agg_coms) {
            if (n_agg_coms & base) {
                if (i >= nagg) return 0;
                for (uint64_t j = 0; j < base; j++) memcpy(coms + 32 * j, sib + ss * (at + j), 32);
                if (!dor_rp_verify(64, (int)base, aggp[i], aggl[i], coms)) return 0;
                i++; at += base;
            }
            base >>= 1;
        }
    }
    for (uint64_t j = 0; j < nind; j++)
        if (!dor_rp_verify(64, 1, ind + 672 * j, 672, sib + ss * (n_agg_coms + j))) return 0;
    return 1;
#undef NEEDB
}
