// Per-thread bodies of the level-synchronous padded sparse-Merkle-sum-tree build.
//
// What the reference does one node at a time on one CPU thread
//   (/root/reference/src/dapol/mod.rs:118-121 -> smtree SparseMerkleTree::build, calling back into
//    DapolNode::new / ::padding / ::merge, src/dapol/node.rs:29-45,64-80,86-88)
// is split here into data-parallel passes over flat per-level arrays in HBM:
//   structure pass  (integer/HBM-bound): pair detection, parent ranks, padding-node slots + RNG ordinals
//   leaf pass       (IMAD-bound): v*B + r*B_blinding by fixed-base comb, compress, D(com)
//   padding pass    (IMAD-bound): ChaCha20 draw -> r*B_blinding, compress, D(com)     [dominant, SURVEY F7]
//   merge sums/level(IMAD/HBM): v, r and point sums of the parents (no compressed point needed)
//   compress pass   (IMAD-bound): compress of the internal nodes of ALL levels, one balanced launch
//   merge hash/level(HBM-bound): D(C_L||C_R||H_L||H_R)
//
// Node store (struct of arrays, one global numbering): level h (0 = root .. H = leaves) occupies
// [level_off[h], level_off[h] + n_h); inside a level the two children of parent j sit at 2j, 2j+1,
// so a node's sibling is position ^ 1 and no child pointers are stored.
#pragma once
#include "ge25519.cuh"
#include "hash_dev.cuh"
#ifndef DAPOL_MAX_TREE_HEIGHT
#define DAPOL_MAX_TREE_HEIGHT 64  // include/dapol_b200.h
#endif

// ------------------------------------------------------------------------------------------------
// structure pass, level h: idx[0..c) = sorted tree indexes of the real (non-padding) nodes
// flags[k] = (starts_new_parent << 32) | is_lone
DAPOL_HD_INLINE uint64_t struct_flags_body(uint64_t k, const uint64_t *idx, uint64_t c) {
    uint64_t x = idx[k];
    int newp = (k == 0) || ((idx[k - 1] >> 1) != (x >> 1));
    int has_sib = (k > 0 && idx[k - 1] == (x ^ 1)) || (k + 1 < c && idx[k + 1] == (x ^ 1));
    return ((uint64_t)newp << 32) | (uint64_t)(!has_sib);
}
// s = exclusive prefix sum of flags at k.  For real node k of the level: its slot in the level's node array
// and its parent's tree index (compaction = next level's real nodes); for a lone node also its padding
// sibling's slot + idx, the pad's destination (global node number) at its build ordinal and its RNG block
// (pad_rng_base + rank inside the level: single tree = running creation ordinal; shard = the level's global base).
DAPOL_HD_INLINE void struct_apply_body(uint64_t k, const uint64_t *idx, uint64_t f, uint64_t s, uint32_t *pos, uint64_t *parent_idx,
                                       uint64_t level_off, const NodeStore &ns, uint64_t *pad_dest, uint64_t pad_ord_base,
                                       uint64_t *pad_rng, uint64_t pad_rng_base, int positional = 0) {
    uint64_t x = idx[k];
    uint32_t newp = (uint32_t)(f >> 32), lone = (uint32_t)f;
    uint64_t j = (s >> 32) - (newp ? 0 : 1);
    uint64_t q = s & 0xffffffffull;
    uint32_t slot = (uint32_t)(x & 1);
    uint64_t p = 2 * j + slot;
    pos[k] = (uint32_t)p;
    ns.idx[level_off + p] = x;
    ns.is_pad[level_off + p] = 0;
    if (newp) parent_idx[j] = x >> 1;
    if (lone) {
        uint64_t pp = 2 * j + (1 - slot);
        ns.idx[level_off + pp] = x ^ 1;
        ns.is_pad[level_off + pp] = 1;
        pad_dest[pad_ord_base + q] = level_off + pp;
        // block of the seeded stream this padding node draws: creation order (RNG contract) or, in the positional mode, the
        // node's own index (pad_rng_base = index of the subtree's first node of the level inside the whole tree)
        pad_rng[pad_ord_base + q] = positional ? pad_rng_base + (x ^ 1) : pad_rng_base + q;
    }
}
// Sizes of every level from one pass over adjacent leaves: msb of idx[k] ^ idx[k-1] (k >= 1).  The number
// of real nodes at level h is 1 + #{k : msb >= H - h}.  Returns -1 for k = 0; sets *bad on unsorted /
// out-of-tree input (smtree rejects those).
DAPOL_HD_INLINE int leaf_pair_msb(uint64_t k, const uint64_t *idx, int height, int *bad) {
    uint64_t x = idx[k];
    if (height < 64 && (x >> height)) *bad = 1;
    if (k == 0) return -1;
    uint64_t y = idx[k - 1];
    if (y >= x) { *bad = 1; return -1; }
    uint64_t d = x ^ y;
    int msb = 0;
    while (d >>= 1) msb++;
    return msb;
}

// ------------------------------------------------------------------------------------------------
// leaf derivation (build_leaf_nodes + shuffle_index, /root/reference/src/dapol/mod.rs:323-441)
DAPOL_HD_INLINE uint64_t seed_to_index(const uint32_t seed[8], int height) {
    // u64::from_be_bytes(seed[..8]) >> (64 - height)   (mod.rs:427-431)
    uint32_t w0 = seed[0], w1 = seed[1];
    uint64_t hi = ((uint64_t)(w0 & 0xff) << 24) | ((uint64_t)((w0 >> 8) & 0xff) << 16) | ((uint64_t)((w0 >> 16) & 0xff) << 8) | (w0 >> 24);
    uint64_t lo = ((uint64_t)(w1 & 0xff) << 24) | ((uint64_t)((w1 >> 8) & 0xff) << 16) | ((uint64_t)((w1 >> 16) & 0xff) << 8) | (w1 >> 24);
    uint64_t be = (hi << 32) | lo;
    return height >= 64 ? be : (be >> (64 - height));
}
// per user: audit_id = D(audit_seed || internal_id); index_seed = D(audit_id || "index_seed" || external_id);
// first candidate = D(index_seed); blinding = from_bits(D(audit_id || "blind_seed" || external_id)).
// Returns 0, or -1 if an input exceeds the single-chunk hashing limit.
DAPOL_HD_INLINE int derive_body(uint64_t i, int hash_id, const uint8_t *iid_blob, const uint64_t *iid_off, const uint8_t *eid_blob,
                                const uint64_t *eid_off, const uint8_t *audit_seed, uint32_t seed_len, int height, uint32_t *audit_out,
                                uint32_t *cur_seed, uint64_t *cand, uint32_t *blind_out) {
    dapol_hasher hs;
    uint32_t audit[8], iseed[8], bs[8];
    int rc = 0;
    hasher_init(hs, hash_id);
    hasher_update(hs, audit_seed, seed_len);
    hasher_update(hs, iid_blob + iid_off[i], (uint32_t)(iid_off[i + 1] - iid_off[i]));
    rc |= hasher_final(hs, audit);
    const uint8_t *eid = eid_blob + eid_off[i];
    uint32_t elen = (uint32_t)(eid_off[i + 1] - eid_off[i]);
    const uint8_t tag_i[10] = {'i', 'n', 'd', 'e', 'x', '_', 's', 'e', 'e', 'd'};
    const uint8_t tag_b[10] = {'b', 'l', 'i', 'n', 'd', '_', 's', 'e', 'e', 'd'};
    hasher_init(hs, hash_id);
    hasher_update_words(hs, audit, 8);
    hasher_update(hs, tag_i, 10);
    hasher_update(hs, eid, elen);
    rc |= hasher_final(hs, iseed);
    dapol_hash32(hash_id, iseed, iseed);  // first shuffle_index iteration (mod.rs:419-424)
    hasher_init(hs, hash_id);
    hasher_update_words(hs, audit, 8);
    hasher_update(hs, tag_b, 10);
    hasher_update(hs, eid, elen);
    rc |= hasher_final(hs, bs);
    bs[7] &= 0x7fffffffu;  // Scalar::from_bits (mod.rs:385)
    store8(audit_out + 8 * i, audit);
    store8(cur_seed + 8 * i, iseed);
    store8(blind_out + 8 * i, bs);
    cand[i] = seed_to_index(iseed, height);
    return rc;
}
// a user that lost its candidate slot to an earlier user re-hashes its seed (mod.rs:416-437); tries counts
// the candidates consumed so far (1 after derive_body); returns 0 if the 128 tries are exhausted.
DAPOL_HD_INLINE int rehash_body(uint64_t u, int hash_id, int height, uint32_t *cur_seed, uint64_t *cand, uint32_t *tries) {
    if (tries[u] >= 128) return 0;
    uint32_t s[8];
    load8(s, cur_seed + 8 * u);
    dapol_hash32(hash_id, s, s);
    store8(cur_seed + 8 * u, s);
    cand[u] = seed_to_index(s, height);
    tries[u] += 1;
    return 1;
}

// Opt-in leaf hash of the DAPOL+ paper (NOT the reference's bytes, which hash the commitment only, node.rs:33-36):
//   salt = D(audit_id || "salt_seed" || external_id),  leaf hash = D("leaf" || external_id || salt).
// Keep this a function: with the two tag arrays declared directly inside the __global__ kernel, nvcc 12.9 (sm_100a, -O3) produced
// wrong digests for every input while the very same statements inside a device function are right (tools/dbg/idsalt8.cu runs
// both shapes side by side on the device; found when tests/test_gpu_ids.py went red without a source change in this path).
DAPOL_HD_INLINE int leaf_id_hash_body(uint64_t i, int hash_id, const uint32_t *audit, const uint8_t *eid_blob, const uint64_t *eid_off, uint32_t *out) {
    const uint8_t tag_s[9] = {'s', 'a', 'l', 't', '_', 's', 'e', 'e', 'd'};
    const uint8_t tag_l[4] = {'l', 'e', 'a', 'f'};
    const uint8_t *eid = eid_blob + eid_off[i];
    const uint32_t elen = (uint32_t)(eid_off[i + 1] - eid_off[i]);
    dapol_hasher hs;
    uint32_t a[8], salt[8], h[8];
    load8(a, audit + 8 * i);
    hasher_init(hs, hash_id);
    hasher_update_words(hs, a, 8); hasher_update(hs, tag_s, 9); hasher_update(hs, eid, elen);
    int rc = hasher_final(hs, salt);
    hasher_init(hs, hash_id);
    hasher_update(hs, tag_l, 4); hasher_update(hs, eid, elen); hasher_update_words(hs, salt, 8);
    rc |= hasher_final(hs, h);
    store8(out + 8 * i, h);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Node passes.  Every commitment is carried as its HALF point Q (com = 2Q) in ns.ext: scalars are halved mod l
// before the comb and parents are Q_L + Q_R, so that compress(com) is the batched double-and-compress of
// ge25519.cuh (one shared inversion per thread batch instead of a 252-squaring chain per node).
// A thread handles up to B units t, t + stride, t + 2 stride, ... (coalesced across the warp).

// DapolNode::new(value, blinding) (node.rs:29-45): com = v*B + r*B_blinding by signed-window comb;
// hash = D(compress(com)) (node.rs:33-36).
template <int W, int B>
DAPOL_HD_INLINE void leaf_batch_body(uint64_t t, uint64_t stride, uint64_t n, const NodeStore &ns, uint64_t level_off, const uint32_t *pos,
                                     int hash_id, const uint64_t *values, const uint32_t *blind /*[n][8]*/, const ge_niels *tab_b,
                                     const ge_niels *tab_bbl) {
    constexpr int WV = comb_value_window<W>::value;
    constexpr int NWR = 253 / W + 1, NWV = 64 / WV + 1;
    ge_dc_batch<B> dc;
    dc.init();
#pragma unroll 1
    for (int b = 0; b < B; b++) {
        uint64_t i = t + (uint64_t)b * stride;
        if (i >= n) break;
        uint32_t rw[8];
        load8(rw, blind + 8 * i);
        sc rs, rh;
#pragma unroll
        for (int k = 0; k < 8; k++) rs.v[k] = rw[k];
        sc_half256(rh, rs);  // blinding may be unreduced (Scalar::from_bits, mod.rs:385)
        uint64_t v = values[i];
        uint32_t vw[2] = {(uint32_t)v, (uint32_t)(v >> 32)};
        int32_t d[NWR > NWV ? NWR : NWV];
        ge acc;
        ge_identity(acc);
        sc_signed_digits<WV, NWV>(d, vw, 2);  // tab_b holds multiples of B/2: v * (B/2) is the half point of v * B
        ge_comb_accumulate<WV, NWV, true>(acc, tab_b, d);
        sc_signed_digits<W, NWR>(d, rh.v, 8);
        ge_comb_accumulate<W, NWR>(acc, tab_bbl, d);
        uint64_t g = level_off + pos[i];
        ns.v[g] = v;
        store8(ns.r + 8 * g, rw);
        store_ge(ns.ext + 32 * g, acc);
        dc.push(acc);
    }
    dc.solve();
#pragma unroll 1
    for (int b = 0; b < dc.n; b++) {
        uint64_t g = level_off + pos[t + (uint64_t)b * stride];
        uint32_t cc[8], hh[8];
        dc.get(b, cc);
        dapol_hash32(hash_id, hh, cc);
        store8(ns.comc + 8 * g, cc);
        store8(ns.hash + 8 * g, hh);
    }
}

// mixed additions of the padding comb with inlined products (experiment, profiles/r02_variants.txt): the loop of 10 additions per
// node is where k_pad spends 70 % of its time, and every called product costs 24 register moves on the multiply pipe
#ifndef DAPOL_PAD_COMB_INL
#define DAPOL_PAD_COMB_INL false
#endif
// ChaCha stream of the padding node with creation ordinal g.  Stream mode: always 0.  Positional mode: the
// node's level inside the whole tree; ordinals run level H first, so level h owns [start[h], start[h - 1]).
struct PadStreams {
    int positional, levels, level0;                // levels = H; level0 = levels above this (sub)tree
    uint64_t start[DAPOL_MAX_TREE_HEIGHT + 2];     // start[h], h = 1 .. H; start[0] = number of padding nodes
};
DAPOL_HD_INLINE uint64_t pad_stream_of(const PadStreams &ps, uint64_t g) {
    if (!ps.positional) return 0;
    int lo = 1, hi = ps.levels;  // smallest h with start[h] <= g
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ps.start[mid] <= g) hi = mid; else lo = mid + 1;
    }
    return (uint64_t)(ps.level0 + lo);
}
// DapolNode::padding (node.rs:86-88): new(0, Scalar::random(rng)); rng draw #g of the seeded stream =
// from_bytes_mod_order_wide(ChaCha20(pad_seed) block pad_rng[g])   (RNG contract)
template <int W, int B>
DAPOL_HD_INLINE void pad_batch_body(uint64_t t, uint64_t stride, uint64_t n, const NodeStore &ns, const uint64_t *pad_dest, int hash_id,
                                    const uint32_t seed[8], const uint64_t *pad_rng, const ge_niels *tab_bbl, const PadStreams &ps) {
    constexpr int NWR = 253 / W + 1;
    ge_dc_batch<B> dc;
    dc.init();
#pragma unroll 1
    for (int b = 0; b < B; b++) {
        uint64_t g = t + (uint64_t)b * stride;
        if (g >= n) break;
        uint32_t ks[16];
        chacha20_block(ks, seed, pad_rng[g], pad_stream_of(ps, g));
        sc r, rh;
        sc_from_wide_with_half(r, rh, ks);
        int32_t d[NWR];
        sc_signed_digits<W, NWR>(d, rh.v, 8);
        ge acc;
        ge_identity(acc);
        ge_comb_accumulate<W, NWR, true, DAPOL_PAD_COMB_INL>(acc, tab_bbl, d);
        uint64_t dest = pad_dest[g];
        ns.v[dest] = 0;
        store8(ns.r + 8 * dest, r.v);
        store_ge(ns.ext + 32 * dest, acc);
        dc.push(acc);
    }
    dc.solve();
#pragma unroll 1
    for (int b = 0; b < dc.n; b++) {
        uint64_t dest = pad_dest[t + (uint64_t)b * stride];
        uint32_t cc[8], hh[8];
        dc.get(b, cc);
        dapol_hash32(hash_id, hh, cc);
        store8(ns.comc + 8 * dest, cc);
        store8(ns.hash + 8 * dest, hh);
    }
}

// Subtree-root record exchanged between shards: 58 LE words = half point X,Y,Z,T (32) | compress(com) (8) |
// hash (8) | blinding (8) | value (2).  The leaves of the top tree are these records instead of fresh commitments.
#define DAPOL_RECORD_WORDS 58
DAPOL_HD_INLINE void record_leaf_body(uint64_t i, const NodeStore &ns, uint64_t level_off, const uint32_t *pos, const uint32_t *recs) {
    const uint32_t *rec = recs + (uint64_t)DAPOL_RECORD_WORDS * i;
    uint64_t g = level_off + pos[i];
    for (int k = 0; k < 32; k++) ns.ext[32 * g + k] = rec[k];
    for (int k = 0; k < 8; k++) {
        ns.comc[8 * g + k] = rec[32 + k];
        ns.hash[8 * g + k] = rec[40 + k];
        ns.r[8 * g + k] = rec[48 + k];
    }
    ns.v[g] = (uint64_t)rec[56] | ((uint64_t)rec[57] << 32);
}

// Mergeable::merge (node.rs:64-80) in three data-parallel steps, so that the compressions of ALL internal nodes form one
// perfectly balanced batch instead of one latency-bound batch per level (the sums need no compressed point; only the
// parent HASH needs the children's):
//   1. per level, bottom-up:  v, r, com = sums           (merge_sum_body; one full addition per parent)
//   2. once:                  compress(com) of every internal node, one shared inversion per thread batch
//   3. per level, bottom-up:  hash = D(C(L)||C(R)||H(L)||H(R))   (merge_hash_body; HBM-bound)
// parent_pos == nullptr: the single root at slot parent_off.
DAPOL_HD_INLINE void merge_sum_body(uint64_t j, const NodeStore &ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos) {
    uint64_t dest = parent_pos ? parent_off + parent_pos[j] : parent_off;
    uint64_t l = child_off + 2 * j, r = l + 1;
    ns.v[dest] = ns.v[l] + ns.v[r];  // u64 wrapping add, as release-mode Rust
    sc a, c, s;
    load8(a.v, ns.r + 8 * l); load8(c.v, ns.r + 8 * r);
    sc_reduce256(a, a); sc_reduce256(c, c);  // leaf blindings may be unreduced (Scalar::from_bits)
    sc_add(s, a, c);
    store8(ns.r + 8 * dest, s.v);
    ge p, q, sum;
    load_ge(p, ns.ext + 32 * l); load_ge(q, ns.ext + 32 * r);
    ge_add(sum, p, q);
    store_ge(ns.ext + 32 * dest, sum);
}
DAPOL_HD_INLINE void merge_hash_body(uint64_t j, const NodeStore &ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos,
                                     int hash_id) {
    uint64_t dest = parent_pos ? parent_off + parent_pos[j] : parent_off;
    uint64_t l = child_off + 2 * j, r = l + 1;
    uint32_t cl[8], cr[8], hl[8], hr[8], hh[8];
    load8(cl, ns.comc + 8 * l); load8(cr, ns.comc + 8 * r);
    load8(hl, ns.hash + 8 * l); load8(hr, ns.hash + 8 * r);
    dapol_hash128(hash_id, hh, cl, cr, hl, hr);
    store8(ns.hash + 8 * dest, hh);
}
// D = Blake2b (64-byte digests, src/tests.rs:104-105; new_blank + build only, mod.rs:101-103).  The IMAD-bound leaf / padding
// kernels stay as they are (they leave a 32-byte placeholder in ns.hash); one extra pass hashes the compressed commitment of
// every leaf-level and padding node -- DapolNode::new: D(compress(com)), node.rs:33-36 -- and the parents' hashes take the
// 192-byte form of Mergeable::merge (node.rs:66-70).  The halves of a digest live in ns.hash / ns.hash_hi.
DAPOL_HD_INLINE void leafpad_hash_b2b_body(uint64_t g, const NodeStore &ns, uint64_t leaf_level_off) {
    if (g < leaf_level_off && !ns.is_pad[g]) return;  // internal node: hashed by its merge
    uint32_t cc[8], lo[8], hi[8];
    load8(cc, ns.comc + 8 * g);
    dapol_b2b_hash32(lo, hi, cc);
    store8(ns.hash + 8 * g, lo);
    store8(ns.hash_hi + 8 * g, hi);
}
DAPOL_HD_INLINE void merge_hash_b2b_body(uint64_t j, const NodeStore &ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos) {
    uint64_t dest = parent_pos ? parent_off + parent_pos[j] : parent_off;
    uint64_t l = child_off + 2 * j, r = l + 1;
    uint32_t cl[8], cr[8], hl[8], hlh[8], hr[8], hrh[8], lo[8], hi[8];
    load8(cl, ns.comc + 8 * l); load8(cr, ns.comc + 8 * r);
    load8(hl, ns.hash + 8 * l); load8(hr, ns.hash + 8 * r);
    load8(hlh, ns.hash_hi + 8 * l); load8(hrh, ns.hash_hi + 8 * r);
    dapol_b2b_hash192(lo, hi, cl, cr, hl, hlh, hr, hrh);
    store8(ns.hash + 8 * dest, lo);
    store8(ns.hash_hi + 8 * dest, hi);
}
// Internal (non-leaf, non-padding) nodes of all levels as one flat unit range: unit u of level h (start[h] <= u <
// start[h + 1], h = 0 .. H - 1) is the (u - start[h])-th real node of that level.
struct InternalMap {
    uint64_t start[DAPOL_MAX_TREE_HEIGHT + 2];
    int levels;                   // H
    const uint64_t *level_off;    // [H + 1]
    uint32_t *const *pos;         // [H + 1], pos[0] unused
};
DAPOL_HD_INLINE uint64_t internal_node_of(const InternalMap &m, uint64_t u) {
    int lo = 0, hi = m.levels - 1;
    while (lo < hi) {  // last level whose start <= u
        int mid = (lo + hi + 1) >> 1;
        if (m.start[mid] <= u) lo = mid; else hi = mid - 1;
    }
    uint64_t j = u - m.start[lo];
    return m.level_off[lo] + (lo ? (uint64_t)m.pos[lo][j] : 0);
}
template <int B>
DAPOL_HD_INLINE void compress_internal_body(uint64_t t, uint64_t stride, uint64_t n, const NodeStore &ns, const InternalMap &m) {
    ge_dc_batch<B> dc;
    uint64_t dest[B];
    dc.init();
#pragma unroll 1
    for (int b = 0; b < B; b++) {
        uint64_t u = t + (uint64_t)b * stride;
        if (u >= n) break;
        uint64_t g = internal_node_of(m, u);
        dest[b] = g;
        ge p;
        load_ge(p, ns.ext + 32 * g);
        dc.push(p);
    }
    dc.solve();
#pragma unroll 1
    for (int b = 0; b < dc.n; b++) {
        uint32_t cc[8];
        dc.get(b, cc);
        store8(ns.comc + 8 * dest[b], cc);
    }
}

// comb table entry (k, e): (e+1) * 2^(W k) * P in affine Niels form; P = B/2 (which = 0, 64/W + 1 windows: values are
// 64-bit) or B_blinding (which = 1, 253/W + 1 windows, used with halved scalars)
template <int W>
DAPOL_HD_INLINE void comb_table_body(uint64_t t, ge_niels *table, int nw, int which /*0 = B, 1 = B_blinding*/) {
    uint32_t half = 1u << (W - 1);
    uint32_t k = (uint32_t)(t / half), e = (uint32_t)(t % half);
    if ((int)k >= nw) return;
    ge base;
    if (which == 0) ge_basepoint_half(base); else ge_bblinding(base);
#pragma unroll 1
    for (uint32_t i = 0; i < k * W; i++) ge_dbl(base, base);
    // (e+1) * base by left-to-right double-and-add
    uint32_t m = e + 1;
    ge acc = base;
    int top = 31;
    while (!((m >> top) & 1u)) top--;
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        ge_dbl(acc, acc);
        if ((m >> b) & 1u) ge_add(acc, acc, base);
    }
    ge_niels n;
    ge_to_niels(n, acc);
    table[t] = n;
}

// The same table built a RUN of consecutive multiples per thread: entry e0 + i of window k = (e0 + 1) * base_k + i * base_k,
// one addition per entry and one inversion per RUN entries (the wide HBM-resident windows have 10^7 .. 10^8 entries).
// bases[k] = 2^(W k) * P (comb_bases_body).  Entries are stored canonical, so both builders give identical bytes.
template <int W>
DAPOL_HD_INLINE void comb_bases_body(ge *bases, int nw, int which) {
    ge base;
    if (which == 0) ge_basepoint_half(base); else ge_bblinding(base);
    for (int k = 0; k < nw; k++) {
        bases[k] = base;
#pragma unroll 1
        for (int i = 0; i < W; i++) ge_dbl(base, base);
    }
}
template <int W, int RUN>
DAPOL_HD_INLINE void comb_table_run_body(uint64_t t, ge_niels *table, int nw, const ge *bases) {
    constexpr uint64_t half = 1ull << (W - 1);
    static_assert(half % RUN == 0, "run length must divide the window size");
    constexpr uint64_t runs = half / RUN;
    uint32_t k = (uint32_t)(t / runs);
    uint64_t e0 = (t % runs) * RUN;
    if ((int)k >= nw) return;
    ge base = bases[k];
    ge_cached cb;
    ge_to_cached(cb, base);
    uint64_t m = e0 + 1;  // first multiple by left-to-right double-and-add
    ge acc = base;
    int top = 63;
    while (!((m >> top) & 1ull)) top--;
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        ge_dbl(acc, acc);
        if ((m >> b) & 1ull) ge_cadd(acc, acc, cb, 0);
    }
    fe X[RUN], Y[RUN], Z[RUN], pre[RUN], prod;  // local memory; Montgomery's trick over the RUN values of Z
    fe_set1(prod);
#pragma unroll 1
    for (int i = 0; i < RUN; i++) {
        if (i) ge_cadd(acc, acc, cb, 0);
        X[i] = acc.X; Y[i] = acc.Y; Z[i] = acc.Z;
        pre[i] = prod;
        fe_mul(prod, prod, acc.Z);  // Z != 0 on the curve
    }
    fe inv;
    fe_invert(inv, prod);
#pragma unroll 1
    for (int i = RUN - 1; i >= 0; i--) {
        fe zi, x, y;
        fe_mul(zi, inv, pre[i]);
        fe_mul(inv, inv, Z[i]);
        fe_mul(x, X[i], zi); fe_mul(y, Y[i], zi);
        ge_niels n;
        fe_add(n.ypx, y, x); fe_sub(n.ymx, y, x);
        fe_mul(n.t2d, x, y); fe_mul(n.t2d, n.t2d, fe_const_d2());
        uint32_t w[8];
        fe_canon(w, n.ypx); fe_fromwords(n.ypx, w);
        fe_canon(w, n.ymx); fe_fromwords(n.ymx, w);
        fe_canon(w, n.t2d); fe_fromwords(n.t2d, w);
        table[(uint64_t)k * half + e0 + i] = n;
    }
}

// path extraction (Dapol::generate_proof's get_merkle_path_ref_batch, mod.rs:173): siblings leaf level
// first.  leaf_pos = slot of the leaf inside level H.  pos_maps[h] = slot map of level h's real nodes.
struct PathOut {
    uint64_t *v;      // [K][H]
    uint32_t *r;      // [K][H][8]
    uint32_t *comc;   // [K][H][8]
    uint32_t *hash;   // [K][H][8]
};
DAPOL_HD_INLINE int64_t find_leaf_slot(const uint64_t *lvl_idx, uint64_t n, uint64_t leaf_idx) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (lvl_idx[mid] < leaf_idx) lo = mid + 1; else hi = mid;
    }
    return (lo < n && lvl_idx[lo] == leaf_idx) ? (int64_t)lo : -1;
}
