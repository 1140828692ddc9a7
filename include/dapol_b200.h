/* dapol_b200 -- C ABI of the B200-native DAPOL+ hot path (tree build, inclusion proofs, range proofs).
 *
 * Drop-in boundary for the reference crate MystenLabs/dapol (/root/reference): each entry point
 * names the reference interface it replaces.  The reference is generic Rust over a digest D and a
 * range-proof policy R; here D is `hash_id` and R is `policy`.  BLAKE3 and BLAKE2s-256 (32-byte digests) work on every
 * entry point; Blake2b (64-byte digests, the other half of the reference's test matrix, src/tests.rs:104-105) works where
 * the reference allows it: trees built from ready nodes (new_blank + build), their paths, proofs and verification --
 * Dapol::new insists on 32 bytes (src/dapol/mod.rs:101-103), so the liability / sharded builds return
 * DAPOL_ERR_INVALID_DIGEST_SIZE for it.  Wherever a buffer holds node hashes, one hash is dapol_digest_len(hash_id) bytes.
 *
 * Conventions: plain pointers and sizes, caller-allocated little-endian flat buffers, `int` return
 * (0 = ok; 1..5 mirror src/errors.rs DapolError variants; >= 16 are boundary errors), never unwinds.
 * Handles are opaque and owned by the library.  A tree handle is immutable after build.
 * There is NO CPU fallback: every compute entry point fails with DAPOL_ERR_CUDA when no sm_100
 * device is usable.
 *
 * Randomness contract (the reference draws from thread_rng(), src/dapol/node.rs:87 and
 * bulletproofs' prove_*; bit-exactness is defined against an injected stream):
 *   ChaCha20 (rand_chacha::ChaCha20Rng layout: 64-bit block counter, 64-bit stream id);
 *   k-th Scalar::random = from_bytes_mod_order_wide(keystream block k).
 *   Padding nodes: key pad_seed, stream 0, block pad_base + creation ordinal (level H..1, left to right).
 *   Range proof #q of the inclusion proof for leaf x (dapol_prove_batch / dapol_prove_to_file): key =
 *     BLAKE3("dapol-b200 prover nonce key v1" || seed || root commitment || root hash || le64 policy ||
 *            le64 aggregation_factor || le64 tree height),  stream x, blocks (q << 32) + draw#
 *     (bulletproofs' party / dealer draw order, 132 m draws per proof at 64 bits).  The key binds the caller's seed to the
 *     tree (whose root commits to every witness), the policy and the aggregation factor: re-using one seed for the next
 *     audit's tree or for another policy / factor never pairs a nonce with two different witnesses, and the prover streams
 *     are domain-separated from the padding stream even if the same 32 bytes are passed as pad_seed and seed.
 *     A Rust-side comparison seeds rand_chacha::ChaCha20Rng with that key, set_stream(x), set_word_pos(q << 36).
 *   The raw range-proof entry points (dapol_rangeproof_prove_batch*) take (seed, stream, base block) as given: the caller
 *     must never prove two different witnesses under the same triple.
 * SECRETS: seed and pad_seed must be secret, high-entropy and known only to the prover; the blindings they generate hide the
 * liabilities.  Table look-ups and Straus windows are indexed by secret scalars (blindings, nonces): this library is NOT
 * constant-time -- run it on hardware the prover controls (the reference's dalek code is constant-time for secrets).
 */
#ifndef DAPOL_B200_H
#define DAPOL_B200_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DAPOL_API __attribute__((visibility("default")))
#else
#define DAPOL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: 1..5 = src/errors.rs:5-17 in declaration order */
#define DAPOL_OK 0
#define DAPOL_ERR_TREE_HEIGHT_TOO_BIG 1
#define DAPOL_ERR_SPARSITY_TOO_SMALL 2
#define DAPOL_ERR_INVALID_DIGEST_SIZE 3
#define DAPOL_ERR_DUPLICATED_INTERNAL_ID 4
#define DAPOL_ERR_FAILED_TO_MAP_INDEX 5
#define DAPOL_ERR_BAD_ARG 16     /* reference: panic (slice OOB, unsorted input, ...) */
#define DAPOL_ERR_NOT_FOUND 17   /* reference: Option::None from generate_proof* */
#define DAPOL_ERR_BUFFER 18      /* caller buffer too small; required size is reported */
#define DAPOL_ERR_CUDA 19        /* no device / CUDA runtime failure (see dapol_last_cuda_error) */
#define DAPOL_ERR_DECODE 20      /* smtree DecodingError / bulletproofs ProofError::FormatError */
#define DAPOL_ERR_IO 21          /* tree / proof file cannot be opened, read or written, or is not a tree file */

#define DAPOL_HASH_BLAKE3 0   /* blake3::Hasher     (benches/dapol.rs:38, src/tests.rs:102-103) */
#define DAPOL_HASH_BLAKE2S 1  /* blake2::Blake2s    (src/dapol/tests.rs:13,21) */
#define DAPOL_HASH_BLAKE2B 2  /* blake2::Blake2b, 64-byte digests (src/tests.rs:104-105); ready-node trees only */

#define DAPOL_POLICY_PADDING 0    /* RangeProofPadding   (src/range/padding.rs) */
#define DAPOL_POLICY_SPLITTING 1  /* RangeProofSplitting (src/range/splitting.rs) */

#define DAPOL_MAX_TREE_HEIGHT 64  /* src/dapol/mod.rs:26 */

typedef struct dapol_ctx dapol_ctx;
typedef struct dapol_tree dapol_tree;

/* Context = one CUDA device + its stream + the precomputed generator tables
 * (PedersenGens::default(), src/dapol/node.rs:31; BulletproofGens::new, src/range/mod.rs:50,66 --
 * the reference re-derives them on every call, here once).  comb_window = window of the fixed-base tables of B/2 and
 * B_blinding (4, 8, 12, 15, 16: L2-resident; 20, 22, 24: up to 9.5 GB in HBM, fewer additions per commitment; 26: 32.8 GB, one
 * addition fewer again -- for a device dedicated to tree building); 0 picks 24 when at least 48 GB of device memory are free
 * and the tables can be allocated, else 15.  Every window gives the same bytes. */
DAPOL_API int dapol_ctx_create(int device, int comb_window, dapol_ctx **out);
DAPOL_API void dapol_ctx_destroy(dapol_ctx *ctx);  /* destroy the context's trees first: a tree holds a pointer to its context */
/* Run this context's kernels and copies on a caller-owned CUDA stream (cudaStream_t), e.g. the framework's
 * current stream, so the caller's events bracket the work.  The stream must outlive the context. */
DAPOL_API int dapol_ctx_set_stream(dapol_ctx *ctx, void *cuda_stream);
/* Where a padding node's blinding comes from (all builds on this context; default DAPOL_PADDING_STREAM).
 * STREAM: the reference's behaviour under the seeded-RNG contract -- the k-th padding node smtree creates draws block
 *   pad_base + k of stream 0 of ChaCha20(pad_seed) (DapolNode::padding ignores idx and secret, src/dapol/node.rs:85-88).
 * POSITIONAL (opt-in, NOT the reference's bytes): the padding node at (level h, index i) of the whole tree
 *   draws block i of stream h of ChaCha20(pad_seed) -- padding(idx, secret) as a function of its arguments, which the
 *   reference leaves as a TODO.  pad_seed is then the padding key derived from the secret and pad_base is ignored; shards need
 *   no exchange of padding counts: dapol_tree_build_shard_dev takes pad_level_base[0] = number of levels above the shard and
 *   pad_level_base[h] = prefix << h (index of the shard's first node of level h inside the whole tree). */
#define DAPOL_PADDING_STREAM 0
#define DAPOL_PADDING_POSITIONAL 1
DAPOL_API int dapol_ctx_set_padding_mode(dapol_ctx *ctx, int mode);
/* What a LEAF's hash is, for trees built from liabilities on this context (default DAPOL_LEAF_HASH_COMMITMENT).
 * COMMITMENT: the reference -- hash = D(compress(com)) (src/dapol/node.rs:33-36).
 * ID_SALT (opt-in, NOT the reference's bytes): the leaf hash of the DAPOL+ paper,
 *   salt = D(audit_id || "salt_seed" || external_id),  hash = D("leaf" || external_id || salt),
 *   so a leaf's hash binds the user's id and a per-user salt.  Padding and internal nodes are unchanged; proofs and their
 *   verification carry (com, hash) of the leaf as before.  Not available in the sharded builds (DAPOL_ERR_BAD_ARG). */
#define DAPOL_LEAF_HASH_COMMITMENT 0
#define DAPOL_LEAF_HASH_ID_SALT 1
DAPOL_API int dapol_ctx_set_leaf_hash_mode(dapol_ctx *ctx, int mode);
DAPOL_API int dapol_digest_len(int hash_id); /* bytes of one node hash: 32, 64 for DAPOL_HASH_BLAKE2B, 0 for an unknown id */
DAPOL_API const char *dapol_strerror(int code);
DAPOL_API const char *dapol_last_cuda_error(void);

/* Dapol::new_blank(height, _) + Dapol::build(&items, &secret)   (src/dapol/mod.rs:196-208; the path
 * benches/dapol.rs:149-158 times).  Items = DapolNode::new(values[i], blindings[i]) at
 * TreeIndex::from_u64(height, leaf_idx[i]); leaf_idx strictly increasing.  Host buffers. */
DAPOL_API int dapol_tree_build_from_nodes(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *leaf_idx,
                                const uint64_t *values, const uint8_t *blindings /* n*32 */, const uint8_t pad_seed[32],
                                uint64_t pad_base, dapol_tree **out);
/* Same with the three input arrays already resident in device memory (HBM). */
DAPOL_API int dapol_tree_build_from_nodes_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *d_leaf_idx,
                                    const uint64_t *d_values, const uint8_t *d_blindings, const uint8_t pad_seed[32],
                                    uint64_t pad_base, dapol_tree **out);

/* Dapol::update(&idx, node, &secret)   (src/dapol/mod.rs:210-213 -> smtree SparseMerkleTree::update), for a batch:
 * *out = the tree over the leaves of `tree` with the k given leaves inserted, a leaf whose index is already in the tree being
 * replaced (leaf_idx strictly increasing; host buffers).  `tree` itself is not modified (handles are immutable) -- destroy it
 * when the new one replaces it.  The union of the two sorted leaf lists is formed on the device and the level-synchronous build
 * runs over it, so one call costs a build (26 ms at 2^20 leaves) whatever k is: batch the updates.  The result is the tree
 * dapol_tree_build_from_nodes gives for the merged leaves with the same pad_seed / pad_base -- in the positional padding mode
 * that makes update and build agree bit for bit; in the stream mode pass a pad_base past the blocks already drawn (num_padding),
 * so that no padding blinding is used twice (the reference draws fresh thread_rng() values either way, and its own test compares
 * the two roots by value only, src/tests.rs:47, src/dapol/node.rs:115-120).  Trees built from liabilities lose their id -> index
 * map (the new leaves have no ids).  DAPOL_ERR_BAD_ARG for a shard with a top tree attached, a height-0 tree, or a tree whose leaves
 * carry id / salt hashes (DAPOL_LEAF_HASH_ID_SALT: they cannot be recomputed from value and blinding). */
DAPOL_API int dapol_tree_update(const dapol_tree *tree, uint64_t k, const uint64_t *leaf_idx, const uint64_t *values,
                                const uint8_t *blindings /* k*32 */, const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out);

/* Dapol::new(liabilities, options)   (src/dapol/mod.rs:100-128): argument checks, build_leaf_nodes
 * (mod.rs:323-399: audit_id / index_seed / shuffle_index / blind_seed), sort, build.
 * Ids are concatenated in a blob with n+1 offsets.  On DUPLICATED_INTERNAL_ID / FAILED_TO_MAP_INDEX
 * *err_pos receives the input position of the offending liability. */
DAPOL_API int dapol_tree_build_from_liabilities(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *iid_blob,
                                      const uint64_t *iid_off, const uint8_t *eid_blob, const uint64_t *eid_off,
                                      const uint64_t *values, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                      const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out, uint64_t *err_pos);

/* Same with ids, offsets and values already resident in device memory (audit_seed stays a host pointer). */
DAPOL_API int dapol_tree_build_from_liabilities_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob,
                                          const uint64_t *d_iid_off, const uint8_t *d_eid_blob, const uint64_t *d_eid_off,
                                          const uint64_t *d_values, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                          const uint8_t pad_seed[32], uint64_t pad_base, dapol_tree **out, uint64_t *err_pos);

/* ---- sharded build (one process per GPU).  The tree of 2^k shards splits at level k: shard r owns the
 * leaves whose index starts with the k-bit prefix r and builds the height-(H-k) subtree below node r of level k; the
 * 2^k subtree roots are exchanged as 232-byte records and every shard builds the top k levels itself.  The reference
 * has no counterpart (it is single-threaded); the result is bit-identical with the single-tree build
 * (dapol_tree_build_from_liabilities) for the same pad_seed.
 *
 * build_leaf_nodes (src/dapol/mod.rs:323-399) in two stages.  Stage 1, per slice of the liabilities (per-user hashing,
 * mod.rs:338-386): audit id, index-seed state after the first shuffle_index iteration, first candidate index, blinding.
 * All outputs are device arrays of n elements (32 B, 32 B, u64, 32 B). */
DAPOL_API int dapol_leaves_derive_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint8_t *d_iid_blob, const uint64_t *d_iid_off,
                                      const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint8_t *audit_seed, uint64_t audit_seed_len,
                                      uint8_t *d_audit, uint8_t *d_seed_state, uint64_t *d_cand, uint8_t *d_blind);
/* Stage 2 over the records of ALL n_total users in input order (slices concatenated): duplicate-id check (mod.rs:345-349),
 * shuffle_index collision rule (mod.rs:408-441), sort by index (mod.rs:396).  d_seed_state and d_cand are updated in
 * place (d_cand[i] = final leaf index of user i = id_to_idx_map).  Outputs the sorted leaves whose index starts with the
 * prefix_bits-bit `prefix`, indexes with the prefix stripped: *n_out leaves (DAPOL_ERR_BUFFER if > cap, *n_out is set). */
DAPOL_API int dapol_leaves_assign_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n_total, const uint8_t *d_audit, uint8_t *d_seed_state,
                                      uint64_t *d_cand, const uint8_t *d_blind, const uint64_t *d_values, int prefix_bits, uint64_t prefix,
                                      uint64_t *d_out_idx, uint64_t *d_out_values, uint8_t *d_out_blind, uint64_t cap, uint64_t *n_out,
                                      uint64_t *err_pos);
/* Padding nodes per level (counts[0..height], counts[0] = 0) of the tree over the given sorted leaves: the shards
 * all-gather these to place their padding draws in the single-tree creation order (level H..1, left to right). */
DAPOL_API int dapol_tree_level_pad_counts_dev(dapol_ctx *ctx, int height, uint64_t n, const uint64_t *d_leaf_idx, uint64_t *counts);
/* dapol_tree_build_from_nodes_dev for one shard: pad_level_base[h] (host, h = 0..height) = block of the seeded padding
 * stream drawn by the first padding node of the shard's level h. */
DAPOL_API int dapol_tree_build_shard_dev(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *d_leaf_idx,
                                         const uint64_t *d_values, const uint8_t *d_blindings, const uint8_t pad_seed[32],
                                         const uint64_t *pad_level_base, dapol_tree **out);
/* Root of a (sub)tree as the record the shards exchange: half point of the commitment X,Y,Z,T (128 B) | compressed
 * commitment (32) | hash (32) | blinding (32) | value (8, LE). */
#define DAPOL_RECORD_BYTES 232
DAPOL_API int dapol_tree_root_record(const dapol_tree *tree, uint8_t *rec /* DAPOL_RECORD_BYTES */);
/* Top tree: the leaves are n subtree-root records at (strictly increasing) indexes leaf_idx of level `height`; missing
 * subtrees are padded like any missing sibling, drawing from the padding stream at pad_base + creation ordinal. */
DAPOL_API int dapol_tree_build_from_records(dapol_ctx *ctx, int hash_id, int height, uint64_t n, const uint64_t *leaf_idx,
                                            const uint8_t *records /* n*DAPOL_RECORD_BYTES */, const uint8_t pad_seed[32], uint64_t pad_base,
                                            dapol_tree **out);
/* Make `tree` the subtree under node `prefix` of the leaf level of `top` (same context; `top` must outlive `tree`):
 * dapol_tree_paths / dapol_prove_batch then take whole-tree leaf indexes and emit whole-tree paths and proofs. */
DAPOL_API int dapol_tree_attach_top(dapol_tree *tree, const dapol_tree *top, uint64_t prefix);

/* ---- sharded build, ONE call per rank (what a multi-GPU host uses; the staged entry points above remain for hosts that
 * bring their own exchange).  Collectives go through a communicator: the built-in backend is NCCL (resolved with dlopen
 * at run time: the copy of libnccl.so.2 the process already holds, e.g. torch's, else the system one; no link-time
 * dependency), or the host plugs its own transport as two callbacks on DEVICE buffers.
 *   rank 0:      dapol_comm_nccl_unique_id(id);  -> ship the 128 bytes to every rank by any means (MPI, TCP, a file, ...)
 *   every rank:  dapol_comm_nccl_create(ctx, id, rank, world, &comm);
 *                dapol_sharded_build(ctx, comm, ..., &subtree, &top, ...);      (Dapol::new over all ranks' liabilities) */
typedef struct dapol_comm dapol_comm;
typedef struct dapol_comm_ops {
    void *user;
    /* every rank contributes `bytes` bytes at d_send; d_recv receives world * bytes in rank order.  Device pointers; the
     * operation is ordered on cuda_stream (enqueue it there, or synchronise the stream, do it, and return when it is done) */
    int (*all_gather)(void *user, const void *d_send, void *d_recv, uint64_t bytes, void *cuda_stream);
    /* rank r is sent send_bytes[r] bytes from d_send + send_off[r] and sends recv_bytes[r] bytes to d_recv + recv_off[r]
     * (host arrays of `world` entries; a rank's own block goes through the same call) */
    int (*all_to_all)(void *user, const void *d_send, const uint64_t *send_off, const uint64_t *send_bytes, void *d_recv,
                      const uint64_t *recv_off, const uint64_t *recv_bytes, void *cuda_stream);
} dapol_comm_ops;
#define DAPOL_NCCL_ID_BYTES 128
DAPOL_API int dapol_comm_create(int rank, int world, const dapol_comm_ops *ops, dapol_comm **out);
DAPOL_API int dapol_comm_nccl_unique_id(uint8_t id[DAPOL_NCCL_ID_BYTES]);
DAPOL_API int dapol_comm_nccl_create(dapol_ctx *ctx, const uint8_t id[DAPOL_NCCL_ID_BYTES], int rank, int world, dapol_comm **out);
DAPOL_API void dapol_comm_destroy(dapol_comm *comm);
DAPOL_API int dapol_comm_rank(const dapol_comm *comm);
DAPOL_API int dapol_comm_world(const dapol_comm *comm);
/* Dapol::new(liabilities, options) (src/dapol/mod.rs:100-128) where the liabilities are the concatenation of every rank's
 * slice in rank order (this rank passes ITS slice: device pointers, n_local may be 0) and the tree is split by the
 * log2(world)-bit leaf-index prefix.  Bit-identical with dapol_tree_build_from_liabilities of the whole input for the same
 * pad_seed / pad_base.  Per rank: 96 B per LOCAL user leave over the all-to-all (audit id for the duplicate check, claim =
 * candidate index + position + value + blinding); index collisions are resolved by the owner of the index prefix, only the
 * losers of a round travel again; then one all-gather of per-level padding counts and one of the 232-byte subtree roots.
 * Outputs: *subtree = this rank's height-(H - k) subtree with *top attached (NULL if no leaf fell into the rank's prefix;
 * world == 1: the whole tree, *top = NULL); *top = the top k levels (replicated); destroy subtree first, then top.
 * dapol_tree_leaf_index_of(*top, pos) answers for the LOCAL slice, pos in [*first_pos, *first_pos + n_local).
 * Errors are collective: every rank returns the same code (DUPLICATED_INTERNAL_ID / FAILED_TO_MAP_INDEX with *err_pos =
 * the earliest offending input position of the whole input).  phase_ms (may be NULL): device time of [0] hashing,
 * [1] exchange (duplicate check + claim rounds), [2] leaves + subtree build, [3] root gather + top tree. */
DAPOL_API int dapol_sharded_build(dapol_ctx *ctx, dapol_comm *comm, int hash_id, int height, uint64_t n_local, const uint8_t *d_iid_blob,
                                  const uint64_t *d_iid_off, const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint64_t *d_values,
                                  const uint8_t *audit_seed, uint64_t audit_seed_len, const uint8_t pad_seed[32], uint64_t pad_base,
                                  dapol_tree **subtree, dapol_tree **top, uint64_t *n_total, uint64_t *first_pos, uint64_t *err_pos,
                                  float phase_ms[4]);

DAPOL_API void dapol_tree_destroy(dapol_tree *tree);

/* Dapol::root_raw / root   (src/dapol/mod.rs:134-141): commitment (compressed), hash, value, blinding. */
DAPOL_API int dapol_tree_root(const dapol_tree *tree, uint8_t com[32], uint8_t *hash /* digest_len */, uint64_t *value, uint8_t blinding[32]);

/* Tree introspection for parity dumps: level 0 = root .. height = leaves; nodes of a level are in tree
 * order with the two children of a parent adjacent. */
DAPOL_API int dapol_tree_height(const dapol_tree *tree);
DAPOL_API int dapol_tree_hash_id(const dapol_tree *tree);  /* DAPOL_HASH_*, -1 for NULL */
DAPOL_API uint64_t dapol_tree_num_nodes(const dapol_tree *tree);
DAPOL_API uint64_t dapol_tree_num_padding(const dapol_tree *tree);
DAPOL_API uint64_t dapol_tree_level_size(const dapol_tree *tree, int level);
DAPOL_API int dapol_tree_level_copy(const dapol_tree *tree, int level, uint64_t *idx, uint64_t *values, uint8_t *blindings,
                          uint8_t *coms, uint8_t *hashes /* n*digest_len */, uint8_t *is_padding); /* any pointer may be NULL */
/* id -> TreeIndex map of Dapol::new (id_to_idx_map, mod.rs:80,389): leaf index of the i-th input liability */
DAPOL_API int dapol_tree_leaf_index_of(const dapol_tree *tree, uint64_t input_pos, uint64_t *leaf_idx);
/* ... and by the internal id itself: Dapol::generate_proof_for_id / generate_proof_batch_for_ids (mod.rs:148-165) =
 * dapol_tree_index_of[_batch] followed by dapol_prove_batch / dapol_generate_proof_batch.  The library keeps the audit ids
 * D(audit_seed || internal_id) of the liabilities it was built from (mod.rs:347-353) and sorts their prefixes on the first
 * lookup; a query hashes the id (any length) and binary-searches.  DAPOL_ERR_NOT_FOUND if the id (batch: ANY id, found[i]
 * says which) is not in the tree or the tree was built from ready nodes (reference: None).  On the handles of a sharded build
 * (*top of dapol_sharded_build) the map covers the rank's own slice of the liabilities. */
DAPOL_API int dapol_tree_index_of(const dapol_tree *tree, const uint8_t *internal_id, uint64_t len, uint64_t *leaf_idx);
DAPOL_API int dapol_tree_index_of_batch(const dapol_tree *tree, uint64_t k, const uint8_t *id_blob, const uint64_t *id_off /* k+1 */,
                                        uint64_t *leaf_idx /* k */, uint8_t *found /* k */);

/* smtree get_merkle_path_ref_batch for one leaf (src/dapol/mod.rs:173-184): the K x height siblings,
 * leaf level first, with their secret (value, blinding) and public (com, hash) parts. */
DAPOL_API int dapol_tree_paths(const dapol_tree *tree, uint64_t k, const uint64_t *leaf_idx, uint64_t *values /* k*h */,
                     uint8_t *blindings /* k*h*32 */, uint8_t *coms /* k*h*32 */, uint8_t *hashes /* k*h*digest_len */,
                     uint8_t *leaf_coms /* k*32 or NULL */, uint8_t *leaf_hashes /* k*digest_len or NULL */);

/* ---- persistence (the reference keeps the tree in memory only and leaves "write the proofs to a local
 * file" as a TODO, src/dapol/mod.rs:250).  dapol_tree_save writes the whole node store of a built tree (every level: index,
 * value, blinding, compressed commitment, hash, padding flag), the slot maps and the id -> leaf-index map to one
 * little-endian file ("DAPOLT01" header; ~113 B per node: 2.9 GB at 2^20 users / height 32); dapol_tree_load brings it back
 * into HBM on any context, bit-identical (levels, paths and inclusion proofs are those of the saved tree).  The file holds
 * the SECRET blindings of every node: protect it like the liabilities themselves.  A shard's attachment to its top tree is
 * not saved (re-attach with dapol_tree_attach_top). */
DAPOL_API int dapol_tree_save(const dapol_tree *tree, const char *path);
DAPOL_API int dapol_tree_load(dapol_ctx *ctx, const char *path, dapol_tree **out);
/* dapol_prove_batch for k leaves streamed to a file in chunks of `chunk` proofs (0 = 8192): proof i at byte i * size, no
 * header; the GPU works on one chunk while nothing larger than a chunk is held on the host.  *proof_size (if not NULL)
 * receives the size of one proof. */
DAPOL_API int dapol_prove_to_file(const dapol_tree *tree, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                  const uint8_t seed[32], uint64_t chunk, const char *path, uint64_t *proof_size);

/* ---- inclusion proofs.
 * Dapol::generate_proof_batch-per-leaf (src/dapol/mod.rs:167-190) for k leaves, each serialised as DapolProof::serialize
 * (src/proof/mod.rs:68-73) = R::serialize (padding.rs:40-53 / splitting.rs:38-59) || MerkleProof::serialize.  All k
 * proofs have the same size (dapol_inclusion_proof_size); proof i is written at out + i * size.  The range proofs of all
 * k leaves run as a few large GPU batches (one per aggregate size, one of singles).  DAPOL_ERR_NOT_FOUND if a leaf index
 * is not a leaf of the tree (reference: None), DAPOL_ERR_BAD_ARG if aggregation_factor > height (reference: panic),
 * DAPOL_ERR_BUFFER (with *proof_size set) if cap < k * size. */
DAPOL_API uint64_t dapol_inclusion_proof_size(int height, uint64_t aggregation_factor, int policy);  /* 32-byte digests */
DAPOL_API uint64_t dapol_inclusion_proof_size_d(int height, uint64_t aggregation_factor, int policy, int hash_id); /* sibling = com || hash: 32 + digest_len */
DAPOL_API int dapol_prove_batch(const dapol_tree *tree, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                const uint8_t seed[32], uint8_t *out, uint64_t cap, uint64_t *proof_size);
/* DapolProof::deserialize + verify (src/proof/mod.rs:41-47,76-95) for k proofs against one root: proof i is
 * proofs[offsets[i] .. offsets[i+1]), leaf i = DapolProofNode{leaf_coms[i], leaf_hashes[i]}.  ok[i] = 1 iff the Merkle
 * path folds to the root (DapolProofNode::merge, src/proof/node.rs:56-69) and R::verify accepts the siblings'
 * commitments (padding.rs:168-197 / splitting.rs:180-211).  Malformed bytes are a reject, never an error. */
DAPOL_API int dapol_verify_batch(dapol_ctx *ctx, int hash_id, int policy, uint64_t k, const uint8_t root_com[32], const uint8_t *root_hash /* digest_len */,
                                 const uint8_t *leaf_coms /* k*32 */, const uint8_t *leaf_hashes /* k*digest_len */, const uint8_t *proofs,
                                 const uint64_t *offsets /* k+1 */, uint8_t *ok /* k */);

/* ---- batch proofs: ONE DapolProof for several leaves.
 * Dapol::generate_proof_batch(&[TreeIndex]) (src/dapol/mod.rs:172-190): the Merkle part carries, level by level from the
 * leaves up and left to right, every sibling that is not itself on the way up from the batch (smtree
 * get_merkle_path_ref_batch; UPSTREAM-RECALL, as is the MerkleProof framing: be16 height, be64 #indexes, path bits of every
 * index, be64 #siblings, siblings), and R::generate_proof runs ONCE over all those siblings with aggregation_factor.
 * leaf_idx strictly increasing (smtree's batch order); k = 1 is dapol_prove_batch of one leaf, byte for byte (mod.rs:167-169).
 * Nonces: the tree's prover key chained over the batch's leaf indexes (BLAKE3, 64 indexes per link, label "dapol-b200 batch
 * proof nonce key v1"), stream 0, blocks (q << 32) + draw#.  DAPOL_ERR_BAD_ARG if aggregation_factor > #siblings (reference:
 * slice panic), if the indexes are not increasing, or if the tree is a shard with a top tree attached; DAPOL_ERR_NOT_FOUND if
 * an index is not a leaf (reference: None).  dapol_batch_proof_size: size of that proof (0 = bad arguments). */
DAPOL_API uint64_t dapol_batch_proof_size(int height, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy);
DAPOL_API uint64_t dapol_batch_proof_size_d(int height, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy, int hash_id);
DAPOL_API int dapol_generate_proof_batch(const dapol_tree *tree, uint64_t k, const uint64_t *leaf_idx, uint64_t aggregation_factor, int policy,
                                         const uint8_t seed[32], uint8_t *out, uint64_t cap, uint64_t *proof_size);
/* DapolProof::deserialize + verify_batch(&root, &leaves) (src/proof/mod.rs:49-54,76-95): *ok = 1 iff the k leaves (in index
 * order) and the proof's siblings fold to the root (MerkleProof::verify_batch by level-synchronous DapolProofNode::merge on the
 * device) and R::verify accepts the siblings' commitments.  Malformed bytes or a wrong number of leaves are a reject. */
DAPOL_API int dapol_proof_verify_batch(dapol_ctx *ctx, int hash_id, int policy, uint64_t k, const uint8_t root_com[32], const uint8_t *root_hash /* digest_len */,
                                       const uint8_t *leaf_coms /* k*32 */, const uint8_t *leaf_hashes /* k*digest_len */, const uint8_t *proof,
                                       uint64_t proof_len, uint8_t *ok);

/* ---- range proofs: src/range/mod.rs:48-119 generate_/verify_{single,aggregated}_range_proof in batches.
 * One call = k independent Bulletproofs of one shape: nbits in {8,16,32,64} (the reference fixes BIT_SIZE = 64,
 * range/mod.rs:16), m parties (power of two <= 64; m = 1 is prove_single / verify_single), each over a fresh
 * Transcript::new(&[]).  Proof bytes = RangeProof::to_bytes(): 32 * (9 + 2 log2(nbits m)) (672 for 64 x 1 =
 * SINGLE_PROOF_BYTE_NUM, range/mod.rs:18).  Randomness of proof i: ChaCha20(seed) stream streams[i], draw j at
 * block base_blocks[i] + j, in bulletproofs' draw order (see the contract at the top of this header).
 * The generator tables (BulletproofGens::new(64, m), range/mod.rs:50,66,85,104) are built on first use per context. */
DAPOL_API uint64_t dapol_rangeproof_size(int nbits, int m); /* 0 for an unsupported shape */
DAPOL_API int dapol_rangeproof_prove_batch(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint64_t *values /* k*m */,
                                           const uint8_t *blindings /* k*m*32, Scalar bytes, may be unreduced */, const uint8_t seed[32],
                                           const uint64_t *streams /* k */, const uint64_t *base_blocks /* k */, uint8_t *proofs /* k*size */);
/* ok[i] = 1 iff RangeProof::from_bytes + verify_multiple accept proof i for its m commitments; malformed input is a
 * reject (0), never an error. */
DAPOL_API int dapol_rangeproof_verify_batch(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint8_t *proofs /* k*proof_len */,
                                            uint64_t proof_len, const uint8_t *commitments /* k*m*32 */, uint8_t *ok /* k */);
/* Same with every array already resident in device memory (HBM). */
DAPOL_API int dapol_rangeproof_prove_batch_dev(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint64_t *d_values, const uint8_t *d_blindings,
                                               const uint8_t seed[32], const uint64_t *d_streams, const uint64_t *d_base_blocks, uint8_t *d_proofs);
DAPOL_API int dapol_rangeproof_verify_batch_dev(dapol_ctx *ctx, int nbits, int m, uint64_t k, const uint8_t *d_proofs, uint64_t proof_len,
                                                const uint8_t *d_commitments, uint8_t *d_ok);
/* How the verifier entry points of this context (dapol_rangeproof_verify_batch[_dev], dapol_verify_batch, dapol_proof_verify_batch)
 * check a batch.  group <= 1 (default): every proof on its own, its 17 .. 102 variable points by Straus' interleaving -- what
 * RangeProof::verify_multiple does per proof (src/range/mod.rs:83-119).  group = G > 1: the proofs are taken G at a time and each
 * group is checked by ONE random linear combination of its G verification equations (secret 253-bit weights, fresh per call
 * from the OS; weight_seed != NULL fixes the 32-byte ChaCha20 key instead, for reproducible tests): the shared generators cost
 * one table MSM per group, and the G * nv variable points are summed by the BUCKET METHOD (Pippenger: window_bits-bit signed
 * digits, one addition per point and window into its bucket, buckets folded by running sums; window_bits = 0 picks
 * log2(G * nv) - 3).  A group whose combination is not the identity, or that holds a malformed proof, is re-verified proof by
 * proof, so ok[] is exactly the per-proof verdict either way (a bad proof slips through a group with probability 2^-252).
 * Large groups pay when (almost) all proofs are valid -- an auditor checking its own output, C3 / the north-star run; with
 * many bad proofs keep G small or 0.  dapol_ctx_verify_fallbacks: proofs re-verified one by one so far. */
DAPOL_API int dapol_ctx_set_verify_mode(dapol_ctx *ctx, uint64_t group, int window_bits, const uint8_t *weight_seed /* 32 bytes or NULL */);
DAPOL_API uint64_t dapol_ctx_verify_fallbacks(const dapol_ctx *ctx);
/* window (8, 12, 13, 14, 15 or 16 bits) of the generator tables; drops tables already built.  0 (the default) = the widest
 * window whose tables fit 70 % of the free HBM, at most 128 GB (16 bits up to m = 8 ... 15 bits at m = 32, 14 at m = 64);
 * if the allocation fails after all, the next narrower window is tried. */
DAPOL_API int dapol_ctx_set_rangeproof_window(dapol_ctx *ctx, int window);
/* HBM budget (bytes) of the generator tables under window 0 (auto); 0 restores the default (70 % of the device memory that is free
 * or cached-but-unused by the stream-ordered pool, at most 128 GB).  A process that shares the GPU (another framework, other
 * contexts, large trees to come) sets it to what it can spare: e.g. 4 GB gives window 12 at m = 1 .. 8.  The shared default
 * memory pool is never trimmed up front; its unused cache is released only if an allocation that fits the budget fails.
 * dapol_ctx_rangeproof_table_bytes: size of the tables currently in place (0 = none built yet). */
DAPOL_API int dapol_ctx_set_rangeproof_table_budget(dapol_ctx *ctx, uint64_t bytes);
DAPOL_API uint64_t dapol_ctx_rangeproof_table_bytes(const dapol_ctx *ctx);
/* device time of the last range-proof batch on this ctx (ms): [0] total, [1] MSM passes, [2] other passes, [3] last table build */
DAPOL_API int dapol_rangeproof_last_times(const dapol_ctx *ctx, float ms[4]);
/* the same plus the MSM passes split per kernel class (CUDA events on the ctx stream; the roofline report of the dominant kernel):
 * [4] k_rp_p10 (L / R of the inner-product rounds over the generator tables), [5] k_rp_p3 (A, S), [6] hybrid late rounds
 * (k_rp_pm / pv / pf), [7] verifier (k_rp_v1 / v2) */
DAPOL_API int dapol_rangeproof_last_kernel_times(const dapol_ctx *ctx, float ms[8]);

/* Micro-entry points used by the parity tests and the roofline microbenchmark. */
DAPOL_API int dapol_commit_batch(dapol_ctx *ctx, uint64_t n, const uint64_t *values, const uint8_t *blindings, uint8_t *coms /* n*32 */);
DAPOL_API int dapol_imad_peak(dapol_ctx *ctx, int variant, double *gmac_per_s); /* measured 32x32->64 MAC/s, variant 0..3 */
DAPOL_API int dapol_fe_bench(dapol_ctx *ctx, int op, double *gop_per_s);        /* field mul(0)/sq(1) throughput, Gop/s */
DAPOL_API uint64_t dapol_kernel_launches(const dapol_ctx *ctx);                 /* kernels launched so far on this ctx */
/* tuning parameters in effect (for the work model of the roofline report): comb window of the tree tables, nodes per
 * thread sharing one field inversion, window of the range-proof generator tables.  Any pointer may be NULL. */
DAPOL_API int dapol_ctx_params(const dapol_ctx *ctx, int *comb_window, int *node_batch, int *rangeproof_window);
/* device-time of the last tree build on this ctx, split per phase (ms, CUDA events on the ctx stream):
 * [0] structure, [1] leaves, [2] padding, [3] merges, [4] total */
DAPOL_API int dapol_last_build_times(const dapol_ctx *ctx, float ms[5]);

#ifdef __cplusplus
}
#endif
#endif
