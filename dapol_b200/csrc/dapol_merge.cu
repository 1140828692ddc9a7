// Merge passes of the tree build, compiled with INLINED field products.
// Every warp of these kernels runs the same short loop in step (one node per thread in k_merge_sum, the same batch of 24 in
// k_compress_internal), so the instruction cache holds the loop once and a call per product is pure overhead on the multiply
// pipe (measured: merges 7.64 ms with called products, 7.18 ms inlined; profiles/r01d_variants.txt).  The node kernels
// (k_pad, k_leaf: warps in different phases of a 100 KB program) call their products -- see fe25519.cuh.
#define DAPOL_FE_CALL 0
#include <cuda_runtime.h>
#include <cstring>
#include "dapol_internal.h"

// merge step 1 (per level): value, blinding and point sums of the parents
__global__ void __launch_bounds__(128) k_merge_sum(uint64_t n, NodeStore ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) merge_sum_body(j, ns, child_off, parent_off, parent_pos);
}
// merge step 2 (once per tree): compress every internal node, B per shared inversion
template <int B>
__global__ void __launch_bounds__(128, DAPOL_MERGE_MINB) k_compress_internal(uint64_t n, uint64_t stride, NodeStore ns,
                                                                              const __grid_constant__ InternalMap m) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < stride) compress_internal_body<B>(t, stride, n, ns, m);
}
// merge step 3 (per level): parent hashes
__global__ void __launch_bounds__(128) k_merge_hash(uint64_t n, NodeStore ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos,
                                                    int hash_id) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) merge_hash_body(j, ns, child_off, parent_off, parent_pos, hash_id);
}
// D = Blake2b: 192-byte parent hashes over the two halves of the children's digests
__global__ void __launch_bounds__(128) k_merge_hash_b2b(uint64_t n, NodeStore ns, uint64_t child_off, uint64_t parent_off, const uint32_t *parent_pos) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) merge_hash_b2b_body(j, ns, child_off, parent_off, parent_pos);
}
void dapol_launch_merges(dapol_ctx *ctx, dapol_tree *t) {
    const int H = t->height;
    const int hash_id = t->hash_id;
    cudaStream_t st = ctx->stream;
    if (H >= 1) {
        for (int h = H; h >= 1; h--) {
            uint64_t np = t->n_real[h - 1];
            k_merge_sum<<<grid_for(np, 128), 128, 0, st>>>(np, t->ns, t->level_off[h], t->level_off[h - 1], h - 1 >= 1 ? t->pos[h - 1] : nullptr);
            ctx->launches++;
        }
        InternalMap im;
        memset(&im, 0, sizeof(im));
        im.levels = H; im.level_off = t->d_level_off; im.pos = t->d_pos;
        uint64_t n_int = 0;
        for (int h = 0; h < H; h++) { im.start[h] = n_int; n_int += t->n_real[h]; }
        im.start[H] = n_int;
        uint64_t stride = batch_stride(n_int, k_compress_internal<NODE_BATCH>, 2.0);
        k_compress_internal<NODE_BATCH><<<grid_for(stride, 128), 128, 0, st>>>(n_int, stride, t->ns, im);
        ctx->launches++;
        for (int h = H; h >= 1; h--) {
            uint64_t np = t->n_real[h - 1];
            if (hash_id == DAPOL_HASH_BLAKE2B)
                k_merge_hash_b2b<<<grid_for(np, 128), 128, 0, st>>>(np, t->ns, t->level_off[h], t->level_off[h - 1], h - 1 >= 1 ? t->pos[h - 1] : nullptr);
            else
                k_merge_hash<<<grid_for(np, 128), 128, 0, st>>>(np, t->ns, t->level_off[h], t->level_off[h - 1], h - 1 >= 1 ? t->pos[h - 1] : nullptr, hash_id);
            ctx->launches++;
        }
    }
}
