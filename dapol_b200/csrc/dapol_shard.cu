// One DAPOL+ tree over several GPUs, one process per GPU, behind ONE C-ABI call (include/dapol_b200.h
// "sharded build, one call").  The reference builds its tree on one CPU thread (src/dapol/mod.rs:100-128) and has no
// multi-device counterpart; the result here is bit-identical with the single-GPU build of the same liabilities.
//
// Exchange protocol (2^k ranks, rank r owns the leaf indexes with k-bit prefix r and the users of its input slice):
//   1. every rank hashes its slice (build_leaf_nodes stage 1, mod.rs:338-386);
//   2. duplicate internal ids (mod.rs:345-349 <=> equal audit ids): all-to-all of (audit id, input position) keyed by the
//      audit id, exact comparison inside runs of equal 64-bit prefixes on the receiver -- 40 B per user, once;
//   3. shuffle_index's first-come-first-served rule (mod.rs:408-441) as a distributed fix-point: every user sends its CLAIM
//      (candidate index, input position, value, blinding: 56 B) to the owner of the candidate's prefix; the owner sorts
//      its claims by (candidate, position), the earliest position keeps the slot, every other claimant is a LOSER whose
//      position goes back to its home rank (8 B), which re-hashes and sends a new claim.  A slot once claimed by position p
//      is held by a position <= p for good, so the fix-point equals the sequential rule.  After the first round only the
//      few hundred losers move;
//   4. the owner's surviving claims, sorted, ARE its leaves: no further exchange of values or blindings;
//   5. all-gather of the per-level padding counts (so each shard draws the blocks of the seeded stream the single-tree
//      creation order gives it), subtree build with no data-path collective, all-gather of the 2^k root records (232 B),
//      top k levels on every rank.
// Per rank this moves 96 B per LOCAL user (all-to-all) instead of all-gathering 112 B per user of the WHOLE input, and sorts
// N / 2^k claims per round instead of all N users: the per-rank exchange cost no longer grows with the number of GPUs.
//
// Collectives go through a small vtable (dapol_comm_ops): the built-in backend is NCCL over NVLink, resolved with dlopen at
// run time (the process may already hold a libnccl, e.g. torch's); a host can also plug its own transport.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: every NCCL symbol is resolved with dlsym (no link-time dependency)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

#include "dapol_internal.h"

// ------------------------------------------------------------------------------------------------ communicator
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;  // the copy the process already uses (torch's)
    if (!h) for (const char *n : names) if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { dapol_cuda_err() = std::string("libnccl.so.2 not found: ") + dlerror(); return nullptr; }
#define LOAD(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); if (!api.name) { dapol_cuda_err() = "nccl" #name " missing"; return nullptr; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllGather) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
    api.lib = h;
    return &api;
}
struct dapol_comm {
    int rank = 0, world = 1;
    dapol_comm_ops ops = {};
    ncclComm_t nccl = nullptr;  // built-in backend
    int device = -1;
};
#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r_ = (expr);                                                                        \
        if (r_ != ncclSuccess) { dapol_cuda_err() = std::string(#expr) + ": " + nccl_api()->GetErrorString(r_); return DAPOL_ERR_CUDA; } \
    } while (0)
static int nccl_all_gather(void *user, const void *d_send, void *d_recv, uint64_t bytes, void *stream) {
    dapol_comm *c = static_cast<dapol_comm *>(user);
    NCCL_TRY(nccl_api()->AllGather(d_send, d_recv, bytes, ncclUint8, c->nccl, static_cast<cudaStream_t>(stream)));
    return DAPOL_OK;
}
static int nccl_all_to_all(void *user, const void *d_send, const uint64_t *send_off, const uint64_t *send_bytes, void *d_recv,
                           const uint64_t *recv_off, const uint64_t *recv_bytes, void *stream) {
    dapol_comm *c = static_cast<dapol_comm *>(user);
    NcclApi *a = nccl_api();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NCCL_TRY(a->GroupStart());
    for (int r = 0; r < c->world; r++) {
        if (send_bytes[r]) NCCL_TRY(a->Send(static_cast<const uint8_t *>(d_send) + send_off[r], send_bytes[r], ncclUint8, r, c->nccl, st));
        if (recv_bytes[r]) NCCL_TRY(a->Recv(static_cast<uint8_t *>(d_recv) + recv_off[r], recv_bytes[r], ncclUint8, r, c->nccl, st));
    }
    NCCL_TRY(a->GroupEnd());
    return DAPOL_OK;
}
extern "C" int dapol_comm_create(int rank, int world, const dapol_comm_ops *ops, dapol_comm **out) {
    if (!out || !ops || !ops->all_gather || !ops->all_to_all || world < 1 || rank < 0 || rank >= world) return DAPOL_ERR_BAD_ARG;
    dapol_comm *c = new dapol_comm();
    c->rank = rank; c->world = world; c->ops = *ops;
    *out = c;
    return DAPOL_OK;
}
extern "C" int dapol_comm_nccl_unique_id(uint8_t id[DAPOL_NCCL_ID_BYTES]) {
    static_assert(DAPOL_NCCL_ID_BYTES == sizeof(ncclUniqueId), "ncclUniqueId size");
    NcclApi *a = nccl_api();
    if (!a || !id) return a ? DAPOL_ERR_BAD_ARG : DAPOL_ERR_CUDA;
    ncclUniqueId u;
    NCCL_TRY(a->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return DAPOL_OK;
}
extern "C" int dapol_comm_nccl_create(dapol_ctx *ctx, const uint8_t id[DAPOL_NCCL_ID_BYTES], int rank, int world, dapol_comm **out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return DAPOL_ERR_BAD_ARG;
    NcclApi *a = nccl_api();
    if (!a) return DAPOL_ERR_CUDA;
    CUDA_TRY(cudaSetDevice(ctx->device));
    dapol_comm *c = new dapol_comm();
    c->rank = rank; c->world = world; c->device = ctx->device;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = a->CommInitRank(&c->nccl, world, u, rank);
    if (r != ncclSuccess) { dapol_cuda_err() = std::string("ncclCommInitRank: ") + a->GetErrorString(r); delete c; return DAPOL_ERR_CUDA; }
    c->ops.user = c; c->ops.all_gather = nccl_all_gather; c->ops.all_to_all = nccl_all_to_all;
    *out = c;
    return DAPOL_OK;
}
extern "C" void dapol_comm_destroy(dapol_comm *c) {
    if (!c) return;
    if (c->nccl) { if (c->device >= 0) cudaSetDevice(c->device); nccl_api()->CommDestroy(c->nccl); }
    delete c;
}
extern "C" int dapol_comm_rank(const dapol_comm *c) { return c ? c->rank : -1; }
extern "C" int dapol_comm_world(const dapol_comm *c) { return c ? c->world : 0; }

// ------------------------------------------------------------------------------------------------ routing kernels
#define MAX_WORLD 256
// histogram of destinations (block-aggregated)
__global__ void k_route_hist(uint64_t n, const uint32_t *dest, int world, unsigned long long *counts) {
    __shared__ unsigned int sh[MAX_WORLD];
    for (int i = threadIdx.x; i < world; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&sh[dest[i]], 1u);
    __syncthreads();
    for (int d = threadIdx.x; d < world; d += blockDim.x) if (sh[d]) atomicAdd(&counts[d], (unsigned long long)sh[d]);
}
// slot of every item inside the send buffer: cursor[d] starts at the first slot of destination d
__global__ void k_route_slots(uint64_t n, const uint32_t *dest, int world, unsigned long long *cursor, uint64_t *slot) {
    __shared__ unsigned int cnt[MAX_WORLD];
    __shared__ unsigned long long base[MAX_WORLD];
    for (int i = threadIdx.x; i < world; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int mine = 0, d = 0;
    if (i < n) { d = dest[i]; mine = atomicAdd(&cnt[d], 1u); }
    __syncthreads();
    for (int t = threadIdx.x; t < world; t += blockDim.x) if (cnt[t]) base[t] = atomicAdd(&cursor[t], (unsigned long long)cnt[t]);
    __syncthreads();
    if (i < n) slot[i] = base[d] + mine;
}
// destination of the audit-id message of local user i: top k bits of the 64-bit audit-id prefix
__global__ void k_dest_audit(uint64_t n, const uint32_t *audit, int k, uint32_t *dest) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = (uint64_t)audit[8 * i] | ((uint64_t)audit[8 * i + 1] << 32);
    dest[i] = k ? (uint32_t)(key >> (64 - k)) : 0u;
}
// audit message = 10 words: audit id (8) | input position (2)
__global__ void k_pack_audit(uint64_t n, const uint32_t *audit, uint64_t first_pos, const uint64_t *slot, uint32_t *send) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t *m = send + 10 * slot[i];
    uint64_t pos = first_pos + i;
    for (int w = 0; w < 8; w++) m[w] = audit[8 * i + w];
    m[8] = (uint32_t)pos; m[9] = (uint32_t)(pos >> 32);
}
// duplicate-id screen: insert the 64-bit audit-id prefix of every message into an open-addressing table (all ones = empty);
// meeting an equal prefix (or the empty pattern itself) counts a suspect -- the exact pass then decides
__global__ void k_dup_probe(uint64_t n, const uint32_t *msgs, unsigned long long *table, uint64_t mask, unsigned long long *suspects) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long key = (unsigned long long)msgs[10 * j] | ((unsigned long long)msgs[10 * j + 1] << 32);
    if (key == ~0ull) { atomicAdd(suspects, 1ull); return; }
    for (uint64_t slot = (key >> 20) & mask;; slot = (slot + 1) & mask) {
        unsigned long long prev = atomicCAS(&table[slot], ~0ull, key);
        if (prev == ~0ull) return;
        if (prev == key) { atomicAdd(suspects, 1ull); return; }
    }
}
__global__ void k_audit_keys(uint64_t n, const uint32_t *msgs, uint64_t *keys, uint32_t *iota) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    keys[j] = (uint64_t)msgs[10 * j] | ((uint64_t)msgs[10 * j + 1] << 32);
    iota[j] = (uint32_t)j;
}
// sorted by 64-bit audit-id prefix (messages arrive in no particular order): every pair of identical ids reports the LATER
// of the two input positions; the reference fails at the first liability whose id was seen before = the minimum of those
__global__ void k_find_dups_msgs(uint64_t n, const uint64_t *keys, const uint32_t *who, const uint32_t *msgs, unsigned long long *first_dup) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n || keys[j] != keys[j - 1]) return;
    const uint32_t *a = msgs + 10 * (uint64_t)who[j];
    uint64_t pa = (uint64_t)a[8] | ((uint64_t)a[9] << 32);
    for (uint64_t t = j; t-- > 0 && keys[t] == keys[j];) {
        const uint32_t *b = msgs + 10 * (uint64_t)who[t];
        uint32_t d = 0;
        for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
        if (d == 0) {
            uint64_t pb = (uint64_t)b[8] | ((uint64_t)b[9] << 32);
            atomicMin(first_dup, (unsigned long long)(pa > pb ? pa : pb));
        }
    }
}
// claims of the local users `who[0..n)` (nullptr: all n local users): destination = k-bit prefix of the candidate index
__global__ void k_dest_claim(uint64_t n, const uint32_t *who, const uint64_t *cand, int shift, uint32_t *dest) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t i = who ? who[j] : j;
    dest[j] = shift >= 64 ? 0u : (uint32_t)(cand[i] >> shift);
}
// claim message = 14 words: candidate index (2) | input position (2) | value (2) | blinding (8)
#define CLAIM_WORDS 14
__global__ void k_pack_claim(uint64_t n, const uint32_t *who, const uint64_t *cand, const uint64_t *values, const uint32_t *blind,
                             uint64_t first_pos, const uint64_t *slot, uint32_t *send) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t i = who ? who[j] : j;
    uint32_t *m = send + CLAIM_WORDS * slot[j];
    uint64_t c = cand[i], p = first_pos + i, v = values[i];
    m[0] = (uint32_t)c; m[1] = (uint32_t)(c >> 32); m[2] = (uint32_t)p; m[3] = (uint32_t)(p >> 32); m[4] = (uint32_t)v; m[5] = (uint32_t)(v >> 32);
    for (int w = 0; w < 8; w++) m[6 + w] = blind[8 * i + w];
}
// owner side: received claims appended to the claim store (struct of arrays)
struct ClaimStore {
    uint64_t *cand = nullptr;  // index inside the shard (prefix stripped)
    uint64_t *pos = nullptr, *val = nullptr;
    uint32_t *blind = nullptr;
    uint8_t *dead = nullptr;
    uint64_t n = 0, cap = 0;
};
__global__ void k_unpack_claims(uint64_t n, const uint32_t *msgs, uint64_t mask, ClaimStore cs, uint64_t at) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t *m = msgs + CLAIM_WORDS * j;
    uint64_t o = at + j;
    cs.cand[o] = ((uint64_t)m[0] | ((uint64_t)m[1] << 32)) & mask;
    cs.pos[o] = (uint64_t)m[2] | ((uint64_t)m[3] << 32);
    cs.val[o] = (uint64_t)m[4] | ((uint64_t)m[5] << 32);
    for (int w = 0; w < 8; w++) cs.blind[8 * o + w] = m[6 + w];
    cs.dead[o] = 0;
}
__global__ void k_claim_pos_keys(uint64_t n, const uint64_t *pos, uint32_t *keys, uint32_t *iota) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    keys[j] = (uint32_t)pos[j];  // n_total < 2^31
    iota[j] = (uint32_t)j;
}
// second sort key: candidate index, dead claims behind every live one (dead bit above the shard's index bits)
__global__ void k_claim_cand_keys(uint64_t n, const uint32_t *by_pos, ClaimStore cs, int dead_bit, uint64_t *keys) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t o = by_pos[j];
    keys[j] = cs.cand[o] | ((uint64_t)cs.dead[o] << dead_bit);
}
// live claims sorted by (candidate, position): every claim but the first of a group lost its slot -- its position goes
// back to the user's home rank, the claim dies
__global__ void k_claim_composite_keys(uint64_t n, ClaimStore cs, int cand_bits, int pos_bits, uint64_t *keys, uint32_t *iota) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    keys[j] = ((((uint64_t)cs.dead[j] << cand_bits) | cs.cand[j]) << pos_bits) | cs.pos[j];
    iota[j] = (uint32_t)j;
}
// key_shift: low bits of the sort key that hold the position (composite keys), 0 for plain candidate keys
__global__ void k_claim_mark_losers(uint64_t n_live, const uint64_t *sorted_keys, int key_shift, const uint32_t *who, ClaimStore cs, unsigned long long *n_losers,
                                    uint64_t *loser_pos) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n_live || (sorted_keys[j] >> key_shift) != (sorted_keys[j - 1] >> key_shift)) return;
    uint32_t o = who[j];
    cs.dead[o] = 1;
    loser_pos[atomicAdd(n_losers, 1ull)] = cs.pos[o];
}
// home rank of an input position: last r with first_pos[r] <= pos
__global__ void k_dest_home(uint64_t n, const uint64_t *pos, const uint64_t *first_pos, int world, uint32_t *dest) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int lo = 0, hi = world - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (first_pos[mid] <= pos[j]) lo = mid; else hi = mid - 1;
    }
    dest[j] = (uint32_t)lo;
}
__global__ void k_pack_u64(uint64_t n, const uint64_t *x, const uint64_t *slot, uint64_t *send) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) send[slot[j]] = x[j];
}
// home side: a user that lost its slot re-hashes (mod.rs:416-437); `again` lists the local users with a new claim
__global__ void k_rehash_losers(uint64_t n, const uint64_t *loser_pos, uint64_t first_pos, int hash_id, int height, uint32_t *cur_seed, uint64_t *cand,
                                uint32_t *tries, unsigned long long *n_again, uint32_t *again, unsigned long long *min_failed) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t i = loser_pos[j] - first_pos;
    if (rehash_body(i, hash_id, height, cur_seed, cand, tries)) again[atomicAdd(n_again, 1ull)] = (uint32_t)i;
    else { tries[i] = 129; atomicMin(min_failed, (unsigned long long)loser_pos[j]); }
}
__global__ void k_fill_u32(uint64_t n, uint32_t *x, uint32_t v) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) x[j] = v;
}
// the shard's leaves: live claims in sorted order
__global__ void k_claims_to_leaves(uint64_t n_live, const uint32_t *who, ClaimStore cs, uint64_t *o_idx, uint64_t *o_val, uint32_t *o_blind) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_live) return;
    uint32_t o = who[j];
    o_idx[j] = cs.cand[o];
    o_val[j] = cs.val[o];
    uint32_t w[8];
    load8(w, cs.blind + 8 * (uint64_t)o);
    store8(o_blind + 8 * j, w);
}

// ------------------------------------------------------------------------------------------------ host orchestration
namespace {
// DAPOL_SHARD_TRACE=1: host time (us since the call began) at every point where the host waited for the stream, on rank 0
struct Trace {
    bool on = false;
    std::chrono::steady_clock::time_point t0;
    std::string log;
    void start(int rank) { const char *e = getenv("DAPOL_SHARD_TRACE"); on = e && *e == '1' && rank == 0; t0 = std::chrono::steady_clock::now(); }
    void mark(const char *what, uint64_t x = 0) {
        if (!on) return;
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        char b[96];
        snprintf(b, sizeof b, "%9.1f us  %s %llu\n", us, what, (unsigned long long)x);
        log += b;
    }
    ~Trace() { if (on) fprintf(stderr, "[dapol_sharded_build trace]\n%s", log.c_str()); }
};
struct DevBuf {  // stream-ordered allocation that frees itself
    void *p = nullptr;
    cudaStream_t st = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { dfree(p, st); }
    cudaError_t alloc(size_t n, cudaStream_t s) { dfree(p, st); p = nullptr; st = s; bytes = n; return dmalloc(&p, n, s); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct Shard {
    dapol_ctx *ctx;
    dapol_comm *comm;
    cudaStream_t st;
    int world, rank, k;
    DevBuf cnt_dev, mat_dev;  // [world] u64 send counts; [world][world] gathered
    std::vector<uint64_t> mat, send_cnt, send_off, recv_cnt, recv_off;
    uint64_t send_total = 0, recv_total = 0;
    int init() {
        CUDA_TRY(cnt_dev.alloc(8 * (size_t)world, st));
        CUDA_TRY(mat_dev.alloc(8 * (size_t)world * world, st));
        mat.resize((size_t)world * world); send_cnt.resize(world); send_off.resize(world); recv_cnt.resize(world); recv_off.resize(world);
        return DAPOL_OK;
    }
    // destinations -> per-destination counts, exchanged: afterwards send_* / recv_* hold ITEM counts and offsets
    int plan(uint64_t n, const uint32_t *dest) {
        CUDA_TRY(cudaMemsetAsync(cnt_dev.p, 0, 8 * (size_t)world, st));
        if (n) { k_route_hist<<<grid_for(n, 256), 256, 0, st>>>(n, dest, world, cnt_dev.as<unsigned long long>()); ctx->launches++; }
        int rc = comm->ops.all_gather(comm->ops.user, cnt_dev.p, mat_dev.p, 8 * (uint64_t)world, st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(mat.data(), mat_dev.p, 8 * (size_t)world * world, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        send_total = recv_total = 0;
        for (int r = 0; r < world; r++) {
            send_cnt[r] = mat[(size_t)rank * world + r]; send_off[r] = send_total; send_total += send_cnt[r];
            recv_cnt[r] = mat[(size_t)r * world + rank]; recv_off[r] = recv_total; recv_total += recv_cnt[r];
        }
        return DAPOL_OK;
    }
    uint64_t grand_total() const { uint64_t s = 0; for (uint64_t x : mat) s += x; return s; }
    // slot of every item in the send buffer (after plan)
    int slots(uint64_t n, const uint32_t *dest, uint64_t *slot) {
        if (!n) return DAPOL_OK;
        CUDA_TRY(cudaMemcpyAsync(cnt_dev.p, send_off.data(), 8 * (size_t)world, cudaMemcpyHostToDevice, st));
        k_route_slots<<<grid_for(n, 256), 256, 0, st>>>(n, dest, world, cnt_dev.as<unsigned long long>(), slot);
        ctx->launches++;
        return DAPOL_OK;
    }
    int all_to_all(const void *send, void *recv, uint64_t item_bytes) {
        std::vector<uint64_t> so(world), sb(world), ro(world), rb(world);
        for (int r = 0; r < world; r++) { so[r] = send_off[r] * item_bytes; sb[r] = send_cnt[r] * item_bytes; ro[r] = recv_off[r] * item_bytes; rb[r] = recv_cnt[r] * item_bytes; }
        return comm->ops.all_to_all(comm->ops.user, send, so.data(), sb.data(), recv, ro.data(), rb.data(), st);
    }
};
// grow the claim store to hold `need` claims (contents kept)
int claims_reserve(ClaimStore &cs, uint64_t need, cudaStream_t st) {
    if (need <= cs.cap) return DAPOL_OK;
    uint64_t cap = std::max<uint64_t>(need + need / 8 + 1024, 2 * cs.cap);
    ClaimStore n = cs;
    n.cap = cap;
    n.cand = n.pos = n.val = nullptr; n.blind = nullptr; n.dead = nullptr;
    CUDA_TRY(dmalloc(&n.cand, cap * 8, st)); CUDA_TRY(dmalloc(&n.pos, cap * 8, st)); CUDA_TRY(dmalloc(&n.val, cap * 8, st));
    CUDA_TRY(dmalloc(&n.blind, cap * 32, st)); CUDA_TRY(dmalloc(&n.dead, cap, st));
    if (cs.n) {
        CUDA_TRY(cudaMemcpyAsync(n.cand, cs.cand, cs.n * 8, cudaMemcpyDeviceToDevice, st)); CUDA_TRY(cudaMemcpyAsync(n.pos, cs.pos, cs.n * 8, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(n.val, cs.val, cs.n * 8, cudaMemcpyDeviceToDevice, st)); CUDA_TRY(cudaMemcpyAsync(n.blind, cs.blind, cs.n * 32, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(n.dead, cs.dead, cs.n, cudaMemcpyDeviceToDevice, st));
    }
    dfree(cs.cand, st); dfree(cs.pos, st); dfree(cs.val, st); dfree(cs.blind, st); dfree(cs.dead, st);
    cs = n;
    return DAPOL_OK;
}
void claims_free(ClaimStore &cs, cudaStream_t st) {
    dfree(cs.cand, st); dfree(cs.pos, st); dfree(cs.val, st); dfree(cs.blind, st); dfree(cs.dead, st);
    cs = ClaimStore();
}
// small fixed-size all-gather of host words through a device staging buffer
int gather_words(Shard &S, const uint64_t *mine, int words, std::vector<uint64_t> &all) {
    DevBuf s, r;
    CUDA_TRY(s.alloc(8 * (size_t)words, S.st)); CUDA_TRY(r.alloc(8 * (size_t)words * S.world, S.st));
    CUDA_TRY(cudaMemcpyAsync(s.p, mine, 8 * (size_t)words, cudaMemcpyHostToDevice, S.st));
    int rc = S.comm->ops.all_gather(S.comm->ops.user, s.p, r.p, 8 * (uint64_t)words, S.st);
    if (rc) return rc;
    all.resize((size_t)words * S.world);
    CUDA_TRY(cudaMemcpyAsync(all.data(), r.p, 8 * (size_t)words * S.world, cudaMemcpyDeviceToHost, S.st));
    CUDA_TRY(cudaStreamSynchronize(S.st));
    return DAPOL_OK;
}
}  // namespace

extern "C" int dapol_sharded_build(dapol_ctx *ctx, dapol_comm *comm, int hash_id, int height, uint64_t n_local, const uint8_t *d_iid_blob,
                                   const uint64_t *d_iid_off, const uint8_t *d_eid_blob, const uint64_t *d_eid_off, const uint64_t *d_values,
                                   const uint8_t *audit_seed, uint64_t audit_seed_len, const uint8_t pad_seed[32], uint64_t pad_base,
                                   dapol_tree **subtree, dapol_tree **top, uint64_t *n_total_out, uint64_t *first_pos_out, uint64_t *err_pos,
                                   float phase_ms[4]) {
    if (!ctx || !comm || !subtree || !top || !pad_seed) return DAPOL_ERR_BAD_ARG;
    *subtree = *top = nullptr;
    if (comm->world > 1 && ctx->leaf_hash_mode != DAPOL_LEAF_HASH_COMMITMENT) return DAPOL_ERR_BAD_ARG;  // claims do not carry id / salt hashes
    const int world = comm->world, rank = comm->rank;
    int k = 0;
    while ((1 << k) < world) k++;
    if ((1 << k) != world || world > MAX_WORLD) return DAPOL_ERR_BAD_ARG;
    if (hash_id != DAPOL_HASH_BLAKE3 && hash_id != DAPOL_HASH_BLAKE2S) return DAPOL_ERR_INVALID_DIGEST_SIZE;
    if (height > DAPOL_MAX_TREE_HEIGHT) return DAPOL_ERR_TREE_HEIGHT_TOO_BIG;
    const int Hs = height - k;
    if ((world > 1 && Hs < 1) || audit_seed_len > 512 || n_local >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (world == 1) {  // nothing to exchange: the single-GPU build; the whole tree comes back as `subtree`, there is no top tree
        if (n_total_out) *n_total_out = n_local;
        if (first_pos_out) *first_pos_out = 0;
        if (phase_ms) phase_ms[0] = phase_ms[1] = phase_ms[2] = phase_ms[3] = 0;
        return dapol_tree_build_from_liabilities_dev(ctx, hash_id, height, n_local, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, d_values, audit_seed,
                                                     audit_seed_len, pad_seed, pad_base, subtree, err_pos);
    }
    cudaEvent_t ev[5];
    for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 5; i++) cudaEventDestroy(e[i]); } } ev_guard{ev};
    CUDA_TRY(cudaEventRecord(ev[0], st));
    Shard S{ctx, comm, st, world, rank, k};
    Trace tr;
    tr.start(rank);
    int rc = S.init();
    if (rc) return rc;

    // ---- 1. per-user hashing of the local slice
    const uint64_t n = n_local, na = n ? n : 1;
    DevBuf b_audit, b_seed, b_cand, b_blind, b_tries, b_dest, b_slot, b_ctr;
    CUDA_TRY(b_audit.alloc(na * 32, st)); CUDA_TRY(b_seed.alloc(na * 32, st)); CUDA_TRY(b_cand.alloc(na * 8, st)); CUDA_TRY(b_blind.alloc(na * 32, st));
    CUDA_TRY(b_tries.alloc(na * 4, st)); CUDA_TRY(b_ctr.alloc(64, st));
    uint32_t *audit = b_audit.as<uint32_t>(), *cur_seed = b_seed.as<uint32_t>(), *blind = b_blind.as<uint32_t>(), *tries = b_tries.as<uint32_t>();
    uint64_t *cand = b_cand.as<uint64_t>();
    // counters: [0] losers of the round, [1] new claims of the round, [2] min failed position, [3] first duplicate position
    unsigned long long *ctr = b_ctr.as<unsigned long long>();
    int derive_rc = DAPOL_OK;
    if (n) {
        derive_rc = dapol_leaves_derive_dev(ctx, hash_id, height, n, d_iid_blob, d_iid_off, d_eid_blob, d_eid_off, audit_seed, audit_seed_len,
                                            b_audit.as<uint8_t>(), b_seed.as<uint8_t>(), cand, b_blind.as<uint8_t>());
        k_fill_u32<<<grid_for(n, 256), 256, 0, st>>>(n, tries, 1u);
        ctx->launches++;
    }
    std::vector<uint64_t> all;
    {
        uint64_t mine[2] = {n, (uint64_t)derive_rc};
        rc = gather_words(S, mine, 2, all);
        if (rc) return rc;
    }
    std::vector<uint64_t> first_pos(world + 1, 0);
    for (int r = 0; r < world; r++) {
        if (all[2 * r + 1]) return (int)all[2 * r + 1];  // a rank could not hash its slice (id too long, CUDA error): same verdict everywhere
        first_pos[r + 1] = first_pos[r] + all[2 * r];
    }
    tr.mark("hashed + sizes gathered", n);
    const uint64_t n_total = first_pos[world], my_first = first_pos[rank];
    if (n_total_out) *n_total_out = n_total;
    if (first_pos_out) *first_pos_out = my_first;
    if (n_total == 0 || n_total >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
    if (height < 64 && (1ull << height) < 2 * n_total) return DAPOL_ERR_SPARSITY_TOO_SMALL;  // MIN_SPARSITY = 2 (mod.rs:27,110)
    DevBuf b_first;
    CUDA_TRY(b_first.alloc(8 * (size_t)(world + 1), st));
    CUDA_TRY(cudaMemcpyAsync(b_first.p, first_pos.data(), 8 * (size_t)(world + 1), cudaMemcpyHostToDevice, st));
    {
        unsigned long long h[4] = {0, 0, ~0ull, ~0ull};
        CUDA_TRY(cudaMemcpyAsync(ctr, h, 32, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(b_dest.alloc(na * 4, st)); CUDA_TRY(b_slot.alloc(na * 8, st));
    uint32_t *dest = b_dest.as<uint32_t>();
    uint64_t *slot = b_slot.as<uint64_t>();
    CUDA_TRY(cudaEventRecord(ev[1], st));

    // ---- 2. duplicate internal ids: audit ids routed by their own top bits, exact comparison on the receiver
    {
        if (n) { k_dest_audit<<<grid_for(n, 256), 256, 0, st>>>(n, audit, k, dest); ctx->launches++; }
        if ((rc = S.plan(n, dest))) return rc;
        tr.mark("dup: planned", S.recv_total);
        if ((rc = S.slots(n, dest, slot))) return rc;
        DevBuf snd, rcv, keys, keys2, who, who2, tmp;
        CUDA_TRY(snd.alloc(40 * (size_t)(S.send_total + 1), st)); CUDA_TRY(rcv.alloc(40 * (size_t)(S.recv_total + 1), st));
        if (n) { k_pack_audit<<<grid_for(n, 256), 256, 0, st>>>(n, audit, my_first, slot, snd.as<uint32_t>()); ctx->launches++; }
        if ((rc = S.all_to_all(snd.p, rcv.p, 40))) return rc;
        const uint64_t nr = S.recv_total;
        // fast path: the 64-bit prefixes go into an open-addressing table; two equal prefixes (practically: a real duplicate) make
        // the exact sort-and-compare pass below run, otherwise the ids are all different and nothing more is needed
        unsigned long long suspects = 0;
        if (nr > 1) {
            uint64_t cap = 1024;
            while (cap < 2 * nr) cap <<= 1;
            DevBuf table;
            CUDA_TRY(table.alloc(8 * cap, st));
            CUDA_TRY(cudaMemsetAsync(table.p, 0xff, 8 * cap, st));
            CUDA_TRY(cudaMemsetAsync(ctr, 0, 8, st));
            k_dup_probe<<<grid_for(nr, 256), 256, 0, st>>>(nr, rcv.as<uint32_t>(), table.as<unsigned long long>(), cap - 1, ctr);
            ctx->launches++;
            CUDA_TRY(cudaMemcpyAsync(&suspects, ctr, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (suspects) {
            CUDA_TRY(keys.alloc(8 * nr, st)); CUDA_TRY(keys2.alloc(8 * nr, st)); CUDA_TRY(who.alloc(4 * nr, st)); CUDA_TRY(who2.alloc(4 * nr, st));
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint64_t>(), keys2.as<uint64_t>(), who.as<uint32_t>(), who2.as<uint32_t>(), (int)nr, 0, 64, st);
            CUDA_TRY(tmp.alloc(tb, st));
            k_audit_keys<<<grid_for(nr, 256), 256, 0, st>>>(nr, rcv.as<uint32_t>(), keys.as<uint64_t>(), who.as<uint32_t>());
            cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<uint64_t>(), keys2.as<uint64_t>(), who.as<uint32_t>(), who2.as<uint32_t>(), (int)nr, 0, 64, st);
            k_find_dups_msgs<<<grid_for(nr, 256), 256, 0, st>>>(nr, keys2.as<uint64_t>(), who2.as<uint32_t>(), rcv.as<uint32_t>(), ctr + 3);
            ctx->launches += 3;
        }
        CUDA_TRY(cudaStreamSynchronize(st));  // the buffers of this block are released in stream order, the host vectors are reused
        tr.mark("dup: exchanged + sorted + checked");
    }

    // ---- 3. claims: distributed fix-point of the first-come-first-served rule
    ClaimStore cs;
    struct CsGuard { ClaimStore &c; cudaStream_t s; ~CsGuard() { claims_free(c, s); } } cs_guard{cs, st};
    const uint64_t mask = Hs >= 64 ? ~0ull : (1ull << Hs) - 1;
    int pos_bits = 1;  // bits of an input position
    while ((1ull << pos_bits) < n_total) pos_bits++;
    const int shift = Hs;  // owner of a candidate = its top k bits
    DevBuf b_again, b_pkeys, b_pkeys2, b_who, b_who2, b_ckeys, b_ckeys2, b_who3, b_sorttmp, b_losers, b_ldest, b_lslot;
    CUDA_TRY(b_again.alloc(na * 4, st));
    uint32_t *again = b_again.as<uint32_t>();
    uint64_t n_claiming = n;       // local users sending a claim this round (round 0: everybody)
    const uint32_t *claim_list = nullptr;
    uint64_t sort_cap = 0, n_dead = 0;
    size_t sort_tmp_bytes = 0;
    uint64_t n_live = 0;
    for (int round = 0;; round++) {
        if (round > 4096) return DAPOL_ERR_BAD_ARG;
        // 3a. route the claims of this round to the owners of their candidates
        if (n_claiming) { k_dest_claim<<<grid_for(n_claiming, 256), 256, 0, st>>>(n_claiming, claim_list, cand, shift, dest); ctx->launches++; }
        if ((rc = S.plan(n_claiming, dest))) return rc;
        tr.mark("claims: planned, round", (uint64_t)round);
        if (round > 0 && S.grand_total() == 0) break;  // nobody re-hashed anywhere: fix-point
        if ((rc = S.slots(n_claiming, dest, slot))) return rc;
        {
            DevBuf snd, rcv;
            CUDA_TRY(snd.alloc(4 * CLAIM_WORDS * (size_t)(S.send_total + 1), st)); CUDA_TRY(rcv.alloc(4 * CLAIM_WORDS * (size_t)(S.recv_total + 1), st));
            if (n_claiming) {
                k_pack_claim<<<grid_for(n_claiming, 256), 256, 0, st>>>(n_claiming, claim_list, cand, d_values, blind, my_first, slot, snd.as<uint32_t>());
                ctx->launches++;
            }
            if ((rc = S.all_to_all(snd.p, rcv.p, 4 * CLAIM_WORDS))) return rc;
            if ((rc = claims_reserve(cs, cs.n + S.recv_total, st))) return rc;
            if (S.recv_total) {
                k_unpack_claims<<<grid_for(S.recv_total, 256), 256, 0, st>>>(S.recv_total, rcv.as<uint32_t>(), mask, cs, cs.n);
                ctx->launches++;
                cs.n += S.recv_total;
            }
        }
        // 3b. owner: sort the claims by (candidate, position); all but the first of a group lose
        const uint64_t nc = cs.n;
        if (nc >= (1ull << 31)) return DAPOL_ERR_BAD_ARG;
        uint64_t n_losers = 0;
        if (nc) {
            if (nc > sort_cap) {
                sort_cap = nc + nc / 8 + 1024;
                CUDA_TRY(b_pkeys.alloc(4 * sort_cap, st)); CUDA_TRY(b_pkeys2.alloc(4 * sort_cap, st)); CUDA_TRY(b_who.alloc(4 * sort_cap, st));
                CUDA_TRY(b_who2.alloc(4 * sort_cap, st)); CUDA_TRY(b_ckeys.alloc(8 * sort_cap, st)); CUDA_TRY(b_ckeys2.alloc(8 * sort_cap, st));
                CUDA_TRY(b_who3.alloc(4 * sort_cap, st)); CUDA_TRY(b_losers.alloc(8 * sort_cap, st));
                CUDA_TRY(b_ldest.alloc(4 * sort_cap, st)); CUDA_TRY(b_lslot.alloc(8 * sort_cap, st));
                size_t t1 = 0, t2 = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, t1, b_pkeys.as<uint32_t>(), b_pkeys2.as<uint32_t>(), b_who.as<uint32_t>(), b_who2.as<uint32_t>(), (int)sort_cap, 0, 31, st);
                cub::DeviceRadixSort::SortPairs(nullptr, t2, b_ckeys.as<uint64_t>(), b_ckeys2.as<uint64_t>(), b_who2.as<uint32_t>(), b_who3.as<uint32_t>(), (int)sort_cap, 0, 64, st);
                sort_tmp_bytes = std::max(t1, t2);
                CUDA_TRY(b_sorttmp.alloc(sort_tmp_bytes, st));
            }
            size_t tb = sort_tmp_bytes;
            int key_shift = 0;
            if (Hs + 1 + pos_bits <= 64) {
                // (dead, candidate, position) fits one 64-bit key: ONE radix sort orders the claims by candidate and, inside a group, by position
                key_shift = pos_bits;
                k_claim_composite_keys<<<grid_for(nc, 256), 256, 0, st>>>(nc, cs, Hs, pos_bits, b_ckeys.as<uint64_t>(), b_who2.as<uint32_t>());
                cub::DeviceRadixSort::SortPairs(b_sorttmp.p, tb, b_ckeys.as<uint64_t>(), b_ckeys2.as<uint64_t>(), b_who2.as<uint32_t>(), b_who3.as<uint32_t>(), (int)nc, 0,
                                                Hs + 1 + pos_bits, st);
                ctx->launches += 2;
            } else {  // tall shards: stable sort by candidate of the claims sorted by position
                k_claim_pos_keys<<<grid_for(nc, 256), 256, 0, st>>>(nc, cs.pos, b_pkeys.as<uint32_t>(), b_who.as<uint32_t>());
                cub::DeviceRadixSort::SortPairs(b_sorttmp.p, tb, b_pkeys.as<uint32_t>(), b_pkeys2.as<uint32_t>(), b_who.as<uint32_t>(), b_who2.as<uint32_t>(), (int)nc, 0, 31, st);
                k_claim_cand_keys<<<grid_for(nc, 256), 256, 0, st>>>(nc, b_who2.as<uint32_t>(), cs, Hs, b_ckeys.as<uint64_t>());
                tb = sort_tmp_bytes;
                cub::DeviceRadixSort::SortPairs(b_sorttmp.p, tb, b_ckeys.as<uint64_t>(), b_ckeys2.as<uint64_t>(), b_who2.as<uint32_t>(), b_who3.as<uint32_t>(), (int)nc, 0,
                                                std::min(Hs + 1, 64), st);
                ctx->launches += 4;
            }
            n_live = nc - n_dead;
            CUDA_TRY(cudaMemsetAsync(ctr, 0, 16, st));
            if (n_live) k_claim_mark_losers<<<grid_for(n_live, 256), 256, 0, st>>>(n_live, b_ckeys2.as<uint64_t>(), key_shift, b_who3.as<uint32_t>(), cs, ctr, b_losers.as<uint64_t>());
            ctx->launches += 1;
            unsigned long long h = 0;
            CUDA_TRY(cudaMemcpyAsync(&h, ctr, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            n_losers = h;
            n_dead += n_losers;
            tr.mark("claims: exchanged + sorted + losers marked", n_losers);
        } else {
            n_live = 0;
            CUDA_TRY(cudaMemsetAsync(ctr, 0, 16, st));
        }
        // 3c. the losers' positions go home; the home rank re-hashes and claims again next round
        uint32_t *ldest = b_ldest.as<uint32_t>();
        uint64_t *lslot = b_lslot.as<uint64_t>();
        if (n_losers) {
            k_dest_home<<<grid_for(n_losers, 256), 256, 0, st>>>(n_losers, b_losers.as<uint64_t>(), b_first.as<uint64_t>(), world, ldest);
            ctx->launches++;
        }
        if ((rc = S.plan(n_losers, ldest))) return rc;
        tr.mark("losers: planned", S.grand_total());
        if (S.grand_total() == 0) { n_live = cs.n - n_dead; break; }  // no collision anywhere: fix-point (the last sort holds the leaves)
        if ((rc = S.slots(n_losers, ldest, lslot))) return rc;
        {
            DevBuf snd, rcv;
            CUDA_TRY(snd.alloc(8 * (size_t)(S.send_total + 1), st)); CUDA_TRY(rcv.alloc(8 * (size_t)(S.recv_total + 1), st));
            if (n_losers) { k_pack_u64<<<grid_for(n_losers, 256), 256, 0, st>>>(n_losers, b_losers.as<uint64_t>(), lslot, snd.as<uint64_t>()); ctx->launches++; }
            if ((rc = S.all_to_all(snd.p, rcv.p, 8))) return rc;
            n_claiming = 0;
            if (S.recv_total) {
                if (S.recv_total > n) return DAPOL_ERR_BAD_ARG;  // a user holds one live claim at a time
                k_rehash_losers<<<grid_for(S.recv_total, 256), 256, 0, st>>>(S.recv_total, rcv.as<uint64_t>(), my_first, hash_id, height, cur_seed, cand, tries,
                                                                              ctr + 1, again, ctr + 2);
                ctx->launches++;
                unsigned long long h = 0;
                CUDA_TRY(cudaMemcpyAsync(&h, ctr + 1, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                n_claiming = h;
                tr.mark("losers: exchanged + re-hashed", n_claiming);
            }
            claim_list = again;
        }
    }
    // ---- global verdict of the leaf stage: DuplicatedInternalId / FailedToMapIndex at the earliest input position
    {
        unsigned long long h[4];
        CUDA_TRY(cudaMemcpyAsync(h, ctr, 32, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        uint64_t mine[2] = {h[2], h[3]};
        if ((rc = gather_words(S, mine, 2, all))) return rc;
        uint64_t f = ~0ull, d = ~0ull;
        for (int r = 0; r < world; r++) { f = std::min<uint64_t>(f, all[2 * r]); d = std::min<uint64_t>(d, all[2 * r + 1]); }
        if (d != ~0ull && d <= f) { if (err_pos) *err_pos = d; return DAPOL_ERR_DUPLICATED_INTERNAL_ID; }
        if (f != ~0ull) { if (err_pos) *err_pos = f; return DAPOL_ERR_FAILED_TO_MAP_INDEX; }
    }
    tr.mark("verdict gathered");
    CUDA_TRY(cudaEventRecord(ev[2], st));

    // ---- 4. the surviving claims, sorted by candidate, are this shard's leaves
    const uint64_t n_mine = n_live;
    DevBuf l_idx, l_val, l_blind;
    CUDA_TRY(l_idx.alloc(8 * (n_mine + 1), st)); CUDA_TRY(l_val.alloc(8 * (n_mine + 1), st)); CUDA_TRY(l_blind.alloc(32 * (n_mine + 1), st));
    if (n_mine) {
        k_claims_to_leaves<<<grid_for(n_mine, 256), 256, 0, st>>>(n_mine, b_who3.as<uint32_t>(), cs, l_idx.as<uint64_t>(), l_val.as<uint64_t>(), l_blind.as<uint32_t>());
        ctx->launches++;
    }
    // ---- 5. padding-stream bases (creation order of the single tree: level H..1, shard by shard inside a level)
    std::vector<uint64_t> level_base(Hs + 1, 0);
    uint64_t top_base = pad_base;
    if (ctx->pad_mode == DAPOL_PADDING_POSITIONAL) {
        level_base[0] = (uint64_t)k;
        for (int h = 1; h <= Hs; h++) level_base[h] = h >= 64 ? 0 : (uint64_t)rank << h;
        top_base = 0;
    } else {
        std::vector<uint64_t> counts(Hs + 1, 0);
        if (n_mine && (rc = dapol_tree_level_pad_counts_dev(ctx, Hs, n_mine, l_idx.as<uint64_t>(), counts.data()))) return rc;
        if ((rc = gather_words(S, counts.data(), Hs + 1, all))) return rc;
        uint64_t below = 0;
        for (int h = Hs; h >= 1; h--) {
            uint64_t before = 0, tot = 0;
            for (int r = 0; r < world; r++) { uint64_t c = all[(size_t)r * (Hs + 1) + h]; tot += c; if (r < rank) before += c; }
            level_base[h] = pad_base + below + before;
            below += tot;
        }
        top_base = pad_base + below;
    }
    // ---- 6. subtree (no data-path collective), root records, top tree
    dapol_tree *sub = nullptr;
    uint64_t rec[(DAPOL_RECORD_BYTES + 8) / 8] = {0};
    if (n_mine) {
        rc = dapol_tree_build_shard_dev(ctx, hash_id, Hs, n_mine, l_idx.as<uint64_t>(), l_val.as<uint64_t>(), l_blind.as<uint8_t>(), pad_seed, level_base.data(), &sub);
        if (rc == DAPOL_OK) rc = dapol_tree_root_record(sub, reinterpret_cast<uint8_t *>(rec));
        rec[DAPOL_RECORD_BYTES / 8] = rc == DAPOL_OK ? 1 : 2 + (uint64_t)rc;
    }
    tr.mark("subtree built", n_mine);
    float build_ms[5] = {0, 0, 0, 0, 0};
    memcpy(build_ms, ctx->last_ms, sizeof build_ms);
    CUDA_TRY(cudaEventRecord(ev[3], st));
    int grc = gather_words(S, rec, (DAPOL_RECORD_BYTES + 8) / 8, all);
    if (grc) { dapol_tree_destroy(sub); return grc; }
    std::vector<uint64_t> present;
    std::vector<uint8_t> recs;
    const size_t rw = (DAPOL_RECORD_BYTES + 8) / 8;
    for (int r = 0; r < world; r++) {
        uint64_t flag = all[r * rw + DAPOL_RECORD_BYTES / 8];
        if (flag >= 2) { dapol_tree_destroy(sub); return (int)(flag - 2); }  // a shard failed to build: same verdict everywhere
        if (flag == 1) {
            present.push_back((uint64_t)r);
            const uint8_t *p = reinterpret_cast<const uint8_t *>(&all[r * rw]);
            recs.insert(recs.end(), p, p + DAPOL_RECORD_BYTES);
        }
    }
    if (present.empty()) { dapol_tree_destroy(sub); return DAPOL_ERR_BAD_ARG; }
    dapol_tree *tp = nullptr;
    rc = dapol_tree_build_from_records(ctx, hash_id, k, present.size(), present.data(), recs.data(), pad_seed, top_base, &tp);
    if (rc) { dapol_tree_destroy(sub); return rc; }
    if (sub && (rc = dapol_tree_attach_top(sub, tp, (uint64_t)rank))) { dapol_tree_destroy(sub); dapol_tree_destroy(tp); return rc; }
    // id_to_idx_map (mod.rs:80,389) of the LOCAL slice lives on the top-tree handle: positions [first_pos, first_pos + n_local)
    if (n) {
        tp->leaf_index_of = cand;
        b_cand.p = nullptr;  // ownership moved to the tree
        tp->index_map_first = my_first;
        tp->index_map_n = n;
        tp->audit_ids = audit;  // dapol_tree_index_of on the top-tree handle: the ids of the local slice
        b_audit.p = nullptr;
        if (audit_seed_len) tp->audit_seed.assign(audit_seed, audit_seed + audit_seed_len);
    }
    memcpy(ctx->last_ms, build_ms, sizeof build_ms);  // dapol_last_build_times: the subtree build, not the tiny top tree
    CUDA_TRY(cudaEventRecord(ev[4], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    tr.mark("top tree built");
    if (phase_ms) for (int i = 0; i < 4; i++) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
    *subtree = sub;
    *top = tp;
    return DAPOL_OK;
}
