#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "/root/repo/dapol_b200/csrc/tree_kernels.cuh"
__device__ const uint8_t c_tag_s[9] = {'s', 'a', 'l', 't', '_', 's', 'e', 'e', 'd'};
__device__ const uint8_t c_tag_l[4] = {'l', 'e', 'a', 'f'};
DAPOL_HD_INLINE int both_stages(int hash_id, const uint32_t a[8], const uint8_t *eid, uint32_t elen, uint32_t h[8]) {
    const uint8_t tag_s[9] = {'s', 'a', 'l', 't', '_', 's', 'e', 'e', 'd'}, tag_l[4] = {'l', 'e', 'a', 'f'};
    dapol_hasher hs; uint32_t salt[8];
    hasher_init(hs, hash_id);
    hasher_update_words(hs, a, 8); hasher_update(hs, tag_s, 9); hasher_update(hs, eid, elen);
    int rc = hasher_final(hs, salt);
    hasher_init(hs, hash_id);
    hasher_update(hs, tag_l, 4); hasher_update(hs, eid, elen); hasher_update_words(hs, salt, 8);
    rc |= hasher_final(hs, h);
    return rc;
}
// W2: each stage its own non-inlined function
static __device__ __noinline__ int stage_salt(int hash_id, const uint32_t *a, const uint8_t *eid, uint32_t elen, uint32_t *salt) {
    const uint8_t tag_s[9] = {'s', 'a', 'l', 't', '_', 's', 'e', 'e', 'd'};
    dapol_hasher hs;
    hasher_init(hs, hash_id);
    hasher_update_words(hs, a, 8); hasher_update(hs, tag_s, 9); hasher_update(hs, eid, elen);
    return hasher_final(hs, salt);
}
static __device__ __noinline__ int stage_leaf(int hash_id, const uint32_t *salt, const uint8_t *eid, uint32_t elen, uint32_t *h) {
    const uint8_t tag_l[4] = {'l', 'e', 'a', 'f'};
    dapol_hasher hs;
    hasher_init(hs, hash_id);
    hasher_update(hs, tag_l, 4); hasher_update(hs, eid, elen); hasher_update_words(hs, salt, 8);
    return hasher_final(hs, h);
}
template <int V>
__global__ void k_var(uint64_t n, int hash_id, const uint32_t *audit, const uint8_t *eid_blob, const uint64_t *eid_off, uint32_t *out, uint32_t *tmp) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *eid = eid_blob + eid_off[i];
    const uint32_t elen = (uint32_t)(eid_off[i + 1] - eid_off[i]);
    uint32_t a[8], h[8], salt[8];
    load8(a, audit + 8 * i);
    if (V == 0) both_stages(hash_id, a, eid, elen, h);
    if (V == 2) { stage_salt(hash_id, a, eid, elen, salt); stage_leaf(hash_id, salt, eid, elen, h); }
    if (V == 3) {
        dapol_hasher hs;
        hasher_init(hs, hash_id);
        hasher_update_words(hs, a, 8); hasher_update(hs, c_tag_s, 9); hasher_update(hs, eid, elen);
        hasher_final(hs, salt);
        hasher_init(hs, hash_id);
        hasher_update(hs, c_tag_l, 4); hasher_update(hs, eid, elen); hasher_update_words(hs, salt, 8);
        hasher_final(hs, h);
    }
    if (V == 4) {
        const uint8_t tag_s[9] = {'s', 'a', 'l', 't', '_', 's', 'e', 'e', 'd'};
        const uint8_t tag_l[4] = {'l', 'e', 'a', 'f'};
        dapol_hasher hs;
        hasher_init(hs, hash_id);
        hasher_update_words(hs, a, 8); hasher_update(hs, tag_s, 9); hasher_update(hs, eid, elen);
        hasher_final(hs, salt);
        hasher_init(hs, hash_id);
        hasher_update(hs, tag_l, 4); hasher_update(hs, eid, elen); hasher_update_words(hs, salt, 8);
        hasher_final(hs, h);
    }
    if (V == 10) { stage_salt(hash_id, a, eid, elen, salt); store8(tmp + 8 * i, salt); return; }   // W1: two kernels
    if (V == 11) { load8(salt, tmp + 8 * i); stage_leaf(hash_id, salt, eid, elen, h); }
    store8(out + 8 * i, h);
}
int main() {
    const int lens[6] = {0, 4, 33, 1200, 64, 1024};
    std::vector<uint8_t> blob; std::vector<uint64_t> off{0}; std::vector<uint32_t> audit;
    for (int c = 0; c < 24; c++) { int L = lens[c % 6]; for (int j = 0; j < L; j++) blob.push_back((uint8_t)(c * 7 + j * 13)); off.push_back(blob.size()); for (int k = 0; k < 8; k++) audit.push_back(0x9e3779b9u * (c * 8 + k + 1)); }
    { const char *hex = "87f05110d6f85eaa048ff82160b73bffc96dfd7dedfd50c52fe1c64af0fba25d"; uint8_t b[32];
      for (int i = 0; i < 32; i++) { unsigned x; sscanf(hex + 2 * i, "%2x", &x); b[i] = (uint8_t)x; }
      for (int k = 0; k < 8; k++) audit.push_back((uint32_t)b[4*k] | ((uint32_t)b[4*k+1] << 8) | ((uint32_t)b[4*k+2] << 16) | ((uint32_t)b[4*k+3] << 24));
      for (int j = 1; j <= 4; j++) blob.push_back((uint8_t)j); off.push_back(blob.size()); }
    const uint64_t n = off.size() - 1;
    uint8_t *d_blob; uint64_t *d_off; uint32_t *d_audit, *d_out, *d_tmp;
    cudaMalloc(&d_blob, blob.size() + 1); cudaMalloc(&d_off, off.size() * 8); cudaMalloc(&d_audit, audit.size() * 4); cudaMalloc(&d_out, n * 32); cudaMalloc(&d_tmp, n * 32);
    cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice); cudaMemcpy(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_audit, audit.data(), audit.size() * 4, cudaMemcpyHostToDevice);
    for (int hid = 0; hid < 2; hid++) {
        std::vector<uint32_t> want(n * 8), got(n * 8);
        for (uint64_t i = 0; i < n; i++) both_stages(hid, &audit[8 * i], blob.data() + off[i], (uint32_t)(off[i + 1] - off[i]), &want[8 * i]);
        auto check = [&](const char *name) {
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), d_out, n * 32, cudaMemcpyDeviceToHost);
            int bad = 0; for (uint64_t i = 0; i < n; i++) bad += memcmp(&got[8 * i], &want[8 * i], 32) != 0;
            printf("hash_id %d %-28s mismatches %d of %d (%s)\n", hid, name, bad, (int)n, cudaGetErrorString(e));
        };
        k_var<0><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); check("W0 current shape");
        k_var<4><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); check("W4 inline in the kernel");
        { auto bad_rows = [&]() { for (uint64_t i = 0; i < n; i++) if (memcmp(&got[8 * i], &want[8 * i], 32)) printf(" %d(len %d)", (int)i, (int)(off[i+1]-off[i])); printf("\n"); }; bad_rows(); }
        k_var<2><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); check("W2 noinline stages");
        k_var<3><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); check("W3 __device__ const tags");
        k_var<10><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); k_var<11><<<1, 128>>>(n, hid, d_audit, d_blob, d_off, d_out, d_tmp); check("W1 two kernels");
    }
}
