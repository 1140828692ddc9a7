// Standalone field-op throughput microbenchmark (one generated mul/sqr variant per build).
#include <cstdio>
#include <cuda_runtime.h>
#include "../dapol_b200/csrc/fe25519.cuh"
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters) {
    fe a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; b.v[i] = a.v[i] ^ 0x9e3779b9u; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { fe_mul(a, a, b); fe_mul(b, b, a); }
        else if (OP == 1) { fe_sq(a, a); fe_sq(b, b); }
        else { fe t; fe_add(t, a, b); fe_sub(b, a, b); fe_mul(a, t, b); fe_sq(b, b); }  // add/sub/mul/sq mix
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a.v[i] ^ b.v[i];
    if (r == 0x12345u) out[0] = r;
}
template <int OP>
static double run(int sms) {
    uint32_t *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = sms * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    return (double)blocks * threads * iters * 2.0 / (best * 1e-3) / 1e9;
}
int main(int argc, char **argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\"variant\": \"%s\", \"fe_mul_Gop_s\": %.1f, \"fe_sq_Gop_s\": %.1f, \"mix_Gop_s\": %.1f}\n", argc > 1 ? argv[1] : "?", run<0>(p.multiProcessorCount),
           run<1>(p.multiProcessorCount), run<2>(p.multiProcessorCount));
    return 0;
}
