// Standalone Fq-multiply throughput experiments (not part of the library): which instruction mix does the
// B200 integer pipe like?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mulbench mulbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp_v0.cuh"
#include "fp_v1.cuh"

template <class F, int CHAINS>
__global__ void __launch_bounds__(256) k_mul(int iters, uint32_t seed, uint32_t *out) {
    F a[CHAINS], c[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) { a[k] = F::one(); a[k].v[0] ^= threadIdx.x + k; c[k] = F::r2(); c[k].v[1] ^= seed + k; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) a[k] = a[k] * c[k];
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++)
#pragma unroll
        for (int i = 0; i < F::N; i++) acc ^= a[k].v[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// raw products only: 12 x 12 mul.wide + xor, no carries (upper bound for the FMA pipe in this register context)
template <int DUMMY>
__global__ void __launch_bounds__(256) k_wide_only(int iters, uint32_t seed, uint32_t *out) {
    uint32_t a[12], b[12];
    unsigned long long acc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = seed + threadIdx.x * 3 + i; b[i] = seed * 7 + i; acc[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 12; i++)
#pragma unroll
            for (int j = 0; j < 12; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[i]), "r"(b[j]));
#pragma unroll
        for (int i = 0; i < 12; i++) a[i] ^= (uint32_t)acc[i];
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
static void run(const char *name, K kernel, int iters, double ops_per_thread_iter, int blocks_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm, threads = 256;
    uint32_t *out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(iters, 1234u, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(iters, 1234u + r, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double ops = ops_per_thread_iter * iters * (double)blocks * threads;
    printf("%-28s blocks/SM=%d  %.3f ms  %.2f Gops/s  (%s)\n", name, blocks_per_sm, best, ops / best / 1e6, cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    const int iters = 1000;
    for (int bps : {1, 2, 4}) {
        run("fq mul v0 (mad.cc) x2", k_mul<v0::fq_t, 2>, iters, 2, bps);
        run("fq mul v1 (wide+add) x2", k_mul<v1::fq_t, 2>, iters, 2, bps);
        run("fq mul v1 (wide+add) x1", k_mul<v1::fq_t, 1>, iters, 1, bps);
        run("fr mul v0 x2", k_mul<v0::fr_t, 2>, iters, 2, bps);
        run("fr mul v1 x2", k_mul<v1::fr_t, 2>, iters, 2, bps);
        run("144 mad.wide only", k_wide_only<0>, iters, 144, bps);
    }
    return 0;
}
