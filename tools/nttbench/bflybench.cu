// Standalone butterfly-throughput experiments for the Fr NTT (not part of the library): how close to the multiplier pipe can a
// register-resident radix-4 round get, without any memory traffic, at a given number of resident warps?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../zcash-gpu-thesis_b200/csrc -o bflybench bflybench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace b200zk;

// ---- variant 0: the library's butterfly (canonical values)
struct V0 {
    static __device__ __forceinline__ void bfly(fr_t &lo, fr_t &hi, const fr_t &w) {
        const fr_t t = fr_t::mul_inline(hi, w);
        hi = lo - t;
        lo = lo + t;
    }
};
// ---- variant 1: products only (no add / sub): the multiplier-bound part alone
struct V1 {
    static __device__ __forceinline__ void bfly(fr_t &lo, fr_t &hi, const fr_t &w) {
        const fr_t t = fr_t::mul_inline(hi, w);
        hi = lo;
        lo = t;
    }
};
#define HAVE_LAZY 1
struct V2 {
    static __device__ __forceinline__ void bfly(fr_t &lo, fr_t &hi, const fr_t &w) {
        const fr_t t = fr_t::mul_inline_t<false>(w, hi);
        hi = fr_t::sub_2p(lo, t);
        lo = fr_t::add_2p(lo, t);
    }
};

template <class V, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_round(int iters, uint32_t seed, uint32_t *out) {
    fr_t x[4], w[3];
#pragma unroll
    for (int k = 0; k < 4; k++) { x[k] = fr_t::one(); x[k].v[0] ^= threadIdx.x + k; x[k].v[3] ^= seed; }
#pragma unroll
    for (int k = 0; k < 3; k++) { w[k] = fr_t::r2(); w[k].v[1] ^= seed + k + (threadIdx.x >> 3); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        V::bfly(x[0], x[1], w[0]);
        V::bfly(x[2], x[3], w[0]);
        V::bfly(x[0], x[2], w[1]);
        V::bfly(x[1], x[3], w[2]);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= x[k].v[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
static void run(const char *name, K kernel, int threads, int blocks_per_sm, int iters) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm;
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
    double bf = 4.0 * iters * (double)blocks * threads;
    double rate = bf / best / 1e6;  // G butterflies / s
    printf("%-34s warps/SM=%2d  %.3f ms  %.2f Gbfly/s  -> 2^24 NTT (2.013e8 bfly) %.3f ms  (%s)\n", name, blocks_per_sm * threads / 32, best, rate,
           2.013e8 / rate / 1e6, cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    const int iters = 2000;
    run("canonical bfly, 16 warps", k_round<V0, 256, 2>, 256, 2, iters);
    run("canonical bfly, 24 warps", k_round<V0, 256, 3>, 256, 3, iters);
    run("canonical bfly, 32 warps", k_round<V0, 256, 4>, 256, 4, iters);
    run("product only, 16 warps", k_round<V1, 256, 2>, 256, 2, iters);
    run("product only, 24 warps", k_round<V1, 256, 3>, 256, 3, iters);
    run("product only, 32 warps", k_round<V1, 256, 4>, 256, 4, iters);
#ifdef HAVE_LAZY
    run("lazy bfly, 16 warps", k_round<V2, 256, 2>, 256, 2, iters);
    run("lazy bfly, 24 warps", k_round<V2, 256, 3>, 256, 3, iters);
    run("lazy bfly, 32 warps", k_round<V2, 256, 4>, 256, 4, iters);
#endif
    return 0;
}
