// Radix-2 NTT over BLS12-381 Fr on sm_100a: the device implementation of bellman's EvaluationDomain
// transforms (bellman/src/domain.rs:83-132 fft/ifft/coset_fft/icoset_fft, :261-374 best_fft/serial_fft/
// parallel_fft) and of the H-polynomial block of the prover (groth16/prover.rs:256-287).
//
// The reference does a bit-reversal permutation followed by log n in-place DIT stages (domain.rs:286-314).
// Here the same butterflies are grouped into ceil(log n / 8) passes over HBM; one pass stages a tile of
// 2^B butterfly-coupled elements x 2^q adjacent columns in shared memory (limb-major, conflict-free),
// runs B stages there and writes the tile back.  The bit reversal is folded into the first pass's loads,
// the coset / inverse scalings (distribute_powers, m^-1; domain.rs:88-118) into the first load / last store.
// Twiddles omega^k are precomputed once per domain size and cached in the context.  Every butterfly is the
// reference's (domain.rs:300-308): t = a[hi]*w; a[hi] = a[lo]-t; a[lo] += t, on canonical values, so the
// output is bit-identical to serial_fft / parallel_fft.
#include "fp.cuh"
#include "internal.h"

namespace b200zk {

static constexpr uint32_t LO_BITS = 12;  // two-level g^i table: g^i = lo[i & 4095] * hi[i >> 12]
static constexpr int NTT_THREADS = 128;
static constexpr uint32_t MAX_TILE_LOG = 10;

enum { C_OMEGA = 0, C_OMEGA_INV = 1, C_N_INV = 2, C_G = 3, C_G_INV = 4, C_Z_INV = 5, C_G_STEP = 6, C_GI_STEP = 7, C_COUNT = 8 };

// fr.rs:50-55 ROOT_OF_UNITY and fr.rs:38-44 GENERATOR (=7), Montgomery limbs
__device__ __constant__ uint32_t FR_ROOT_OF_UNITY[8] = {0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u, 0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau};
__device__ __constant__ uint32_t FR_GENERATOR[8] = {0xfffffff1u, 0x0000000eu, 0x00189c0fu, 0x17e363d3u, 0x6f8457b0u, 0xff9c5787u, 0x8fc5a8c4u, 0x35133220u};

// EvaluationDomain::from_coeffs constants (domain.rs:64-79) + z(g)^-1 (domain.rs:136-148), one thread
__global__ void k_ntt_setup(fr_t *consts, uint32_t log_n) {
    fr_t omega, g;
#pragma unroll
    for (int i = 0; i < 8; i++) { omega.v[i] = FR_ROOT_OF_UNITY[i]; g.v[i] = FR_GENERATOR[i]; }
    for (uint32_t i = log_n; i < 32; i++) omega = omega.sqr();
    fr_t n = fr_t::zero();
    uint64_t nn = 1ull << log_n;
    n.v[0] = (uint32_t)nn;
    n.v[1] = (uint32_t)(nn >> 32);
    n = n.to_mont();
    fr_t g_inv = g.inverse();
    consts[C_OMEGA] = omega;
    consts[C_OMEGA_INV] = omega.inverse();
    consts[C_N_INV] = n.inverse();
    consts[C_G] = g;
    consts[C_G_INV] = g_inv;
    consts[C_Z_INV] = (g.pow(nn) - fr_t::one()).inverse();
    consts[C_G_STEP] = g.pow(1ull << LO_BITS);
    consts[C_GI_STEP] = g_inv.pow(1ull << LO_BITS);
}

// out[k] = base^k * (scale ? *scale : 1), k < count; 16 consecutive powers per thread
__global__ void k_pow_table(fr_t *__restrict__ out, const fr_t *__restrict__ base, const fr_t *__restrict__ scale, size_t count) {
    size_t k0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (k0 >= count) return;
    fr_t b = base[0];
    fr_t cur = b.pow(k0);
    if (scale) cur = cur * scale[0];
    for (int i = 0; i < 16 && k0 + i < count; i++) {
        out[k0 + i] = cur;
        cur = cur * b;
    }
}

struct NttPass {
    uint32_t log_n, s0, B, q;
    int bitrev_load;  // gather in[bitrev(g)] (first pass)
    int pre_scale;    // multiply loaded element i by lo[i & mask] * hi[i >> LO_BITS]   (coset_fft: g^i)
    int post_scale;   // 1: multiply stored element by consts[C_N_INV] (ifft); 2: by lo*hi tables (icoset_fft: g^-i / n)
};

// Tile layout in shared memory: two planes of 16-byte halves (limbs 0-3 of every element, then limbs 4-7).  Consecutive
// threads touch consecutive 16-byte words of a plane, so an element moves with 2 conflict-free LDS.128 / STS.128 -- the
// kernel is bound by instruction issue as much as by the multiplier (one butterfly = 114 wide multiplies x 4 issue
// cycles against ~450 instructions), and the limb-major layout used before cost 32 LDS/STS.32 per butterfly instead of 8.
#ifdef B200ZK_NTT_LIMB_MAJOR
__device__ __forceinline__ fr_t sm_load(const uint32_t *sm, uint32_t tile, uint32_t e) {
    fr_t x;
#pragma unroll
    for (int l = 0; l < 8; l++) x.v[l] = sm[l * tile + e];
    return x;
}
__device__ __forceinline__ void sm_store(uint32_t *sm, uint32_t tile, uint32_t e, const fr_t &x) {
#pragma unroll
    for (int l = 0; l < 8; l++) sm[l * tile + e] = x.v[l];
}
#else
__device__ __forceinline__ fr_t sm_load(const uint32_t *sm, uint32_t tile, uint32_t e) {
    const uint4 *p = reinterpret_cast<const uint4 *>(sm);
    const uint4 lo = p[e], hi = p[tile + e];
    fr_t x;
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
    x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    return x;
}
__device__ __forceinline__ void sm_store(uint32_t *sm, uint32_t tile, uint32_t e, const fr_t &x) {
    uint4 *p = reinterpret_cast<uint4 *>(sm);
    p[e] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    p[tile + e] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
#endif

#ifndef B200ZK_NTT_UNROLL
#define B200ZK_NTT_UNROLL 1  // 2 and 4 measured no faster (3.86 / 3.88 / 3.88 ms at 2^24)
#endif
// CB / CQ: the pass shape as compile-time constants (0 = take it from `p`): the 2^24 transform runs three (8, 2) passes,
// and with constant shifts the index arithmetic of a butterfly shrinks (the kernel is issue-bound, see sm_load).
template <int CB, int CQ>
__global__ void __launch_bounds__(NTT_THREADS, 7) k_ntt_pass(const fr_t *__restrict__ in, fr_t *__restrict__ out, const fr_t *__restrict__ tw,
                                                         const fr_t *__restrict__ sc_lo, const fr_t *__restrict__ sc_hi,
                                                         const fr_t *__restrict__ consts, NttPass p) {
    extern __shared__ __align__(16) uint32_t sm[];
    if (CB) { p.B = CB; p.q = CQ; }
    constexpr int UNROLL = CB ? B200ZK_NTT_UNROLL : 1;  // butterflies of one stage unrolled per thread (specialised shapes)
    const uint32_t T = p.B + p.q, TILE = 1u << T;
    const uint32_t tile = blockIdx.x;
    const uint32_t vshift = p.s0 == 0 ? 0 : p.q;
    const uint32_t midbits = p.s0 == 0 ? 0 : p.s0 - p.q;
    const uint32_t mid = tile & ((1u << midbits) - 1), high = tile >> midbits;
    auto gidx = [&](uint32_t e) -> uint32_t {
        if (p.s0 == 0) return (tile << T) | e;
        uint32_t u = e & ((1u << p.q) - 1), v = e >> p.q;
        return (high << (p.s0 + p.B)) | (v << p.s0) | (mid << p.q) | u;
    };
    for (uint32_t e = threadIdx.x; e < TILE; e += NTT_THREADS) {
        uint32_t g = gidx(e);
        uint32_t src = p.bitrev_load ? (__brev(g) >> (32 - p.log_n)) : g;
        fr_t x = in[src];
        if (p.pre_scale) x = (x * sc_lo[src & ((1u << LO_BITS) - 1)]) * sc_hi[src >> LO_BITS];
        sm_store(sm, TILE, e, x);
    }
    __syncthreads();
    for (uint32_t k = 0; k < p.B; k++) {
        const uint32_t s = p.s0 + k, bitpos = k + vshift;
#pragma unroll(UNROLL)
        for (uint32_t bf = threadIdx.x; bf < TILE / 2; bf += NTT_THREADS) {
            uint32_t lo = ((bf >> bitpos) << (bitpos + 1)) | (bf & ((1u << bitpos) - 1));
            uint32_t hi = lo | (1u << bitpos);
            uint32_t j = gidx(lo) & ((1u << s) - 1);
            fr_t w = tw[(size_t)j << (p.log_n - 1 - s)];
            fr_t a = sm_load(sm, TILE, lo);
            fr_t t = sm_load(sm, TILE, hi) * w;
            sm_store(sm, TILE, lo, a + t);
            sm_store(sm, TILE, hi, a - t);
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < TILE; e += NTT_THREADS) {
        uint32_t g = gidx(e);
        fr_t x = sm_load(sm, TILE, e);
        if (p.post_scale == 1) x = x * consts[C_N_INV];
        else if (p.post_scale == 2) x = (x * sc_lo[g & ((1u << LO_BITS) - 1)]) * sc_hi[g >> LO_BITS];
        out[g] = x;
    }
}

// distribute_powers with an arbitrary g (domain.rs:105-118): 16 consecutive elements per thread
__global__ void k_distribute_powers(fr_t *__restrict__ a, const fr_t *__restrict__ g, size_t n) {
    size_t k0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (k0 >= n) return;
    fr_t b = g[0];
    fr_t cur = b.pow(k0);
    for (int i = 0; i < 16 && k0 + i < n; i++) {
        a[k0 + i] = a[k0 + i] * cur;
        cur = cur * b;
    }
}

// prover.rs:267-271: a = (a*b - c) * z(g)^-1
__global__ void k_h_combine(fr_t *__restrict__ a, const fr_t *__restrict__ b, const fr_t *__restrict__ c, const fr_t *__restrict__ consts, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = (a[i] * b[i] - c[i]) * consts[C_Z_INV];
}
// prover.rs:287: Fr -> FrRepr for the first n elements
__global__ void k_into_repr(const fr_t *__restrict__ a, fr_t *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = a[i].from_mont();
}

// divide_by_z_on_coset (domain.rs:146-159): a[i] *= 1 / (g^m - 1), the constant cached with the domain tables
__global__ void k_scale_by_const(fr_t *__restrict__ a, const fr_t *__restrict__ consts, int which, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = a[i] * consts[which];
}
// z(tau) = tau^m - 1 (domain.rs:136-141), one thread
__global__ void k_domain_z(const fr_t *tau, uint32_t log_m, fr_t *out) {
    fr_t t = tau[0];
    for (uint32_t i = 0; i < log_m; i++) t = t.sqr();
    out[0] = t - fr_t::one();
}

static void free_tables(NttTables &t) {
    cudaFree(t.tw); cudaFree(t.tw_inv); cudaFree(t.g_lo); cudaFree(t.g_hi); cudaFree(t.gi_lo); cudaFree(t.gi_hi); cudaFree(t.consts);
    t = NttTables();
}

int ntt_get_tables(Ctx *ctx, uint32_t log_n, NttTables **out) {
    auto it = ctx->ntt_tables.find(log_n);
    if (it != ctx->ntt_tables.end()) { *out = &it->second; return B200ZK_OK; }
    NttTables t;
    t.log_n = log_n;
    const size_t n = (size_t)1 << log_n;
    const size_t half = n > 1 ? n / 2 : 1;
    const size_t lo_cnt = (size_t)1 << LO_BITS;
    const size_t hi_cnt = (n >> LO_BITS) ? (n >> LO_BITS) : 1;
    B200ZK_CUDA(ctx, cudaMalloc(&t.consts, C_COUNT * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.tw, half * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.tw_inv, half * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.g_lo, lo_cnt * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.gi_lo, lo_cnt * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.g_hi, hi_cnt * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.gi_hi, hi_cnt * sizeof(fr_t)));
    fr_t *c = (fr_t *)t.consts;
    k_ntt_setup<<<1, 1, 0, ctx->stream>>>(c, log_n);
    auto blocks = [](size_t cnt) { return (unsigned)((cnt + 16 * 128 - 1) / (16 * 128)); };
    k_pow_table<<<blocks(half), 128, 0, ctx->stream>>>((fr_t *)t.tw, c + C_OMEGA, nullptr, half);
    k_pow_table<<<blocks(half), 128, 0, ctx->stream>>>((fr_t *)t.tw_inv, c + C_OMEGA_INV, nullptr, half);
    k_pow_table<<<blocks(lo_cnt), 128, 0, ctx->stream>>>((fr_t *)t.g_lo, c + C_G, nullptr, lo_cnt);
    k_pow_table<<<blocks(lo_cnt), 128, 0, ctx->stream>>>((fr_t *)t.gi_lo, c + C_G_INV, nullptr, lo_cnt);
    k_pow_table<<<blocks(hi_cnt), 128, 0, ctx->stream>>>((fr_t *)t.g_hi, c + C_G_STEP, nullptr, hi_cnt);
    k_pow_table<<<blocks(hi_cnt), 128, 0, ctx->stream>>>((fr_t *)t.gi_hi, c + C_GI_STEP, c + C_N_INV, hi_cnt);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { free_tables(t); return set_error(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e)); }
    auto ins = ctx->ntt_tables.emplace(log_n, t);
    *out = &ins.first->second;
    return B200ZK_OK;
}

void ntt_free_all_tables(Ctx *ctx) {
    for (auto &kv : ctx->ntt_tables) free_tables(kv.second);
    ctx->ntt_tables.clear();
}

int ntt_run(Ctx *ctx, void *d_coeffs, uint32_t log_n, int kind) {
    if (log_n >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");  // domain.rs:59-61
    if (kind < B200ZK_FFT || kind > B200ZK_ICOSET_FFT) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad ntt kind");
    if (log_n == 0) return B200ZK_OK;  // m = 1: omega = 1, m^-1 = 1, g^0 = 1 -> every transform is the identity
    if (log_n > 30) return set_error(ctx, B200ZK_ERR_BAD_ARG, "log_m > 30 not supported on one GPU");
    NttTables *t;
    int st = ntt_get_tables(ctx, log_n, &t);
    if (st) return st;
    const size_t n = (size_t)1 << log_n;
    st = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, n * sizeof(fr_t));
    if (st) return st;
    fr_t *A = (fr_t *)d_coeffs, *S = (fr_t *)ctx->scratch;
    const bool inverse = kind == B200ZK_IFFT || kind == B200ZK_ICOSET_FFT;
    const fr_t *tw = (const fr_t *)(inverse ? t->tw_inv : t->tw);

    uint32_t npass = log_n <= MAX_TILE_LOG ? 1 : (log_n + 7) / 8;
    uint32_t base = log_n / npass, extra = log_n % npass;
    uint32_t s0 = 0;
    for (uint32_t ps = 0; ps < npass; ps++) {
        NttPass p;
        p.log_n = log_n;
        p.s0 = s0;
        p.B = base + (ps < extra ? 1 : 0);
        p.q = npass == 1 ? 0 : (MAX_TILE_LOG - p.B < 2 ? MAX_TILE_LOG - p.B : 2);
        while (p.q > 0 && (n >> (p.B + p.q)) < (size_t)2 * ctx->sm_count) p.q--;  // small transforms: more, narrower tiles to fill the SMs
        p.bitrev_load = ps == 0;
        p.pre_scale = (ps == 0 && kind == B200ZK_COSET_FFT) ? 1 : 0;
        p.post_scale = ps + 1 == npass ? (kind == B200ZK_IFFT ? 1 : kind == B200ZK_ICOSET_FFT ? 2 : 0) : 0;
        const fr_t *lo = (const fr_t *)(p.pre_scale ? t->g_lo : t->gi_lo), *hi = (const fr_t *)(p.pre_scale ? t->g_hi : t->gi_hi);
        const fr_t *src = ps == 0 ? A : S;
        fr_t *dst = (ps + 1 == npass && npass > 1) ? A : S;
        uint32_t T = p.B + p.q;
        size_t smem = (size_t)8 * sizeof(uint32_t) << T;
        unsigned tiles = (unsigned)(n >> T);
#define B200ZK_NTT_SHAPE(CB, CQ) \
    if (p.B == CB && p.q == CQ) k_ntt_pass<CB, CQ><<<tiles, NTT_THREADS, smem, ctx->stream>>>(src, dst, tw, lo, hi, (const fr_t *)t->consts, p); else
        B200ZK_NTT_SHAPE(8, 2) B200ZK_NTT_SHAPE(7, 2) B200ZK_NTT_SHAPE(6, 2) B200ZK_NTT_SHAPE(5, 2)
        k_ntt_pass<0, 0><<<tiles, NTT_THREADS, smem, ctx->stream>>>(src, dst, tw, lo, hi, (const fr_t *)t->consts, p);
#undef B200ZK_NTT_SHAPE
        ctx->launches++;
        s0 += p.B;
    }
    if (npass == 1) B200ZK_CUDA(ctx, cudaMemcpyAsync(A, S, n * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

int ntt_distribute_powers(Ctx *ctx, void *d_coeffs, size_t n, const void *d_g) {
    if (n == 0) return B200ZK_OK;
    k_distribute_powers<<<(unsigned)((n + 16 * 128 - 1) / (16 * 128)), 128, 0, ctx->stream>>>((fr_t *)d_coeffs, (const fr_t *)d_g, n);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

int ntt_divide_by_z_on_coset(Ctx *ctx, void *d_coeffs, uint32_t log_n) {
    if (log_n >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");
    NttTables *t;
    int st = ntt_get_tables(ctx, log_n, &t);
    if (st) return st;
    const size_t n = (size_t)1 << log_n;
    k_scale_by_const<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((fr_t *)d_coeffs, (const fr_t *)t->consts, C_Z_INV, n);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}
int ntt_domain_z(Ctx *ctx, const void *d_tau, uint32_t log_n, void *d_out) {
    k_domain_z<<<1, 1, 0, ctx->stream>>>((const fr_t *)d_tau, log_n, (fr_t *)d_out);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

int ntt_h_poly(Ctx *ctx, void *d_a, void *d_b, void *d_c, uint32_t log_n, void *d_out_repr) {
    int st;
    const size_t n = (size_t)1 << log_n;
    void *v[3] = {d_a, d_b, d_c};
    for (int i = 0; i < 3; i++) {
        if ((st = ntt_run(ctx, v[i], log_n, B200ZK_IFFT))) return st;
        if ((st = ntt_run(ctx, v[i], log_n, B200ZK_COSET_FFT))) return st;
    }
    NttTables *t;
    if ((st = ntt_get_tables(ctx, log_n, &t))) return st;
    ctx->launches += 2;
    k_h_combine<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((fr_t *)d_a, (const fr_t *)d_b, (const fr_t *)d_c, (const fr_t *)t->consts, n);
    if ((st = ntt_run(ctx, d_a, log_n, B200ZK_ICOSET_FFT))) return st;
    if (n > 1) k_into_repr<<<(unsigned)((n - 1 + 255) / 256), 256, 0, ctx->stream>>>((const fr_t *)d_a, (fr_t *)d_out_repr, n - 1);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
