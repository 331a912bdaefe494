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
#include <algorithm>
#include <cstdlib>
#include "fp.cuh"
#include "internal.h"

namespace b200zk {

enum { C_OMEGA = 0, C_OMEGA_INV = 1, C_N_INV = 2, C_G = 3, C_G_INV = 4, C_Z_INV = 5, C_COUNT = 8 };

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

// ------------------------------------------------------------------------------------------------ the pass kernels
// One pass = B consecutive radix-2 DIT stages (bits [s0, s0 + B) of the element index) over tiles of 2^B coupled elements x 2^Q
// independent columns.  A thread owns 2^XB elements in registers and runs XB stages on them without touching shared memory;
// between such rounds the tile is regrouped (warp shuffles / shared memory).  The first round reads its elements straight from
// HBM, the last one writes them straight back.  Twiddles omega^k come from the per-domain table (precomputed once).
//   k_ntt_pass  (XB = 1): small, latency-bound transforms -- one butterfly per thread and stage, out-of-line butterfly body.
//   k_ntt_pass4 (XB = 2): large, work-bound transforms -- radix-4 rounds, everything inlined (further down).
//   FIRST pass: input gathered in bit-reversed order (domain.rs:286-295 folded into the loads: the columns are the TOP index
//   bits, so the four columns of a tile are four ADJACENT source elements = one 128-byte line) and the twiddles of its first
//   round are 1, omega^(n/4), ...: the products by 1 are skipped (x * 1 = x exactly, the output stays bit-identical).
//   coset_fft / icoset_fft / ifft scalings: ONE product per element at the first load (g^i table) / the last store
//   (g^-i / n table or the constant 1 / n).
struct NttPass {
    const fr_t *in;
    fr_t *out;
    const fr_t *tw;       // omega^k (or omega^-k), k < n / 2
    const fr_t *scale;    // pre_scale: g^i, i < n;  post_scale == 2: g^-i / n, i < n;  post_scale == 1: &n^-1
    uint32_t log_n, s0;
    int pre_scale, post_scale;
    int last;             // the last pass of the transform (k_ntt_pass4 delivers canonical values only there)
    size_t in_stride, out_stride;  // blockIdx.y = which vector of a batch of equal transforms: element offsets of its input / output
    uint32_t batch;
};

struct FrPair { fr_t lo, hi; };
// The product inside is the fused operand-scanning Montgomery product of fp.cuh.
static __device__ __noinline__ FrPair ntt_bfly_call(fr_t lo, fr_t hi, fr_t w) {
    const fr_t t = fr_t::mul_inline(hi, w);
    return {lo + t, lo - t};
}
static __device__ __noinline__ fr_t ntt_mul_call(fr_t a, fr_t b) { return fr_t::mul_inline(a, b); }

// shared-memory tile: two planes of 16-byte halves, XOR-swizzled so that the 8 lanes of a quarter warp always hit 8 different
// 16-byte bank groups whatever bit positions the round's register-resident index bits occupy
__device__ __forceinline__ uint32_t ntt_swz(uint32_t e) { return e ^ ((e >> 3) & 7u); }
template <int TILE>
__device__ __forceinline__ fr_t ntt_sm_load(const uint4 *sm, uint32_t e) {
    const uint32_t w = ntt_swz(e);
    const uint4 lo = sm[w], hi = sm[TILE + w];
    fr_t x;
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
    x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    return x;
}
template <int TILE>
__device__ __forceinline__ void ntt_sm_store(uint4 *sm, uint32_t e, const fr_t &x) {
    const uint32_t w = ntt_swz(e);
    sm[w] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    sm[TILE + w] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ void ntt_prefetch(const fr_t *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// K stages on the E = 2^XB register-resident elements x[0..E): x's index bit i is the i-th of the round's stage bits.
// idx0 = the table index of the first stage's twiddle; stage j, butterfly with low index bits xl = xi & (2^j - 1):
// twiddle index = (idx0 >> j) + (xl << (log_n - 1 - j)).  TRIVIAL: idx0 = 0 (first round of the first pass), so the butterflies
// with xl = 0 multiply by omega^0 = 1 and skip the product.
// the twiddle addresses of a round only depend on the thread: requested into L1 before the round's data is touched
template <int K, bool TRIVIAL>
__device__ __forceinline__ void ntt_round_prefetch(const fr_t *__restrict__ tw, uint32_t idx0, uint32_t log_n) {
#pragma unroll
    for (int j = 0; j < K; j++)
#pragma unroll
        for (int xl = TRIVIAL ? 1 : 0; xl < (1 << j); xl++) ntt_prefetch(tw + ((idx0 >> j) + ((uint32_t)xl << (log_n - 1 - j))));
}
template <int XB, int K, bool TRIVIAL>
__device__ __forceinline__ void ntt_round(fr_t (&x)[1 << XB], const fr_t *__restrict__ tw, uint32_t idx0, uint32_t log_n) {
    constexpr int E = 1 << XB;
    if (XB > 1) ntt_round_prefetch<K, TRIVIAL>(tw, idx0, log_n);  // measured: prefetching every round's twiddles at kernel start is slower
#pragma unroll
    for (int j = 0; j < K; j++) {
        fr_t w0;
        if (j == 0 && !TRIVIAL) w0 = tw[idx0];
#pragma unroll
        for (int xi = 0; xi < E; xi++) {
            if (xi & (1 << j)) continue;
            const int xl = xi & ((1 << j) - 1), hi = xi | (1 << j);
            if (TRIVIAL && xl == 0) {  // w = 1
                const fr_t a = x[xi], b = x[hi];
                x[xi] = a + b;
                x[hi] = a - b;
            } else {
                const fr_t w = j == 0 ? w0 : tw[(idx0 >> j) + ((uint32_t)xl << (log_n - 1 - j))];
                const FrPair r = ntt_bfly_call(x[xi], x[hi], w);
                x[xi] = r.lo;
                x[hi] = r.hi;
            }
        }
    }
}

// XB = 1: one butterfly per thread and stage; the elements change hands by warp shuffle while the partner is a lane of the same
// warp, through shared memory afterwards -- for small transforms, whose time is the serial chain of a thread, not the work
template <int B, int Q, int XB>
struct NttShape {
    static constexpr int T = B + Q, TILE = 1 << T, THREADS = 1 << (T - XB);
    static constexpr int K0 = (B - 1) % XB + 1, ROUNDS = (B + XB - 1) / XB;  // the short round comes first: rounds of K0, XB, XB ... stages
    static constexpr int WARPS = 32;  // resident warps per SM the register budget is set for
    static constexpr int MINBLOCKS = (WARPS * 32 / THREADS) < 1 ? 1 : (WARPS * 32 / THREADS);
    static constexpr size_t SMEM = ROUNDS > 1 ? (size_t)2 * TILE * sizeof(uint4) : 0;
};

template <int B, int Q, int XB, bool FIRST>
__global__ void __launch_bounds__(NttShape<B, Q, XB>::THREADS, NttShape<B, Q, XB>::MINBLOCKS) k_ntt_pass(NttPass p) {
    typedef NttShape<B, Q, XB> S;
    constexpr int E = 1 << XB;
    extern __shared__ __align__(16) uint4 ntt_sm[];
    const uint32_t t = threadIdx.x, tile = blockIdx.x;
    const uint32_t log_n = p.log_n, s0 = FIRST ? 0u : p.s0;
    // tile-local element e = (v << Q) | col  ->  global index
    //   FIRST:  g = col << (log_n - Q) | tile << B | v      (columns = top bits: adjacent after the bit reversal)
    //   else:   g = high << (s0 + B) | v << s0 | mid << Q | col,  tile = high << (s0 - Q) | mid
    const uint32_t mid = FIRST ? 0u : tile & ((1u << (s0 - Q)) - 1u), high = FIRST ? 0u : tile >> (s0 - Q);
    auto gidx = [&](uint32_t e) -> uint32_t {
        const uint32_t v = e >> Q, col = e & ((1u << Q) - 1u);
        if (FIRST) return (Q ? col << (log_n - Q) : 0u) | (tile << B) | v;
        return (high << (s0 + B)) | (v << s0) | (mid << Q) | col;
    };
    // twiddle base index of a round's first stage: (g mod 2^(s0 + b)) << (log_n - 1 - s0 - b), b = the round's first stage bit
    auto round_idx0 = [&](int P) -> uint32_t {
        const uint32_t e0 = ((t >> P) << (P + XB)) | (t & ((1u << P) - 1u));
        const uint32_t sb = s0 + (uint32_t)(P - Q);
        return sb == 0 ? 0u : (gidx(e0) & ((1u << sb) - 1u)) << (log_n - 1 - sb);
    };
    fr_t x[E];
#pragma unroll
    for (int r = 0; r < S::ROUNDS; r++) {
        const int P = (r == 0 ? 0 : S::K0 + XB * (r - 1)) + Q;  // position of the XB register-resident bits inside e
        const uint32_t e0 = ((t >> P) << (P + XB)) | (t & ((1u << P) - 1u));
        const uint32_t idx0 = round_idx0(P);
        if (r == 0) {
#pragma unroll
            for (int xi = 0; xi < E; xi++) {
                const uint32_t g = gidx(e0 | ((uint32_t)xi << P));
                const uint32_t src = FIRST ? __brev(g) >> (32 - log_n) : g;
                x[xi] = (p.in + (size_t)blockIdx.y * p.in_stride)[src];
                if (FIRST && p.pre_scale) x[xi] = ntt_mul_call(x[xi], p.scale[src]);
            }
        } else if (!(XB == 1 && P - 1 < 5)) {  // (otherwise the previous round handed its elements over by warp shuffle)
            __syncthreads();
#pragma unroll
            for (int xi = 0; xi < E; xi++) x[xi] = ntt_sm_load<S::TILE>(ntt_sm, e0 | ((uint32_t)xi << P));
        }
        if (r == 0) {
            if (S::K0 == 1) ntt_round<XB, 1, FIRST>(x, p.tw, idx0, log_n);
            else if (S::K0 == 2) ntt_round<XB, 2, FIRST>(x, p.tw, idx0, log_n);
            else ntt_round<XB, 3, FIRST>(x, p.tw, idx0, log_n);
        } else {
            ntt_round<XB, XB, false>(x, p.tw, idx0, log_n);
        }
        if (r == S::ROUNDS - 1) {
#pragma unroll
            for (int xi = 0; xi < E; xi++) {
                const uint32_t g = gidx(e0 | ((uint32_t)xi << P));
                if (p.post_scale == 1) x[xi] = ntt_mul_call(x[xi], p.scale[0]);
                else if (p.post_scale == 2) x[xi] = ntt_mul_call(x[xi], p.scale[g]);
                (p.out + (size_t)blockIdx.y * p.out_stride)[g] = x[xi];
            }
        } else if (XB == 1 && P < 5) {
            // Warp-shuffle butterfly exchange.  The next stage pairs elements whose index differs in the bit this thread's id holds
            // at position P, and its partner there is lane ^ (1 << P) of the same warp: each of the two keeps the element whose
            // stage bit equals its own id bit and swaps the other one -- no shared memory, no barrier.
            const bool up = (t >> P) & 1u;
            const fr_t send = up ? x[0] : x[E - 1];
            fr_t recv;
#pragma unroll
            for (int l = 0; l < 8; l++) recv.v[l] = __shfl_xor_sync(0xffffffffu, send.v[l], 1u << P);
            if (up) x[0] = recv; else x[E - 1] = recv;
        } else {
#pragma unroll
            for (int xi = 0; xi < E; xi++) ntt_sm_store<S::TILE>(ntt_sm, e0 | ((uint32_t)xi << P), x[xi]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ radix-4 rounds, inlined
// The large-transform pass: a thread owns FOUR elements and runs two stages on them (4 butterflies, 3 twiddles), fully inlined --
// no call, so none of the ~38 register moves per butterfly that marshal operands into and out of an out-of-line body (all of
// which ptxas splits between the ALU and the *multiplier* pipe as IMAD.MOV) -- and the rounds of a pass are a real loop, so the
// kernel stays ~20 KB and inside the instruction cache.  The three twiddles of a round only depend on the thread: they are
// requested right after the previous round's elements went to shared memory (their registers are free then) and arrive while the
// block waits at the barrier, so no butterfly waits on a dependent table load.  Index conventions as in k_ntt_pass.
template <int B, int Q>
struct NttShape4 {
    static constexpr int T = B + Q, TILE = 1 << T, THREADS = 1 << (T - 2);
    static constexpr int K0 = (B & 1) ? 1 : 2, ROUNDS = (B + 1) / 2;  // an odd pass starts with a one-stage round
    static constexpr int MINBLOCKS = (768 / THREADS) < 1 ? 1 : (768 / THREADS);  // 24 resident warps: 80 registers (2^24 fft: 20 warps / 94 registers 3.20 ms, 28 / 72 3.22, 24 / 80 3.13)
    static constexpr size_t SMEM = (size_t)2 * TILE * sizeof(uint4);
};
__device__ __forceinline__ uint32_t ntt_swz4(uint32_t e) { return e ^ ((e >> 2) & 6u); }  // bits 3, 4 -> bits 1, 2: conflict-free for every P >= 1
template <int TILE>
__device__ __forceinline__ fr_t ntt_sm_load4(const uint4 *sm, uint32_t e) {
    const uint32_t w = ntt_swz4(e);
    const uint4 lo = sm[w], hi = sm[TILE + w];
    fr_t x;
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
    x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    return x;
}
template <int TILE>
__device__ __forceinline__ void ntt_sm_store4(uint4 *sm, uint32_t e, const fr_t &x) {
    const uint32_t w = ntt_swz4(e);
    sm[w] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    sm[TILE + w] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
// Between the stages of a large transform the values live in [0, 2r) (2r < 2^256): the product of such a value with a canonical
// twiddle is < 2r without its final subtraction (fp.cuh mul_inline_t<false>), sums and differences are folded back below 2r, and
// the last round of the last pass subtracts r once more where needed.  Same residues, hence the same canonical outputs.
__device__ __forceinline__ void ntt_bfly_inline(fr_t &lo, fr_t &hi, const fr_t &w) {
    const fr_t t = fr_t::mul_inline_t<false>(w, hi);  // the canonical operand is the one multiplied as a whole (see fp.cuh)
    hi = fr_t::sub_2p(lo, t);
    lo = fr_t::add_2p(lo, t);
}
__device__ __forceinline__ void ntt_bfly_trivial(fr_t &lo, fr_t &hi) {
    const fr_t t = hi;
    hi = fr_t::sub_2p(lo, t);
    lo = fr_t::add_2p(lo, t);
}

template <int B, int Q, bool FIRST>
__global__ void __launch_bounds__(NttShape4<B, Q>::THREADS, NttShape4<B, Q>::MINBLOCKS) k_ntt_pass4(NttPass p) {
    typedef NttShape4<B, Q> S;
    extern __shared__ __align__(16) uint4 ntt_sm[];
    const uint32_t t = threadIdx.x, tile = blockIdx.x;
    const uint32_t log_n = p.log_n, s0 = FIRST ? 0u : p.s0;
    const uint32_t mid = FIRST ? 0u : tile & ((1u << (s0 - Q)) - 1u), high = FIRST ? 0u : tile >> (s0 - Q);
    auto gidx = [&](uint32_t e) -> uint32_t {
        const uint32_t v = e >> Q, col = e & ((1u << Q) - 1u);
        if (FIRST) return (Q ? col << (log_n - Q) : 0u) | (tile << B) | v;
        return (high << (s0 + B)) | (v << s0) | (mid << Q) | col;
    };
    const fr_t *__restrict__ tw = p.tw;
    fr_t x[4];
    fr_t w0, w1, w2;
#pragma unroll 1
    for (int r = 0; r < S::ROUNDS; r++) {
        const uint32_t P = (uint32_t)Q + (r == 0 ? 0u : (uint32_t)(S::K0 + 2 * (r - 1)));
        const uint32_t e0 = ((t >> P) << (P + 2)) | (t & ((1u << P) - 1u));
        const uint32_t sb = s0 + P - (uint32_t)Q;
        const bool two = r > 0 || S::K0 == 2;
        const bool trivial = FIRST && r == 0;  // omega^0 everywhere in stage 0, and for the pair (0, 2) in stage 1
        const uint32_t idx0 = sb == 0 ? 0u : (gidx(e0) & ((1u << sb) - 1u)) << (log_n - 1 - sb);
        if (!trivial) {
            w0 = tw[idx0];
            if (two) w1 = tw[idx0 >> 1];
        }
        if (two) w2 = tw[(idx0 >> 1) + (1u << (log_n - 2))];
        if (r == 0) {
#pragma unroll
            for (int xi = 0; xi < 4; xi++) {
                const uint32_t g = gidx(e0 | ((uint32_t)xi << P));
                const uint32_t src = FIRST ? __brev(g) >> (32 - log_n) : g;
                x[xi] = (p.in + (size_t)blockIdx.y * p.in_stride)[src];
                if (FIRST && p.pre_scale) x[xi] = ntt_mul_call(x[xi], p.scale[src]);
            }
        } else {
            __syncthreads();
#pragma unroll
            for (int xi = 0; xi < 4; xi++) x[xi] = ntt_sm_load4<S::TILE>(ntt_sm, e0 | ((uint32_t)xi << P));
        }
        if (r + 1 < S::ROUNDS) {  // the next round's twiddles travel to L1 while this round computes (no registers held)
            const uint32_t Pn = (uint32_t)(Q + S::K0 + 2 * r), sbn = s0 + Pn - (uint32_t)Q;
            const uint32_t e0n = ((t >> Pn) << (Pn + 2)) | (t & ((1u << Pn) - 1u));
            const uint32_t idxn = (gidx(e0n) & ((1u << sbn) - 1u)) << (log_n - 1 - sbn);
            ntt_prefetch(tw + idxn);
            ntt_prefetch(tw + (idxn >> 1));
            ntt_prefetch(tw + (idxn >> 1) + (1u << (log_n - 2)));
        }
        if (trivial) {
            ntt_bfly_trivial(x[0], x[1]);
            ntt_bfly_trivial(x[2], x[3]);
            if (two) {
                ntt_bfly_trivial(x[0], x[2]);
                ntt_bfly_inline(x[1], x[3], w2);
            }
        } else {
            ntt_bfly_inline(x[0], x[1], w0);
            ntt_bfly_inline(x[2], x[3], w0);
            if (two) {
                ntt_bfly_inline(x[0], x[2], w1);
                ntt_bfly_inline(x[1], x[3], w2);
            }
        }
        if (r == S::ROUNDS - 1) {
#pragma unroll
            for (int xi = 0; xi < 4; xi++) {
                const uint32_t g = gidx(e0 | ((uint32_t)xi << P));
                if (p.post_scale == 1) x[xi] = ntt_mul_call(p.scale[0], x[xi]);        // (the product of a value < 2r is canonical)
                else if (p.post_scale == 2) x[xi] = ntt_mul_call(p.scale[g], x[xi]);
                else if (p.last) x[xi] = x[xi].canonical();
                (p.out + (size_t)blockIdx.y * p.out_stride)[g] = x[xi];
            }
        } else {
#pragma unroll
            for (int xi = 0; xi < 4; xi++) ntt_sm_store4<S::TILE>(ntt_sm, e0 | ((uint32_t)xi << P), x[xi]);
        }
    }
}

// (Measured, no gain: __syncwarp() instead of the block barrier for the round transitions whose exchange stays inside a warp --
// 3.13 against 3.11 ms; the barrier waits are covered by the other resident blocks.)
// (Measured and rejected: the same pass as a persistent kernel whose next tile travels HBM -> shared memory by cp.async into a
// second buffer while the current one is computed -- 2^24 fft 3.56 against 3.11 ms: the second buffer halves what is left of L1
// for the twiddles and adds a barrier and a shared-memory round trip per tile.)
// n <= 4: one thread, the reference's loop as is (bit reversal, log n stages; domain.rs:272-315) with the scalings around it
__global__ void k_ntt_tiny(fr_t *a, const fr_t *tw, const fr_t *g_pow, const fr_t *gi_pow, const fr_t *consts, uint32_t log_n, int kind) {
    const uint32_t n = 1u << log_n;
    fr_t x[4];
    for (uint32_t i = 0; i < n; i++) {
        x[i] = a[i];
        if (kind == B200ZK_COSET_FFT) x[i] = x[i] * g_pow[i];
    }
    if (log_n == 2) { fr_t t = x[1]; x[1] = x[2]; x[2] = t; }  // bit reversal of 2-bit indices
    for (uint32_t s = 0; s < log_n; s++) {
        const uint32_t m = 1u << s;
        for (uint32_t k = 0; k < n; k += 2 * m)
            for (uint32_t j = 0; j < m; j++) {
                fr_t t = x[k + j + m] * tw[(size_t)j << (log_n - 1 - s)];
                fr_t lo = x[k + j];
                x[k + j + m] = lo - t;
                x[k + j] = lo + t;
            }
    }
    for (uint32_t i = 0; i < n; i++) {
        if (kind == B200ZK_IFFT) x[i] = x[i] * consts[C_N_INV];
        else if (kind == B200ZK_ICOSET_FFT) x[i] = x[i] * gi_pow[i];
        a[i] = x[i];
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
    cudaFree(t.tw); cudaFree(t.tw_inv); cudaFree(t.g_pow); cudaFree(t.gi_pow); cudaFree(t.consts);
    t = NttTables();
}

int ntt_get_tables(Ctx *ctx, uint32_t log_n, NttTables **out) {
    auto it = ctx->ntt_tables.find(log_n);
    if (it != ctx->ntt_tables.end()) { *out = &it->second; return B200ZK_OK; }
    NttTables t;
    t.log_n = log_n;
    const size_t n = (size_t)1 << log_n;
    const size_t half = n > 1 ? n / 2 : 1;
    B200ZK_CUDA(ctx, cudaMalloc(&t.consts, C_COUNT * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.tw, half * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t.tw_inv, half * sizeof(fr_t)));
    fr_t *c = (fr_t *)t.consts;
    k_ntt_setup<<<1, 1, 0, ctx->stream>>>(c, log_n);
    auto blocks = [](size_t cnt) { return (unsigned)((cnt + 16 * 128 - 1) / (16 * 128)); };
    k_pow_table<<<blocks(half), 128, 0, ctx->stream>>>((fr_t *)t.tw, c + C_OMEGA, nullptr, half);
    k_pow_table<<<blocks(half), 128, 0, ctx->stream>>>((fr_t *)t.tw_inv, c + C_OMEGA_INV, nullptr, half);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { free_tables(t); return set_error(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e)); }
    auto ins = ctx->ntt_tables.emplace(log_n, t);
    *out = &ins.first->second;
    return B200ZK_OK;
}
// the coset tables g^i and g^-i / n (n entries each, natural order), built on the first coset transform of a domain size
static int ntt_coset_tables(Ctx *ctx, NttTables *t) {
    if (t->g_pow) return B200ZK_OK;
    const size_t n = (size_t)1 << t->log_n;
    B200ZK_CUDA(ctx, cudaMalloc(&t->g_pow, n * sizeof(fr_t)));
    B200ZK_CUDA(ctx, cudaMalloc(&t->gi_pow, n * sizeof(fr_t)));
    fr_t *c = (fr_t *)t->consts;
    const unsigned blocks = (unsigned)((n + 16 * 128 - 1) / (16 * 128));
    k_pow_table<<<blocks, 128, 0, ctx->stream>>>((fr_t *)t->g_pow, c + C_G, nullptr, n);
    k_pow_table<<<blocks, 128, 0, ctx->stream>>>((fr_t *)t->gi_pow, c + C_G_INV, c + C_N_INV, n);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

void ntt_free_all_tables(Ctx *ctx) {
    for (auto &kv : ctx->ntt_tables) free_tables(kv.second);
    ctx->ntt_tables.clear();
}

template <int B, int Q, int XB, bool FIRST>
static int ntt_launch(Ctx *ctx, const NttPass &p) {
    typedef NttShape<B, Q, XB> S;
    static bool opted_in[64] = {};  // per device: dynamic shared memory above 48 KiB needs the opt-in
    if (S::SMEM > 48 * 1024 && ctx->device < 64 && !opted_in[ctx->device]) {
        B200ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<B, Q, XB, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        opted_in[ctx->device] = true;
    }
    const unsigned tiles = (unsigned)(((size_t)1 << p.log_n) >> S::T);
    k_ntt_pass<B, Q, XB, FIRST><<<dim3(tiles, p.batch), S::THREADS, S::SMEM, ctx->stream>>>(p);
    ctx->launches++;
    return B200ZK_OK;
}
template <int B, int Q, bool FIRST>
static int ntt_launch4(Ctx *ctx, const NttPass &p) {
    typedef NttShape4<B, Q> S;
    static bool opted_in[64] = {};
    if (S::SMEM > 48 * 1024 && ctx->device < 64 && !opted_in[ctx->device]) {
        B200ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass4<B, Q, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        opted_in[ctx->device] = true;
    }
    const unsigned tiles = (unsigned)(((size_t)1 << p.log_n) >> S::T);
    k_ntt_pass4<B, Q, FIRST><<<dim3(tiles, p.batch), S::THREADS, S::SMEM, ctx->stream>>>(p);
    ctx->launches++;
    return B200ZK_OK;
}
template <bool FIRST>
static int ntt_dispatch(Ctx *ctx, const NttPass &p, uint32_t B, uint32_t Q, uint32_t XB) {
    if (XB == 2 && Q == 1) {
        if (B == 6) return ntt_launch4<6, 1, FIRST>(ctx, p);
        if (B == 7) return ntt_launch4<7, 1, FIRST>(ctx, p);
        if (B == 8) return ntt_launch4<8, 1, FIRST>(ctx, p);
    }
    if (XB == 2 && Q == 2) {
        if (B == 6) return ntt_launch4<6, 2, FIRST>(ctx, p);
        if (B == 7) return ntt_launch4<7, 2, FIRST>(ctx, p);
        if (B == 8) return ntt_launch4<8, 2, FIRST>(ctx, p);
        if (B == 9) return ntt_launch4<9, 2, FIRST>(ctx, p);
    }
#define B200ZK_NTT_CASE(CB, CQ, CX) if (B == CB && Q == CQ && XB == CX) return ntt_launch<CB, CQ, CX, FIRST>(ctx, p);
    // small transforms: two elements per thread
    B200ZK_NTT_CASE(5, 2, 1) B200ZK_NTT_CASE(6, 2, 1) B200ZK_NTT_CASE(7, 2, 1) B200ZK_NTT_CASE(8, 2, 1)
    B200ZK_NTT_CASE(5, 0, 1) B200ZK_NTT_CASE(6, 0, 1) B200ZK_NTT_CASE(7, 0, 1) B200ZK_NTT_CASE(8, 0, 1) B200ZK_NTT_CASE(9, 0, 1)
    if (FIRST) { B200ZK_NTT_CASE(3, 0, 1) B200ZK_NTT_CASE(4, 0, 1) }
#undef B200ZK_NTT_CASE
    return set_error(ctx, B200ZK_ERR_BAD_ARG, "internal: no NTT kernel for this pass shape");
}

// How a transform of 2^log_n elements is cut into passes (host logic only).  Large transforms (work-bound): k_ntt_pass4, passes of
// 8 or 6 stages where possible (whole radix-4 rounds; an odd size gets one pass of 7), two adjacent columns per tile (64-byte runs
// in HBM: 512-element tiles, 128 threads, six blocks per SM -- a barrier waits for four warps instead of eight and more tiles are in
// different phases; 2^24 fft 3.19 -> 3.11 ms against four columns).  Small ones (latency-bound: the time is the serial chain of one
// thread): k_ntt_pass with one butterfly per thread and stage, passes of at most 9 stages, four columns only while they leave two
// tiles per SM.  Returns the number of passes; B[ps] stages and 2^Q[ps] columns per pass.
uint32_t ntt_plan(uint32_t log_n, int large_from, int sm_count, uint32_t batch, uint32_t *B, uint32_t *Q, int *radix4) {
    const size_t n = (size_t)1 << log_n;
    const bool large = log_n >= (uint32_t)large_from && log_n >= 12;
    uint32_t npass = (log_n + 8) / 9;
    for (uint32_t ps = 0; ps < npass; ps++) B[ps] = log_n / npass + (ps < log_n % npass ? 1 : 0);
    if (large && 6 * ((log_n + 7) / 8) <= log_n) {
        npass = (log_n + 7) / 8;
        uint32_t rem = log_n - 6 * npass;
        for (uint32_t ps = 0; ps < npass; ps++) B[ps] = 6;
        if (rem & 1) { B[0] = 7; rem--; }
        for (uint32_t ps = (B[0] == 7 ? 1 : 0); ps < npass && rem >= 2; ps++) { B[ps] = 8; rem -= 2; }
        if (rem >= 2) { B[0] += 2; rem -= 2; }  // (7 -> 9; not reached for log_n <= 30)
    }
    for (uint32_t ps = 0; ps < npass; ps++) {
        Q[ps] = npass == 1 ? 0 : 2;
        if (!large && Q[ps] && ((n >> (B[ps] + Q[ps])) * batch < (size_t)2 * sm_count || B[ps] + Q[ps] > 10)) Q[ps] = 0;
        if (large && B[ps] <= 8) Q[ps] = 1;
    }
    *radix4 = large ? 1 : 0;
    return npass;
}

int ntt_run(Ctx *ctx, void *d_coeffs, uint32_t log_n, int kind) { return ntt_run_batch(ctx, d_coeffs, log_n, kind, 1, (size_t)1 << (log_n < 32 ? log_n : 0)); }

// `batch` equal transforms in one set of launches (blockIdx.y = the vector): vector v starts at d_coeffs + v * stride elements.
// The H blocks of a lock-step batch of proofs are such batches: 16 launches for K proofs instead of 16 K half-empty ones.
int ntt_run_batch(Ctx *ctx, void *d_coeffs, uint32_t log_n, int kind, uint32_t batch, size_t stride) {
    if (batch == 0) return B200ZK_OK;
    if (batch > 65535) return set_error(ctx, B200ZK_ERR_BAD_ARG, "more than 65535 transforms in one batch");
    if (log_n >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");  // domain.rs:59-61
    if (kind < B200ZK_FFT || kind > B200ZK_ICOSET_FFT) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad ntt kind");
    if (log_n == 0) return B200ZK_OK;  // m = 1: omega = 1, m^-1 = 1, g^0 = 1 -> every transform is the identity
    if (log_n > 30) return set_error(ctx, B200ZK_ERR_BAD_ARG, "log_m > 30 not supported on one GPU");
    NttTables *t;
    int st = ntt_get_tables(ctx, log_n, &t);
    if (st) return st;
    const bool coset = kind == B200ZK_COSET_FFT || kind == B200ZK_ICOSET_FFT;
    if (coset && (st = ntt_coset_tables(ctx, t))) return st;
    const size_t n = (size_t)1 << log_n;
    fr_t *A = (fr_t *)d_coeffs;
    const bool inverse = kind == B200ZK_IFFT || kind == B200ZK_ICOSET_FFT;
    const fr_t *tw = (const fr_t *)(inverse ? t->tw_inv : t->tw);
    if (log_n <= 2) {
        for (uint32_t v = 0; v < batch; v++)
            k_ntt_tiny<<<1, 1, 0, ctx->stream>>>(A + v * stride, tw, (const fr_t *)t->g_pow, (const fr_t *)t->gi_pow, (const fr_t *)t->consts, log_n, kind);
        ctx->launches += batch;
        B200ZK_CUDA(ctx, cudaGetLastError());
        return B200ZK_OK;
    }
    uint32_t Bs[8], Qs[8];
    int radix4 = 0;
    const uint32_t npass = ntt_plan(log_n, ctx->ntt_large_from, ctx->sm_count, batch, Bs, Qs, &radix4);
    st = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, (size_t)batch * n * sizeof(fr_t));
    if (st) return st;
    fr_t *S = (fr_t *)ctx->scratch;  // vector v of the batch at S + v * n
    uint32_t s0 = 0;
    const fr_t *src = A;
    for (uint32_t ps = 0; ps < npass; ps++) {
        NttPass p;
        const uint32_t B = Bs[ps], Q = Qs[ps];
        p.log_n = log_n;
        p.s0 = s0;
        p.tw = tw;
        p.pre_scale = (ps == 0 && kind == B200ZK_COSET_FFT) ? 1 : 0;
        p.last = ps + 1 == npass;
        p.post_scale = ps + 1 == npass ? (kind == B200ZK_IFFT ? 1 : kind == B200ZK_ICOSET_FFT ? 2 : 0) : 0;
        p.scale = p.pre_scale ? (const fr_t *)t->g_pow : p.post_scale == 2 ? (const fr_t *)t->gi_pow : (const fr_t *)t->consts + C_N_INV;
        p.in = src;
        p.batch = batch;
        p.in_stride = src == A ? stride : n;
        // The first pass gathers in bit-reversed order, so it writes to the other buffer; every later pass reads and writes the
        // same positions (a tile's reads all precede its first barrier, its writes follow the last), so the middle passes run in
        // place in the scratch buffer and the last one delivers into the caller's vector: no copy for any number of passes.
        fr_t *dst = ps == 0 ? S : ps + 1 == npass ? A : S;
        p.out = dst;
        p.out_stride = dst == A ? stride : n;
        const uint32_t XB = radix4 ? 2 : 1;
        st = ps == 0 ? ntt_dispatch<true>(ctx, p, B, Q, XB) : ntt_dispatch<false>(ctx, p, B, Q, XB);
        if (st) return st;
        src = dst;
        s0 += B;
    }
    if (src != A)  // (single-pass transforms only)
        B200ZK_CUDA(ctx, cudaMemcpy2DAsync(A, stride * sizeof(fr_t), src, n * sizeof(fr_t), n * sizeof(fr_t), batch, cudaMemcpyDeviceToDevice, ctx->stream));
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

// prover.rs:257-265 for ONE of the three vectors: ifft, then coset_fft
int ntt_h_poly_front(Ctx *ctx, void *d_v, uint32_t log_n) {
    int st;
    if ((st = ntt_run(ctx, d_v, log_n, B200ZK_IFFT))) return st;
    return ntt_run(ctx, d_v, log_n, B200ZK_COSET_FFT);
}
// prover.rs:267-287 once a, b, c are on the coset: a = (a * b - c) / z(g), icoset_fft, into_repr of the first m - 1 coefficients
int ntt_h_poly_tail(Ctx *ctx, void *d_a, const void *d_b, const void *d_c, uint32_t log_n, void *d_out_repr) {
    int st;
    const size_t n = (size_t)1 << log_n;
    NttTables *t;
    if ((st = ntt_get_tables(ctx, log_n, &t))) return st;
    ctx->launches += 2;
    k_h_combine<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((fr_t *)d_a, (const fr_t *)d_b, (const fr_t *)d_c, (const fr_t *)t->consts, n);
    if ((st = ntt_run(ctx, d_a, log_n, B200ZK_ICOSET_FFT))) return st;
    if (n > 1) k_into_repr<<<(unsigned)((n - 1 + 255) / 256), 256, 0, ctx->stream>>>((const fr_t *)d_a, (fr_t *)d_out_repr, n - 1);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// The H blocks of K proofs at once: d_abc = a_0 .. a_(K-1) | b_0 .. b_(K-1) | c_0 .. c_(K-1), each vector 2^log_n elements, contiguous;
// d_out_repr = K vectors of 2^log_n slots (the first 2^log_n - 1 of each are the H coefficients in FrRepr form).
int ntt_h_poly_batch(Ctx *ctx, void *d_abc, uint32_t log_n, void *d_out_repr, uint32_t K) {
    int st;
    const size_t n = (size_t)1 << log_n, total = (size_t)K * n;
    fr_t *a = (fr_t *)d_abc, *b = a + total, *c = b + total;
    NttTables *t;
    if ((st = ntt_get_tables(ctx, log_n, &t))) return st;
    if ((st = ntt_run_batch(ctx, a, log_n, B200ZK_IFFT, 3 * K, n))) return st;
    if ((st = ntt_run_batch(ctx, a, log_n, B200ZK_COSET_FFT, 3 * K, n))) return st;
    ctx->launches += 2;
    k_h_combine<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(a, b, c, (const fr_t *)t->consts, total);
    if ((st = ntt_run_batch(ctx, a, log_n, B200ZK_ICOSET_FFT, K, n))) return st;
    k_into_repr<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(a, (fr_t *)d_out_repr, total);  // (slot n - 1 of a vector is not used)
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

int ntt_h_poly(Ctx *ctx, void *d_a, void *d_b, void *d_c, uint32_t log_n, void *d_out_repr) {
    int st;
    void *v[3] = {d_a, d_b, d_c};
    for (int i = 0; i < 3; i++)
        if ((st = ntt_h_poly_front(ctx, v[i], log_n))) return st;
    return ntt_h_poly_tail(ctx, d_a, d_b, d_c, log_n, d_out_repr);
}

}  // namespace b200zk
