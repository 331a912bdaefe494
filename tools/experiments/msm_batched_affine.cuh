// Bucket accumulation by batched affine additions (included by msm_impl.cuh).
//
// The XYZZ accumulation (k_msm_accumulate) already keeps the integer multiplier pipe 93 % busy (ncu), so the only way to
// go faster is to spend fewer field products per added point.  An affine addition P + Q costs one inversion plus
// 2M + 1S; with Montgomery's trick the inversion is shared by a whole batch at 3M per element: 5M + 1S instead of the
// 8M + 2S of madd-2008-s.  Additions must be independent to be batched, so the buckets are reduced as a forest of
// pairwise trees, level by level over the *whole* sorted array:
//
//   level l:  for every bucket b with cnt_l[b] points at P_l[start_l[b] ..): pairs (2i, 2i + 1) are added into
//             P_{l+1}[start_{l+1}[b] + i]; an odd leftover is copied.  cnt_{l+1} = ceil(cnt_l / 2).
//
// One thread owns a run of CH consecutive pairs (across bucket borders): forward pass = running product of the
// denominators (stored, coalesced, in a scratch array), one inversion, backward pass = the additions.  Level 0 reads the
// points through the sorted references (sign bit = negate y); later levels read the previous level's output.  After
// ceil(log2(max load)) levels every bucket holds at most one point.  The identity is the marker (0, 0), which is not on
// y^2 = x^3 + b.  Equal points (doubling) and opposite points (sum = identity) are handled in-line, as the reference's
// add_assign_mixed does (ec.rs:446-526).
#pragma once
#include "ec.cuh"
#include "internal.h"

namespace b200zk {

template <class F>
__device__ __forceinline__ bool ba_is_identity(const Affine<F> &p) { return p.x.is_zero() && p.y.is_zero(); }

template <class F>
__device__ __forceinline__ Affine<F> ba_load(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const Affine<F> *__restrict__ pin,
                                             uint32_t idx) {
    if (sorted) {
        uint32_t e = sorted[idx];
        Affine<F> p = bases[e & 0x7fffffffu];
        if (e >> 31) p.y = p.y.neg();
        return p;
    }
    return pin[idx];
}

// x coordinate only (the forward pass needs y just for the rare equal-x pairs)
template <class F>
__device__ __forceinline__ F ba_load_x(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const Affine<F> *__restrict__ pin, uint32_t idx) {
    if (sorted) return bases[sorted[idx] & 0x7fffffffu].x;
    return pin[idx].x;
}

// kind of a pair and its denominator: 0 = generic (d = x2 - x1), 1 = doubling (d = 2 y1), 2 = result is the identity,
// 3 = P1 is the identity (result P2), 4 = P2 is the identity (result P1); d = 1 when no inversion is needed
template <class F>
__device__ __forceinline__ int ba_classify(const Affine<F> &p1, const Affine<F> &p2, F &d) {
    if (ba_is_identity(p1)) { d = F::one(); return 3; }
    if (ba_is_identity(p2)) { d = F::one(); return 4; }
    d = p2.x - p1.x;
    if (!d.is_zero()) return 0;
    if (p1.y == p2.y && !p1.y.is_zero()) { d = p1.y.dbl(); return 1; }
    d = F::one();
    return 2;
}

static __global__ void k_ba_counts_from_offsets(const uint32_t *__restrict__ offsets, uint32_t n_buckets, uint32_t *__restrict__ cnt, uint32_t *__restrict__ maxcnt) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t c = offsets[b + 1] - offsets[b];
    cnt[b] = c;
    atomicMax(maxcnt, c);
}
static __global__ void k_ba_next_counts(const uint32_t *__restrict__ cnt_in, uint32_t n_buckets, uint32_t *__restrict__ cnt_out, uint32_t *__restrict__ pairs) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t c = cnt_in[b];
    cnt_out[b] = (c + 1) >> 1;
    pairs[b] = c >> 1;
}

// one level of the forest; n_threads = ceil(total_pairs / CH) threads, scratch arrays are [k * n_threads + t]
template <class F>
__global__ void __launch_bounds__(128, 4) k_ba_level(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const Affine<F> *__restrict__ pin,
                                                 const uint32_t *__restrict__ start_in, const uint32_t *__restrict__ cnt_in,
                                                 const uint32_t *__restrict__ pair_start, uint32_t n_buckets, const uint32_t *__restrict__ start_out,
                                                 Affine<F> *__restrict__ pout, F *__restrict__ prefix, uint2 *__restrict__ meta, uint32_t CH,
                                                 uint32_t n_threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = pair_start[n_buckets];
    const uint64_t g0 = (uint64_t)t * CH;
    if (t >= n_threads || g0 >= total) return;
    const uint32_t npairs = (uint32_t)min((uint64_t)CH, (uint64_t)total - g0);
    // bucket of the first pair: pair_start[b] <= g0 < pair_start[b + 1]
    uint32_t lo = 0, hi = n_buckets;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (pair_start[mid] <= (uint32_t)g0) lo = mid; else hi = mid;
    }
    uint32_t b = lo, i = (uint32_t)g0 - pair_start[lo];
    uint32_t bp = cnt_in[b] >> 1, bs = start_in[b], bo = start_out[b];
    F run = F::one();
    for (uint32_t k = 0; k < npairs; k++) {
        if (i >= bp) {  // next bucket that still has pairs: usually the neighbour, otherwise (sparse levels) a binary search
            b++;
            bp = cnt_in[b] >> 1;
            if (bp == 0) {
                const uint32_t g = (uint32_t)g0 + k;
                uint32_t l2 = b, h2 = n_buckets;
                while (h2 - l2 > 1) {
                    uint32_t mid = (l2 + h2) >> 1;
                    if (pair_start[mid] <= g) l2 = mid; else h2 = mid;
                }
                b = l2;
                bp = cnt_in[b] >> 1;
            }
            i = 0;
            bs = start_in[b];
            bo = start_out[b];
        }
        const uint32_t in1 = bs + 2 * i;
        F x1 = ba_load_x(bases, sorted, pin, in1), x2 = ba_load_x(bases, sorted, pin, in1 + 1);
        F d = x2 - x1;
        if (d.is_zero() || x1.is_zero() || x2.is_zero()) {  // rare: equal x, or possibly the identity marker -> full classification
            Affine<F> p1 = ba_load(bases, sorted, pin, in1), p2 = ba_load(bases, sorted, pin, in1 + 1);
            ba_classify(p1, p2, d);
        }
        run = run * d;
        prefix[(size_t)k * n_threads + t] = run;
        meta[(size_t)k * n_threads + t] = make_uint2(in1, bo + i);
        i++;
    }
    F inv = run.inverse();
    for (uint32_t k = npairs; k-- > 0;) {
        const uint2 m = meta[(size_t)k * n_threads + t];
        Affine<F> p1 = ba_load(bases, sorted, pin, m.x), p2 = ba_load(bases, sorted, pin, m.x + 1);
        F d;
        const int kind = ba_classify(p1, p2, d);
        F dinv = k ? inv * prefix[(size_t)(k - 1) * n_threads + t] : inv;
        inv = inv * d;
        Affine<F> r;
        if (kind <= 1) {
            F lam;
            if (kind == 0) {
                lam = (p2.y - p1.y) * dinv;
            } else {
                F xx = p1.x.sqr();
                lam = (xx.dbl() + xx) * dinv;
            }
            r.x = lam.sqr() - p1.x - p2.x;
            r.y = lam * (p1.x - r.x) - p1.y;
        } else if (kind == 2) {
            r.x = F::zero();
            r.y = F::zero();
        } else {
            r = kind == 3 ? p2 : p1;
        }
        pout[m.y] = r;
    }
}

// odd leftovers move to the end of their bucket's next-level segment
template <class F>
__global__ void k_ba_leftover(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const Affine<F> *__restrict__ pin,
                              const uint32_t *__restrict__ start_in, const uint32_t *__restrict__ cnt_in, uint32_t n_buckets,
                              const uint32_t *__restrict__ start_out, Affine<F> *__restrict__ pout) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t c = cnt_in[b];
    if (c & 1) pout[start_out[b] + (c >> 1)] = ba_load(bases, sorted, pin, start_in[b] + c - 1);
}

// every bucket holds 0 or 1 point now -> XYZZ bucket array for the reduction
template <class F>
__global__ void k_ba_finish(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted, const Affine<F> *__restrict__ pin,
                            const uint32_t *__restrict__ start, const uint32_t *__restrict__ cnt, uint32_t n_buckets, XYZZ<F> *__restrict__ buckets) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    XYZZ<F> r = XYZZ<F>::zero();
    if (cnt[b]) {
        Affine<F> p = ba_load(bases, sorted, pin, start[b]);
        if (!ba_is_identity(p)) r = XYZZ<F>::from_affine(p);
    }
    buckets[b] = r;
}

// Host driver: reduces every bucket of the sorted array to one point.  `refs_bound` >= offsets[n_buckets].
// Workspace `ws` must hold ba_workspace_bytes<F>(n_buckets, refs_bound) bytes.
template <class F>
static size_t ba_workspace_bytes(size_t n_buckets, size_t refs_bound) {
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t n1 = refs_bound / 2 + n_buckets + 1, n2 = n1 / 2 + n_buckets + 1;
    size_t pairs0 = (refs_bound + n_buckets) / 2 + 2048;
    return 6 * al((n_buckets + 2) * sizeof(uint32_t)) + al(256) + al(n1 * sizeof(Affine<F>)) + al(n2 * sizeof(Affine<F>)) + al(pairs0 * sizeof(F)) +
           al(pairs0 * sizeof(uint2)) + al((n_buckets / 2048 + 4) * sizeof(uint32_t));
}

template <class F, class ScanFn>
static int ba_accumulate(Ctx *ctx, const Affine<F> *bases, const uint32_t *sorted, const uint32_t *offsets, uint32_t n_buckets, size_t refs_bound,
                         XYZZ<F> *buckets, char *ws, ScanFn scan) {
    cudaStream_t st = ctx->stream;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = ws + off; off += al(bytes); return p; };
    const size_t words = (n_buckets + 2) * sizeof(uint32_t);
    uint32_t *cnt[2] = {(uint32_t *)take(words), (uint32_t *)take(words)};
    uint32_t *start[2] = {(uint32_t *)take(words), (uint32_t *)take(words)};
    uint32_t *pairs = (uint32_t *)take(words), *pair_start = (uint32_t *)take(words);
    uint32_t *maxcnt = (uint32_t *)take(256);
    const size_t n1 = refs_bound / 2 + n_buckets + 1, n2 = n1 / 2 + n_buckets + 1, pairs0 = (refs_bound + n_buckets) / 2 + 2048;
    Affine<F> *pbuf[2] = {(Affine<F> *)take(n1 * sizeof(Affine<F>)), (Affine<F> *)take(n2 * sizeof(Affine<F>))};
    F *prefix = (F *)take(pairs0 * sizeof(F));
    uint2 *meta = (uint2 *)take(pairs0 * sizeof(uint2));
    uint32_t *sums = (uint32_t *)take((n_buckets / 2048 + 4) * sizeof(uint32_t));
    const unsigned bb = (unsigned)((n_buckets + 255) / 256);

    B200ZK_CUDA(ctx, cudaMemsetAsync(maxcnt, 0, sizeof(uint32_t), st));
    k_ba_counts_from_offsets<<<bb, 256, 0, st>>>(offsets, n_buckets, cnt[0], maxcnt);
    uint32_t h_max = 0;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(&h_max, maxcnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(st));  // the number of levels depends on the fullest bucket
    uint32_t levels = 0;
    while ((1ull << levels) < h_max) levels++;
    ctx->launches += 1;

    const uint32_t *cur_start = offsets, *cur_cnt = cnt[0];
    const Affine<F> *pin = nullptr;
    const uint32_t *refs = sorted;
    size_t bound = refs_bound;  // upper bound of the points at the current level
    int ci = 0;
    const size_t target_threads = (size_t)ctx->sm_count * 384 * 2;
    for (uint32_t l = 0; l < levels; l++) {
        uint32_t *ncnt = cnt[ci ^ 1], *nstart = start[l & 1];
        k_ba_next_counts<<<bb, 256, 0, st>>>(cur_cnt, n_buckets, ncnt, pairs);
        ctx->launches += 1 + scan(ncnt, n_buckets, nstart, sums) + scan(pairs, n_buckets, pair_start, sums);
        const size_t pair_bound = bound / 2;
        uint32_t CH = (uint32_t)((pair_bound + target_threads - 1) / target_threads);
        CH = CH < 128 ? 128 : CH > 1024 ? 1024 : CH;
        const uint32_t n_threads = (uint32_t)((pair_bound + CH - 1) / CH) + 1;
        Affine<F> *pout = pbuf[l & 1];
        k_ba_level<F><<<(n_threads + 127) / 128, 128, 0, st>>>(bases, refs, pin, cur_start, cur_cnt, pair_start, n_buckets, nstart, pout, prefix, meta, CH,
                                                             n_threads);
        k_ba_leftover<F><<<bb, 256, 0, st>>>(bases, refs, pin, cur_start, cur_cnt, n_buckets, nstart, pout);
        ctx->launches += 2;
        cur_start = nstart;
        cur_cnt = ncnt;
        ci ^= 1;
        pin = pout;
        refs = nullptr;
        bound = bound / 2 + n_buckets;
    }
    k_ba_finish<F><<<bb, 256, 0, st>>>(bases, refs, pin, cur_start, cur_cnt, n_buckets, buckets);
    ctx->launches += 1;
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
