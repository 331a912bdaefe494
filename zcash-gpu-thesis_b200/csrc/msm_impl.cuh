// Pippenger multi-scalar multiplication over BLS12-381 G1 / G2 on sm_100a: the device implementation of
// bellman::multiexp (bellman/src/multiexp.rs:140-335).
//
// The reference runs one CPU task per c-bit window; each task walks all (exponent, density) pairs, adds the
// consumed base into bucket[digit-1] with a mixed Jacobian add (multiexp.rs:174-196), folds the buckets with
// a running sum (multiexp.rs:202-206) and the windows are joined by c doublings each (multiexp.rs:223-229).
// The B200 pipeline keeps the same decomposition (windows x buckets) but turns the scalar loop inside out:
//
//   1. k_msm_digits<COUNT>   per exponent: density rank -> base index, error detection (EOF / identity, the
//                            Source semantics of multiexp.rs:42-68), signed c-bit digits; histogram per
//                            (window, bucket) with global reductions
//   2. scan                  exclusive prefix sum of the histogram = bucket offsets
//   3. k_msm_digits<SCATTER> counting-sort of (base index, sign) by bucket
//   4. k_msm_accumulate      one thread per bucket (G1) / k_msm_accumulate_pair: one lane pair per bucket, one Fq component of
//                            every Fq2 coordinate per lane (G2): XYZZ accumulator in registers, mixed adds (the hot loop);
//                            oversized buckets are cut into tasks and folded by k_msm_combine_small / _big
//   5. k_msm_ladder_step     bucket reduction sum (i+1)*S_i as a log-depth ladder on lane pairs, k_msm_ladder_final
//   6. k_msm_window_combine  Horner over the windows (c doublings each), Jacobian result + status word
// With the precomputed table 2^(c w) P (b200zk_bases_precompute) all windows feed ONE bucket set: steps 5 and 6 run once.
//
// Signed digits halve the bucket count (bucket ids 1..2^(c-1), negative digits add -P).  exp == 0 is skipped
// and exp == 1 needs no special case: it is digit 1 of window 0.  The group result equals the reference's;
// the Jacobian representative differs (see ec.cuh), parity is checked after into_affine.
#pragma once
#include <algorithm>

#include <cstdlib>

#include "ec.cuh"
#include "internal.h"

namespace b200zk {

static constexpr uint32_t NO_POS = 0xffffffffu;

// ------------------------------------------------------------------------------------------------ u32 exclusive scan
// 3-kernel scan (block sums, scan of block sums, block scan + offset); input is u32 counts or u8 density flags.
static constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

template <class T>
__device__ __forceinline__ uint32_t scan_val(const T *in, size_t i, size_t n) {
    if (i >= n) return 0;
    if (sizeof(T) == 1) return in[i] != 0 ? 1u : 0u;
    return (uint32_t)in[i];
}
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += y; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    return base + x - v;
}
template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const T *__restrict__ in, size_t n, uint32_t *__restrict__ sums) {
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) s += scan_val(in, base + i, n);
    uint32_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *sums, size_t nblocks, uint32_t *total_out) {
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nblocks; b0 += 1024) {
        size_t i = b0 + threadIdx.x;
        uint32_t v = i < nblocks ? sums[i] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan(v, &total);
        uint32_t c = carry;
        if (i < nblocks) sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const T *__restrict__ in, size_t n, const uint32_t *__restrict__ sums,
                                                            uint32_t *__restrict__ out, uint32_t *__restrict__ out2) {
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = scan_val(in, base + i, n); s += v[i]; }
    uint32_t ex = block_excl_scan(s, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) { out[base + i] = ex; if (out2) out2[base + i] = ex; }
        ex += v[i];
    }
}
// out[i] = exclusive prefix of in; out[n] (and *total) = sum.  `sums` needs ceil(n / SCAN_BLOCK) + 1 words.
template <class T>
static int scan_u32(cudaStream_t st, const T *in, size_t n, uint32_t *out, uint32_t *out2, uint32_t *sums) {
    if (n == 0) { cudaMemsetAsync(out, 0, sizeof(uint32_t), st); if (out2) cudaMemsetAsync(out2, 0, sizeof(uint32_t), st); return 0; }
    size_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    k_scan_block_sums<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
    k_scan_sums<<<1, 1024, 0, st>>>(sums, nb, out + n);
    k_scan_apply<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums, out, out2);
    if (out2) cudaMemcpyAsync(out2 + n, out + n, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
    return 3;
}

// ------------------------------------------------------------------------------------------------ digits / sort
struct MsmShape {
    uint32_t c, W, B;  // window bits, windows (c*W >= 256), buckets per window = 2^(c-1)
    uint32_t pre_n;    // 0, or the stride of the precomputed base table (then all windows share one bucket set)
    uint32_t sets;     // bucket sets per multiexp: 1 with the precomputed table, W without
};
static constexpr uint32_t ST_PER_K = 8;  // offset of the per-multiexp status triples inside the status area
static constexpr int ST_NONCANONICAL = 7;  // word 7 of the status area: some scalar of the call needs more than c W bits

__device__ __forceinline__ uint32_t extract_bits(const uint32_t *s, uint32_t pos, uint32_t c) {
    uint32_t limb = pos >> 5, sh = pos & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1);
}

// MODE 0: histogram, 1: scatter.  One thread per exponent (multiexp.rs:174-196 turned inside out).
template <int MODE>
__global__ void __launch_bounds__(256) k_msm_digits(const uint32_t *__restrict__ scalars, size_t n_exp, const uint8_t *__restrict__ density,
                                                   const uint32_t *__restrict__ rank, size_t base_offset, size_t n_bases,
                                                   const uint8_t *__restrict__ base_inf, MsmShape sh, MsmBatch batch,
                                                   uint32_t *__restrict__ counts_or_cursor, uint32_t *__restrict__ sorted,
                                                   uint32_t *__restrict__ status, uint32_t slot_lo, uint32_t slot_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_exp) return;
    const uint32_t k = blockIdx.y;  // which multiexp of the batch
    status += ST_PER_K + 3 * k;
    if (density && !density[k * batch.density_stride + i]) return;  // no base consumed (multiexp.rs:175)
    size_t idx = base_offset + (rank ? rank[k * (n_exp + 1) + i] : i);
    if (idx >= n_bases) {  // Source::{skip, add_assign_mixed} both fail first on an exhausted source (multiexp.rs:44, 60)
        if (MODE == 0) atomicMin(&status[0], (uint32_t)i);
        return;
    }
    uint32_t s[8];
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * (k * batch.scalar_stride + i));
    uint4 lo = sp[0], hi = sp[1];
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;  // exp == zero: skip(1)
    if (base_inf && base_inf[idx]) {  // consumed identity base (multiexp.rs:48-50)
        if (MODE == 0) atomicMin(&status[1], (uint32_t)i);
        return;
    }
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sh.W; w++) {
        uint32_t raw = extract_bits(s, w * sh.c, sh.c) + carry;
        uint32_t neg = raw > sh.B;
        uint32_t d = neg ? (1u << sh.c) - raw : raw;
        carry = neg;
        if (d == 0) continue;
        // with precomputed 2^(c w) * P the digit of window w is a digit of window 0 for the point w * n + idx
        uint32_t slot = (k * sh.sets + (sh.pre_n ? 0u : w)) * sh.B + d - 1;
        if (slot < slot_lo || slot >= slot_hi) continue;  // another pass of the scatter owns this bucket range
        if (MODE == 0) {
            atomicAdd(&counts_or_cursor[slot], 1u);
        } else {
            uint32_t pos = atomicAdd(&counts_or_cursor[slot], 1u);
            sorted[pos] = (uint32_t)(idx + (size_t)w * sh.pre_n) | (neg << 31);
        }
    }
    // A carry out of the top window only happens for a scalar of 256 significant bits when c divides 256 (c W = 256): no canonical
    // FrRepr (r < 2^255) gets here, but the ABI takes any 4 x u64, so the multiexp reports it instead of dropping 2^256 P silently
    if (MODE == 0 && carry) status[ST_NONCANONICAL - (int)(ST_PER_K + 3 * k)] = 1u;
}

// ------------------------------------------------------------------------------------------------ bucket accumulation
// resident blocks per SM the G1 accumulation kernel is compiled for: 4 x 128 threads at 128 registers (16 warps; round 2: 73.0 ms at
// 2^24 against 73.9 with 3 x 128 at 168 registers -- the extra warps are worth more than the ~90 extra bytes of spills; 7 x 64
// threads 73.7).  G2 runs on lane pairs (k_msm_accumulate_pair).
template <class F> struct AccShape {
    static constexpr unsigned THREADS = 128;
    static constexpr unsigned MINBLOCKS = 4;
};
// Oversized buckets (witness scalars are full of 0/1/small values: half of a Sapling witness lands in bucket 1 of
// window 0) are split into tasks of at most `cap` points so that no thread walks a bucket alone; the partial sums of a
// split bucket are folded by k_msm_combine_small / k_msm_combine_big.  With uniform scalars only the few buckets that the
// short top window feeds are split (the cap is 1.25 x the actual mean load + 8, k_msm_pick_cap).
static constexpr uint32_t SPLIT_SERIAL_MAX = 32;  // up to here one thread per bucket beats a block per bucket
// The chain length above which a bucket is split, from the histogram itself: 1.25 x the ACTUAL mean load + 8, never above
// the host's bound.  The nominal mean n W / buckets overestimates the typical load of a witness (half of its scalars are
// 0 or 1 and contribute one digit or none), and the one or two heavy buckets would be cut into chains several times longer
// than everybody else's: the whole kernel then waits for them (7.7 ms instead of ~3 for the batched G2 multiexp of 8 proofs).
// atomicAdd(&counter[key], 1) for a warp whose lanes mostly share a few keys (bucket loads cluster around the mean): one
// atomic per distinct key of the warp, every lane gets its own position.  All lanes of the warp must call it.
__device__ __forceinline__ uint32_t warp_grouped_increment(uint32_t *counter, uint32_t key, bool active) {
    const unsigned lanes = __ballot_sync(0xffffffffu, active);
    uint32_t pos = 0;
    if (active) {
        const unsigned peers = __match_any_sync(lanes, key);
        const unsigned lane = threadIdx.x & 31u, leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counter[key], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        pos = base + __popc(peers & ((1u << lane) - 1u));
    }
    return pos;
}
static __global__ void k_msm_pick_cap(const uint32_t *__restrict__ offsets, uint32_t n_buckets, uint32_t cap_host, uint32_t *__restrict__ cap_out) {
    const uint32_t mean = offsets[n_buckets] / n_buckets;
    *cap_out = max(8u, min(cap_host, mean + mean / 4 + 8));
}
static __global__ void k_msm_count_tasks(const uint32_t *__restrict__ offsets, uint32_t n_buckets, const uint32_t *__restrict__ cap_dev, uint32_t *__restrict__ task_cnt,
                                         uint32_t *__restrict__ split_list, uint32_t *__restrict__ n_split, uint32_t *__restrict__ size_hist, uint32_t serial_max) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = b < n_buckets;
    const uint32_t cap = *cap_dev;
    uint32_t cnt = 0;
    if (live) {
        cnt = offsets[b + 1] - offsets[b];
        uint32_t tasks = cnt <= cap ? 0u : (cnt + cap - 1) / cap;
        task_cnt[b] = tasks;
        // n_split[0]: buckets with a few partial sums, listed from the front; n_split[1]: buckets with many, listed from the back
        if (tasks > serial_max) split_list[n_buckets - 1 - atomicAdd(n_split + 1, 1u)] = b;
        else if (tasks) split_list[atomicAdd(n_split, 1u)] = b;
    }
    warp_grouped_increment(size_hist, cap - min(cnt, cap), live);  // key 0 = fullest
}
// Counting sort of the bucket ids by load (fullest first): the 32 buckets of a warp then carry (almost) the same number of
// points, so no lane idles while its neighbours finish (ncu: 29.2 of 32 lanes active before this, Poisson spread of the loads).
static __global__ void k_msm_order_buckets(const uint32_t *__restrict__ offsets, uint32_t n_buckets, const uint32_t *__restrict__ cap_dev, uint32_t *__restrict__ size_cursor,
                                           uint32_t *__restrict__ order) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = b < n_buckets;
    const uint32_t cap = *cap_dev;
    const uint32_t cnt = live ? offsets[b + 1] - offsets[b] : 0u;
    const uint32_t pos = warp_grouped_increment(size_cursor, cap - min(cnt, cap), live);
    if (live) order[pos] = b;
}

template <class F>
__global__ void __launch_bounds__(AccShape<F>::THREADS, AccShape<F>::MINBLOCKS) k_msm_accumulate(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                       const uint32_t *__restrict__ offsets, uint32_t n_buckets, const uint32_t *__restrict__ task_cnt,
                                                       const uint32_t *__restrict__ task_off, const uint32_t *__restrict__ order,
                                                       uint32_t max_tasks, XYZZ<F> *__restrict__ buckets, XYZZ<F> *__restrict__ partials) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t beg, end;
    XYZZ<F> *dst;
    // the split tasks (the longest chains) come first in the grid so they start in the first wave
    if (t >= max_tasks) {
        t -= max_tasks;
        if (t >= n_buckets) return;
        uint32_t b = order[t];
        if (task_cnt[b]) return;  // handled by its split tasks
        beg = offsets[b];
        end = offsets[b + 1];
        dst = buckets + b;
    } else {
        uint32_t task = t;
        if (task >= task_off[n_buckets]) return;
        // bucket b with task_off[b] <= task < task_off[b + 1]
        uint32_t lo = 0, hi = n_buckets;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (task_off[mid] <= task) lo = mid; else hi = mid;
        }
        // the bucket's points are shared evenly among its tasks (equal chains, not cap, cap, ..., remainder)
        const uint32_t j = task - task_off[lo], tasks = task_off[lo + 1] - task_off[lo];
        const uint32_t first = offsets[lo], cnt = offsets[lo + 1] - first;
        beg = first + (uint32_t)(((uint64_t)cnt * j) / tasks);
        end = first + (uint32_t)(((uint64_t)cnt * (j + 1)) / tasks);
        dst = partials + task;
    }
    XYZZ<F> acc = XYZZ<F>::zero();
    uint32_t e = beg < end ? sorted[beg] : 0u;
    for (uint32_t k = beg; k < end; k++) {
        // the next reference is read one addition ahead and its point is requested from HBM while this addition runs
        // (a prefetch costs no registers; the gather is a ~1 us dependent load in front of ~10 us of arithmetic)
        const uint32_t e_next = k + 1 < end ? sorted[k + 1] : 0u;
        if (k + 1 < end) {
            const char *nxt = reinterpret_cast<const char *>(bases + (e_next & 0x7fffffffu));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + sizeof(Affine<F>) - 1));
        }
        Affine<F> p = bases[e & 0x7fffffffu];
        acc.add_mixed(p, (e >> 31) != 0);
        e = e_next;
    }
    *dst = acc;
}

// The same accumulation for G2 with TWO lanes per bucket chain: each lane holds one Fq component of every Fq2 coordinate (fq2h_t,
// fq2.cuh).  Round 1 ran one thread per chain at 255 registers with a 1.5 KB stack frame (97 M local loads per 2^20 multiexp);
// a lane of a pair needs about what a G1 thread needs.  Layout in memory is unchanged (x.c0 x.c1 y.c0 y.c1 | X Y ZZ ZZZ).
struct AccPairShape {
    static constexpr unsigned THREADS = 128, MINBLOCKS = 3;
};
static __global__ void __launch_bounds__(AccPairShape::THREADS, AccPairShape::MINBLOCKS)
k_msm_accumulate_pair(const Affine<fq2_t> *__restrict__ bases, const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ offsets, uint32_t n_buckets,
                      const uint32_t *__restrict__ task_cnt, const uint32_t *__restrict__ task_off, const uint32_t *__restrict__ order, uint32_t max_tasks,
                      XYZZ<fq2_t> *__restrict__ buckets, XYZZ<fq2_t> *__restrict__ partials) {
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;  // both lanes of a pair take the same path everywhere below
    const unsigned role = threadIdx.x & 1u;
    uint32_t beg, end;
    XYZZ<fq2_t> *dst;
    if (t >= max_tasks) {
        t -= max_tasks;
        if (t >= n_buckets) return;
        uint32_t b = order[t];
        if (task_cnt[b]) return;
        beg = offsets[b];
        end = offsets[b + 1];
        dst = buckets + b;
    } else {
        uint32_t task = t;
        if (task >= task_off[n_buckets]) return;
        uint32_t lo = 0, hi = n_buckets;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (task_off[mid] <= task) lo = mid; else hi = mid;
        }
        const uint32_t j = task - task_off[lo], tasks = task_off[lo + 1] - task_off[lo];
        const uint32_t first = offsets[lo], cnt = offsets[lo + 1] - first;
        beg = first + (uint32_t)(((uint64_t)cnt * j) / tasks);
        end = first + (uint32_t)(((uint64_t)cnt * (j + 1)) / tasks);
        dst = partials + task;
    }
    XYZZ<fq2h_t> acc = XYZZ<fq2h_t>::zero();
    uint32_t e = beg < end ? sorted[beg] : 0u;
    for (uint32_t k = beg; k < end; k++) {
        const uint32_t e_next = k + 1 < end ? sorted[k + 1] : 0u;
        if (k + 1 < end) {  // the 192-byte point spans two lines: one prefetch per lane
            const char *nxt = reinterpret_cast<const char *>(bases + (e_next & 0x7fffffffu));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (role ? sizeof(Affine<fq2_t>) - 1 : 0)));
        }
        const fq_t *src = reinterpret_cast<const fq_t *>(bases + (e & 0x7fffffffu));
        Affine<fq2h_t> p;
        p.x.c = src[role];
        p.y.c = src[2 + role];
        acc.add_mixed(p, (e >> 31) != 0);
        e = e_next;
    }
    fq_t *out = reinterpret_cast<fq_t *>(dst);
    out[role] = acc.x.c;
    out[2 + role] = acc.y.c;
    out[4 + role] = acc.zz.c;
    out[6 + role] = acc.zzz.c;
}
template <class F>
static void msm_launch_accumulate(cudaStream_t st, const void *point_table, const uint32_t *sorted, const uint32_t *offsets, size_t nbk, const uint32_t *task_cnt,
                                  const uint32_t *task_off, const uint32_t *order, size_t max_tasks, XYZZ<F> *buckets, XYZZ<F> *partials) {
    if constexpr (sizeof(F) > 48) {
        // measured at 2^22: 3 blocks of 128 per SM (168 registers, small spills) 63.95 ms, 2 blocks (254 registers, none) 66.0 ms,
        // one thread per chain (k_msm_accumulate<fq2_t>, 255 registers, 1.5 KB frame) 67.8 ms
        const size_t threads = 2 * (nbk + max_tasks);
        k_msm_accumulate_pair<<<(unsigned)((threads + AccPairShape::THREADS - 1) / AccPairShape::THREADS), AccPairShape::THREADS, 0, st>>>(
            (const Affine<fq2_t> *)point_table, sorted, offsets, (uint32_t)nbk, task_cnt, task_off, order, (uint32_t)max_tasks, buckets, partials);
    } else {
        k_msm_accumulate<F><<<(unsigned)((nbk + max_tasks + AccShape<F>::THREADS - 1) / AccShape<F>::THREADS), AccShape<F>::THREADS, 0, st>>>(
            (const Affine<F> *)point_table, sorted, offsets, (uint32_t)nbk, task_cnt, task_off, order, (uint32_t)max_tasks, buckets, partials);
    }
}

// Folds the partial sums of the split buckets.  A bucket with few partials (small multiexps cut every chain, so most
// buckets have 2-4) is summed by one thread; one with many (bucket 1 of a witness, the few buckets the short top window
// feeds) by a whole block: threads stride over its partials, then a shared-memory tree.
template <class F>
__global__ void __launch_bounds__(64) k_msm_combine_small(const uint32_t *__restrict__ split_list, const uint32_t *__restrict__ n_split,
                                                         const uint32_t *__restrict__ task_cnt, const uint32_t *__restrict__ task_off,
                                                         const XYZZ<F> *__restrict__ partials, XYZZ<F> *__restrict__ buckets) {
    // one thread per split bucket: there are as many of them as buckets, so this kernel is throughput-bound and one thread per
    // chain uses the multiplier best (lane pairs -- used by the reduction ladder below -- measured slower here and in the
    // accumulation, at every size down to 2^12: G1 2^16 0.94 -> 1.08 ms, G2 61 300 points 2.8 -> 3.4 ms)
    const uint32_t total = n_split[0];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
        const uint32_t b = split_list[s], cnt = task_cnt[b], off = task_off[b];
        XYZZ<F> acc = partials[off];
        for (uint32_t j = 1; j < cnt; j++) acc.add(partials[off + j]);
        buckets[b] = acc;
    }
}
static constexpr uint32_t COMBINE_BIG_THREADS = 256;  // 128 lane pairs
template <class F>
__global__ void __launch_bounds__(COMBINE_BIG_THREADS) k_msm_combine_big(const uint32_t *__restrict__ split_list, const uint32_t *__restrict__ n_split,
                                                                        uint32_t n_buckets, const uint32_t *__restrict__ task_cnt,
                                                                        const uint32_t *__restrict__ task_off, const XYZZ<F> *__restrict__ partials,
                                                                        XYZZ<F> *__restrict__ buckets) {
    extern __shared__ __align__(16) unsigned char combine_smem[];
    XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(combine_smem);
    constexpr uint32_t PAIRS = COMBINE_BIG_THREADS / 2;
    const uint32_t total = n_split[1], pj = threadIdx.x >> 1;
    for (uint32_t s = blockIdx.x; s < total; s += gridDim.x) {
        const uint32_t b = split_list[n_buckets - 1 - s], cnt = task_cnt[b], off = task_off[b];
        PairXYZZ<F> acc = PairXYZZ<F>::zero();
        for (uint32_t j0 = 0; j0 < cnt; j0 += PAIRS) {  // block-uniform trip count
            const uint32_t j = j0 + pj;
            const PairXYZZ<F> o = j < cnt ? PairXYZZ<F>::load(partials + off + j) : PairXYZZ<F>::zero();
            acc.add(o);
        }
        acc.store(&sm[pj]);
        __syncthreads();
        for (uint32_t stride = PAIRS / 2; stride > 0; stride >>= 1) {
            const bool on = pj < stride;
            PairXYZZ<F> x = on ? PairXYZZ<F>::load(&sm[pj]) : PairXYZZ<F>::zero();
            const PairXYZZ<F> y = on ? PairXYZZ<F>::load(&sm[pj + stride]) : PairXYZZ<F>::zero();
            x.add(y);
            __syncthreads();
            if (on) x.store(&sm[pj]);
            __syncthreads();
        }
        if (pj == 0) PairXYZZ<F>::load(&sm[0]).store(buckets + b);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ bucket reduction
// The reference folds a window's buckets with a running sum (multiexp.rs:202-206): 2 (2^c - 1) dependent additions.  On the GPU
// a dependent point addition costs ~10 us (G1) / ~25 us (G2) of latency in one thread, so the reduction is arranged for DEPTH:
//
//   sum_i (i + 1) E_i  =  R + sum_j 2^j T_j ,   R = sum_i E_i ,   T_j = sum over {i : bit j of i} E_i          (E_i = bucket i + 1)
//
// computed as a pairwise ladder: C^0 = E, C^(s+1)_k = C^s_2k + C^s_(2k+1) (so C^s_k is the sum of the buckets whose index,
// shifted right by s, is k), and T_s = the sum of the ODD entries of C^s, itself a pairwise tree that starts in the same step
// and stays one halving behind.  Every step is one launch in which every thread does ONE addition; after log2(B) - 1 steps all
// T_j and R are known.  2 B additions in total (as many as the running sum), log2(B) deep instead of 2 B.
// (Round 1 used an 8-ary (R, A) tree above 4096 entries and computed all masked sums at once below: ~3 x deeper, 17 x the
// additions on the last 4096 entries.)
// Every addition of the ladder is done by a PAIR of lanes (ec.cuh PairXYZZ): the steps are pure latency.
template <class F>
__global__ void __launch_bounds__(64) k_msm_ladder_step(const XYZZ<F> *__restrict__ Cin, const XYZZ<F> *__restrict__ Din, uint32_t m, uint32_t s, uint32_t sets,
                                                       XYZZ<F> *__restrict__ Cout, XYZZ<F> *__restrict__ Dout) {
    // per bucket set: m / 2 pair sums of C, and (s + 1) arrays of nD = m / 4 pair sums of the odd-entry trees
    const uint32_t nC = m >> 1, nD = m >> 2, per_set = nC + (s + 1) * nD;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;  // two lanes per addition
    const bool live = t < per_set * sets;
    const XYZZ<F> *p0 = nullptr, *p1 = nullptr;
    XYZZ<F> *dst = nullptr;
    if (live) {
        const uint32_t w = t / per_set, u = t % per_set;
        const XYZZ<F> *C = Cin + (size_t)w * m;
        if (u < nC) {
            p0 = C + 2 * u; p1 = C + 2 * u + 1;
            dst = Cout + (size_t)w * nC + u;
        } else {
            const uint32_t j = (u - nC) / nD, k = (u - nC) % nD;
            if (j == s) {  // the tree of T_s starts from the odd entries of C^s
                p0 = C + 4 * k + 1; p1 = C + 4 * k + 3;
            } else {       // the tree of T_j, j < s: one more halving (its arrays hold 2 nD entries in Din)
                const XYZZ<F> *D = Din + ((size_t)w * s + j) * (2 * nD);
                p0 = D + 2 * k; p1 = D + 2 * k + 1;
            }
            dst = Dout + ((size_t)w * (s + 1) + j) * nD + k;
        }
    }
    PairXYZZ<F> x = live ? PairXYZZ<F>::load(p0) : PairXYZZ<F>::zero();
    const PairXYZZ<F> y = live ? PairXYZZ<F>::load(p1) : PairXYZZ<F>::zero();
    x.add(y);
    if (live) x.store(dst);
}
// The last step and the weights: C (2 entries) and the finished trees D (nb - 1 single entries) of one bucket set ->
// value = (C_0 + C_1) + 2^(nb-1) C_1 + sum_(j < nb-1) 2^j T_j, written as the (R, A) pair (value, 0) that k_msm_window_combine takes.
// 64 threads = 32 lane pairs per bucket set: pair j raises its term to its power of two by j doublings, a shared-memory tree
// adds the terms.
template <class F>
__global__ void __launch_bounds__(64) k_msm_ladder_final(const XYZZ<F> *__restrict__ C, const XYZZ<F> *__restrict__ D, uint32_t nb, uint32_t sets, XYZZ<F> *__restrict__ outR,
                                                        XYZZ<F> *__restrict__ outA) {
    __shared__ XYZZ<F> sm[32];
    const uint32_t w = blockIdx.x, j = threadIdx.x >> 1;
    const XYZZ<F> *c2 = C + (size_t)w * 2;
    const XYZZ<F> *src = nullptr;
    if (j + 1 < nb) src = D + (size_t)w * (nb - 1) + j;  // T_j
    else if (j + 1 == nb || j == nb) src = c2 + 1;        // T_(nb-1) = the odd entry of the last C; R = C_0 + C_1 starts from C_1 too
    PairXYZZ<F> acc = src ? PairXYZZ<F>::load(src) : PairXYZZ<F>::zero();
    {
        const PairXYZZ<F> c0 = j == nb ? PairXYZZ<F>::load(c2) : PairXYZZ<F>::zero();
        acc.add(c0);  // only the R pair adds something
    }
    for (uint32_t d = 0; d + 1 < nb; d++) {  // uniform trip count: the shuffles inside need every lane
        PairXYZZ<F> t = acc;
        t.dbl();
        const bool mine = j < nb && d < j;
        acc.a = lane_select(mine, t.a, acc.a);
        acc.b = lane_select(mine, t.b, acc.b);
    }
    acc.store(&sm[j]);
    __syncthreads();
    for (uint32_t stride = 16; stride > 0; stride >>= 1) {
        const bool on = j < stride;
        PairXYZZ<F> x = on ? PairXYZZ<F>::load(&sm[j]) : PairXYZZ<F>::zero();
        const PairXYZZ<F> y = on ? PairXYZZ<F>::load(&sm[j + stride]) : PairXYZZ<F>::zero();
        x.add(y);
        __syncthreads();
        if (on) x.store(&sm[j]);
        __syncthreads();
    }
    if (j == 0) {
        PairXYZZ<F>::load(&sm[0]).store(&outR[w]);
        PairXYZZ<F>::zero().store(&outA[w]);
    }
}

// multiexp.rs:223-229: higher = 2^c * higher + this, from the top window down; then the status word.
template <class F>
__global__ void __launch_bounds__(32) k_msm_window_combine(const XYZZ<F> *__restrict__ R, const XYZZ<F> *__restrict__ A, MsmShape sh, Jacobian<F> *__restrict__ out,
                                                          uint32_t *__restrict__ status, uint32_t *__restrict__ status_out) {
    // one warp per multiexp of the batch; its first lane pair does the work (the other pairs run along on the identity)
    const uint32_t k = blockIdx.x;  // multiexp of the batch: its bucket sets are [k * sets, (k + 1) * sets)
    const bool mine = threadIdx.x < 2;
    R += (size_t)k * sh.sets;
    A += (size_t)k * sh.sets;
    PairXYZZ<F> acc = PairXYZZ<F>::zero();
    for (uint32_t w = sh.sets; w-- > 0;) {
        if (w + 1 < sh.sets) for (uint32_t d = 0; d < sh.c; d++) acc.dbl();
        PairXYZZ<F> t = mine ? PairXYZZ<F>::load(A + w) : PairXYZZ<F>::zero();
        const PairXYZZ<F> r = mine ? PairXYZZ<F>::load(R + w) : PairXYZZ<F>::zero();
        t.add(r);
        acc.add(t);
    }
    const bool zero = acc.is_zero();
    F zz;
    const F h = acc.to_jacobian_half(&zz);  // X ZZ | Y ZZZ
    if (threadIdx.x == 0) {
        out[k].x = zero ? F::zero() : h;
        out[k].z = zero ? F::zero() : zz;
        status += ST_PER_K + 3 * k;
        uint32_t eof = status[0], ident = status[1];
        uint32_t st = B200ZK_OK;
        if (eof != NO_POS || ident != NO_POS) st = eof < ident ? B200ZK_ERR_UNEXPECTED_EOF : B200ZK_ERR_UNEXPECTED_IDENTITY;
        else if (status[ST_NONCANONICAL - (int)(ST_PER_K + 3 * k)]) st = B200ZK_ERR_BAD_ARG;
        status[2] = st;
        if (status_out) status_out[k] = st;
    } else if (threadIdx.x == 1) {
        out[k].y = zero ? F::one() : h;
    }
}

template <class F>
__global__ void k_write_zero_point(Jacobian<F> *out, uint32_t *status, uint32_t *status_out) {
    const uint32_t k = blockIdx.x;
    out[k] = Jacobian<F>::zero();
    status[ST_PER_K + 3 * k + 2] = 0;
    if (status_out) status_out[k] = 0;
}

static uint32_t msm_default_window(size_t n) {
    if (n < (1u << 7)) return 4;
    if (n < (1u << 10)) return 7;
    if (n < (1u << 12)) return 8;
    if (n < (1u << 14)) return 10;
    if (n < (1u << 16)) return 11;
    if (n < (1u << 18)) return 12;
    if (n < (1u << 20)) return 13;
    if (n < (1u << 22)) return 14;
    // never 15 or 17: 255 = 15 x 17, so with those widths the signed-digit carry of the top window gets a window of its own, whose
    // single bucket receives half of all the points
    return 16;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <class F>
static int msm_run_t(Ctx *ctx, const Bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp, const uint8_t *d_density,
                     void *d_out_jac, void *d_status_out, int window_bits, MsmBatch batch) {
    cudaStream_t st = ctx->stream;
    if (n_exp >= (1ull << 31) || bases->n >= (1ull << 31)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "n >= 2^31 not supported");
    if (batch.K == 0 || batch.K > 4096) return set_error(ctx, B200ZK_ERR_BAD_ARG, "batch must be in [1, 4096]");
    const uint32_t K = batch.K;
    MsmShape sh;
    const bool use_pre = bases->pre != nullptr && (window_bits <= 0 || (uint32_t)window_bits == bases->pre_c);
    sh.c = use_pre ? bases->pre_c : window_bits > 0 ? (uint32_t)window_bits : msm_default_window(n_exp);
    if (sh.c < 2 || sh.c > 24) return set_error(ctx, B200ZK_ERR_BAD_ARG, "window bits must be in [2, 24]");
    sh.W = (256 + sh.c - 1) / sh.c;
    sh.B = 1u << (sh.c - 1);
    sh.pre_n = use_pre ? (uint32_t)bases->n : 0u;
    sh.sets = use_pre ? 1u : sh.W;
    const uint32_t bw = K * sh.sets;  // bucket sets of the whole batch
    const size_t nbk = (size_t)bw * sh.B;
    const size_t refs_max = (size_t)K * n_exp * sh.W;  // digits of the whole batch
    if (nbk + refs_max >= (1ull << 32)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "batch * n * windows >= 2^32 not supported");
    const void *point_table = use_pre ? bases->pre : bases->points;

    // workspace carve-up
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_status = take((ST_PER_K + 3 * (size_t)K) * sizeof(uint32_t));
    size_t o_counts = take((nbk + 1) * sizeof(uint32_t));
    size_t o_offsets = take((nbk + 1) * sizeof(uint32_t));
    size_t o_cursor = take((nbk + 1) * sizeof(uint32_t));
    size_t o_sums = take((std::max(nbk, n_exp) / SCAN_BLOCK + 2) * sizeof(uint32_t));
    size_t o_rank = take(d_density ? K * (n_exp + 1) * sizeof(uint32_t) : 0);
    size_t o_sorted = take(refs_max * sizeof(uint32_t));
    size_t o_buckets = take(nbk * sizeof(XYZZ<F>));
    // bucket splitting: the host's bound on the cap (the device lowers it from the histogram, k_msm_pick_cap)
    const size_t mean_load = n_exp * (sh.W / sh.sets) / sh.B;
    uint32_t cap = (uint32_t)(mean_load + mean_load / 2 + 32);  // chains longer than ~1.5 x the mean are split
    {
        // A small multiexp cannot fill the machine with one thread per bucket: its time is (longest chain) x (latency of one
        // dependent point addition).  Cut the chains so that there are about `waves` tasks per resident thread slot.
        // Automatic for multiexps with fewer than 64 additions per resident thread (a Spend proof's multiexps, a strong-scaling
        // shard): measured on B200, two waves of tasks: G1 2^16 1.53 -> 1.26 ms, G2 61 300 points 5.9 -> 4.8 ms.  Large
        // multiexps are throughput-bound and keep whole buckets (splitting only adds partial sums to fold).
        const size_t slots = sizeof(F) > 48 ? (size_t)ctx->sm_count * AccPairShape::MINBLOCKS * AccPairShape::THREADS / 2  // G2: a lane pair per chain
                                            : (size_t)ctx->sm_count * AccShape<F>::MINBLOCKS * AccShape<F>::THREADS;
        double waves = refs_max < 64 * slots ? 2.0 : 0.0;
        if (const char *e = getenv("B200ZK_MSM_WAVES")) waves = atof(e);
        if (waves > 0) {
            const size_t fill = (size_t)((double)refs_max / (waves * (double)slots));
            cap = (uint32_t)std::min<size_t>(cap, std::max<size_t>(8, fill));
        }
    }
    // tasks <= 2 T / cap with the device's cap >= 1.25 T / buckets, hence <= 1.6 x buckets whatever the scalars are
    const size_t max_tasks = std::max((size_t)2 * refs_max / cap, 2 * nbk) + 2;
    size_t o_tcnt = take((nbk + 1) * sizeof(uint32_t)), o_toff = take((nbk + 1) * sizeof(uint32_t));
    size_t o_split = take((nbk + 1) * sizeof(uint32_t)), o_partials = take(max_tasks * sizeof(XYZZ<F>));
    size_t o_shist = take((cap + 2) * sizeof(uint32_t)), o_scur = take((cap + 2) * sizeof(uint32_t)), o_order = take((nbk + 1) * sizeof(uint32_t));
    // reduction ladder: C ping-pong (B / 2 and B / 4 entries per set), the odd-entry trees ping-pong (<= B / 4 entries per set each)
    const size_t half_b = std::max<size_t>(sh.B / 2, 1), quarter_b = std::max<size_t>(sh.B / 4, 1);
    uint32_t log_b = 0;
    while ((1u << log_b) < sh.B) log_b++;
    size_t o_c0 = take((size_t)bw * half_b * sizeof(XYZZ<F>)), o_c1 = take((size_t)bw * quarter_b * sizeof(XYZZ<F>));
    size_t o_d0 = take((size_t)bw * quarter_b * sizeof(XYZZ<F>)), o_d1 = take((size_t)bw * quarter_b * sizeof(XYZZ<F>));
    size_t o_fr = take((size_t)bw * sizeof(XYZZ<F>)), o_fa = take((size_t)bw * sizeof(XYZZ<F>));
    int rc = ensure_scratch(ctx, &ctx->scratch2, &ctx->scratch2_bytes, off);
    if (rc) return rc;
    char *ws = (char *)ctx->scratch2;
    uint32_t *status = (uint32_t *)(ws + o_status), *counts = (uint32_t *)(ws + o_counts), *offsets = (uint32_t *)(ws + o_offsets);
    uint32_t *cursor = (uint32_t *)(ws + o_cursor), *sums = (uint32_t *)(ws + o_sums), *rank = d_density ? (uint32_t *)(ws + o_rank) : nullptr;
    uint32_t *sorted = (uint32_t *)(ws + o_sorted);
    XYZZ<F> *buckets = (XYZZ<F> *)(ws + o_buckets);
    uint32_t *task_cnt = (uint32_t *)(ws + o_tcnt), *task_off = (uint32_t *)(ws + o_toff), *split_list = (uint32_t *)(ws + o_split);
    XYZZ<F> *partials = (XYZZ<F> *)(ws + o_partials);
    uint32_t *size_hist = (uint32_t *)(ws + o_shist), *size_cur = (uint32_t *)(ws + o_scur), *order = (uint32_t *)(ws + o_order);
    XYZZ<F> *lc[2] = {(XYZZ<F> *)(ws + o_c0), (XYZZ<F> *)(ws + o_c1)}, *ld[2] = {(XYZZ<F> *)(ws + o_d0), (XYZZ<F> *)(ws + o_d1)};

    B200ZK_CUDA(ctx, cudaMemsetAsync(status + ST_PER_K, 0xff, 3 * (size_t)K * sizeof(uint32_t), st));
    B200ZK_CUDA(ctx, cudaMemsetAsync(status + ST_NONCANONICAL, 0, sizeof(uint32_t), st));
    if (n_exp == 0) {
        k_write_zero_point<F><<<K, 1, 0, st>>>((Jacobian<F> *)d_out_jac, status, (uint32_t *)d_status_out);
        B200ZK_CUDA(ctx, cudaGetLastError());
        return B200ZK_OK;
    }
    B200ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (nbk + 1) * sizeof(uint32_t), st));
    if (d_density)
        for (uint32_t k = 0; k < K; k++)
            ctx->launches += scan_u32<uint8_t>(st, d_density + k * batch.density_stride, n_exp, rank + k * (n_exp + 1), nullptr, sums);
    const dim3 eb((unsigned)((n_exp + 255) / 256), K);
    k_msm_digits<0><<<eb, 256, 0, st>>>((const uint32_t *)d_scalars, n_exp, d_density, rank, base_offset, bases->n, bases->infinity, sh, batch,
                                        counts, nullptr, status, 0u, 0xffffffffu);
    ctx->launches += scan_u32<uint32_t>(st, counts, nbk, offsets, cursor, sums);
    // The scatter writes 4-byte entries at random places of `sorted`; when that array is far larger than the L2 every entry
    // costs a 32-byte read-modify-write in HBM.  Scattering one bucket range at a time (re-deriving the digits, which is
    // cheap) keeps the live part of `sorted` L2-resident, so a bucket's entries merge into full sectors before they leave.
    // Measured on B200 (2^22 ... 2^26 points): four passes save 3 % of the whole multiexp, more passes cost what they save.
    const size_t sorted_bytes = refs_max * sizeof(uint32_t);
    uint32_t passes = !use_pre ? 1u : sorted_bytes >= (256u << 20) ? 4u : sorted_bytes >= (128u << 20) ? 2u : 1u;
    if (const char *e = getenv("B200ZK_SCATTER_PASSES")) passes = (uint32_t)std::max(1, atoi(e));
    passes = std::max(1u, std::min<uint32_t>(passes, (uint32_t)nbk));
    for (uint32_t ps = 0; ps < passes; ps++) {
        const uint32_t lo = (uint32_t)((size_t)nbk * ps / passes), hi = (uint32_t)((size_t)nbk * (ps + 1) / passes);
        k_msm_digits<1><<<eb, 256, 0, st>>>((const uint32_t *)d_scalars, n_exp, d_density, rank, base_offset, bases->n, bases->infinity, sh, batch,
                                            cursor, sorted, status, lo, hi);
    }
    ctx->launches += passes - 1;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (ctx->prof_on) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); cudaEventRecord(pe0, st); }
    {
        uint32_t serial_max = SPLIT_SERIAL_MAX;
        if (const char *e = getenv("B200ZK_SPLIT_SERIAL_MAX")) serial_max = (uint32_t)atoi(e);
        uint32_t *n_split = status + 4;  // two counters: few-partial buckets, many-partial buckets
        B200ZK_CUDA(ctx, cudaMemsetAsync(n_split, 0, 2 * sizeof(uint32_t), st));
        B200ZK_CUDA(ctx, cudaMemsetAsync(size_hist, 0, (cap + 2) * sizeof(uint32_t), st));
        uint32_t *cap_dev = status + 6;
        k_msm_pick_cap<<<1, 1, 0, st>>>(offsets, (uint32_t)nbk, cap, cap_dev);
        k_msm_count_tasks<<<(unsigned)((nbk + 255) / 256), 256, 0, st>>>(offsets, (uint32_t)nbk, cap_dev, task_cnt, split_list, n_split, size_hist, serial_max);
        ctx->launches += scan_u32<uint32_t>(st, task_cnt, nbk, task_off, nullptr, sums);
        ctx->launches += scan_u32<uint32_t>(st, size_hist, cap + 1, size_cur, nullptr, sums);
        ctx->launches += 9;  // digits x2, pick_cap, count_tasks, order_buckets, accumulate, combine_small, combine_big, window_combine
        k_msm_order_buckets<<<(unsigned)((nbk + 255) / 256), 256, 0, st>>>(offsets, (uint32_t)nbk, cap_dev, size_cur, order);
        msm_launch_accumulate<F>(st, point_table, sorted, offsets, nbk, task_cnt, task_off, order, max_tasks, buckets, partials);
        k_msm_combine_small<F><<<1024, 64, 0, st>>>(split_list, n_split, task_cnt, task_off, partials, buckets);
        const size_t big_smem = COMBINE_BIG_THREADS / 2 * sizeof(XYZZ<F>);  // one entry per lane pair: 24 KiB (G1) / 48 KiB (G2) of dynamic shared memory
        bool &opted_in = ctx->combine_smem_opt_in[sizeof(F) > 48 ? 1 : 0];
        if (!opted_in) {
            B200ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_combine_big<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem));
            opted_in = true;
        }
        k_msm_combine_big<F><<<256, COMBINE_BIG_THREADS, big_smem, st>>>(split_list, n_split, (uint32_t)nbk, task_cnt, task_off, partials, buckets);
    }
    if (ctx->prof_on) { cudaEventRecord(pe1, st); ctx->prof_events.emplace_back(pe0, pe1); }
    // reduction ladder (see k_msm_ladder_step): log2(B) - 1 one-addition steps, then the weights
    const XYZZ<F> *cin = buckets, *din = nullptr;
    uint32_t m = sh.B;
    for (uint32_t step = 0; m > 2; step++, m >>= 1) {
        const size_t threads = (size_t)bw * ((m >> 1) + (size_t)(step + 1) * (m >> 2));
        XYZZ<F> *cout = lc[step & 1], *dout = ld[step & 1];
        k_msm_ladder_step<F><<<(unsigned)((2 * threads + 63) / 64), 64, 0, st>>>(cin, din, m, step, bw, cout, dout);
        ctx->launches++;
        cin = cout;
        din = dout;
    }
    XYZZ<F> *inR = (XYZZ<F> *)(ws + o_fr), *inA = (XYZZ<F> *)(ws + o_fa);
    k_msm_ladder_final<F><<<bw, 64, 0, st>>>(cin, din, log_b, bw, inR, inA);
    ctx->launches++;
    k_msm_window_combine<F><<<K, 32, 0, st>>>(inR, inA, sh, (Jacobian<F> *)d_out_jac, status, (uint32_t *)d_status_out);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------------ base precomputation
// pre[w * n + i] = 2^(c w) * P_i as affine points (w < W).  180 GB of HBM buys the W-fold table: every window then adds
// into ONE bucket set -- no per-window bucket reduction, no c doublings per window join (multiexp.rs:223-229 disappears),
// and c can grow until the single bucket set costs as much as a window (fewer mixed adds per point).
template <class F>
__global__ void __launch_bounds__(64) k_bases_precompute(const Affine<F> *__restrict__ points, size_t n, uint32_t c, uint32_t W, Affine<F> *__restrict__ pre) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = points[i];
    pre[i] = p;
    Jacobian<F> j = {p.x, p.y, F::one()};
    F zs[31], pz[31];
    F run = F::one();
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t k = 0; k < c; k++) jacobian_double(j);
        Affine<F> t = {j.x, j.y};
        pre[(size_t)w * n + i] = t;
        zs[w - 1] = j.z;
        run = run * j.z;
        pz[w - 1] = run;
    }
    F inv = run.inverse();  // one inversion per base (Montgomery's trick over its W - 1 multiples)
    for (uint32_t w = W - 1; w >= 1; w--) {
        F zinv = w > 1 ? inv * pz[w - 2] : inv;
        inv = inv * zs[w - 1];
        Affine<F> t = pre[(size_t)w * n + i];
        F zi2 = zinv.sqr();
        t.x = t.x * zi2;
        t.y = t.y * (zi2 * zinv);
        pre[(size_t)w * n + i] = t;
    }
}
template <class F>
static int msm_precompute_t(Ctx *ctx, Bases *bases, uint32_t c) {
    if (c < 2 || c > 24) return set_error(ctx, B200ZK_ERR_BAD_ARG, "window bits must be in [2, 24]");
    uint32_t W = (256 + c - 1) / c;
    if (W > 32) return set_error(ctx, B200ZK_ERR_BAD_ARG, "window too narrow for precomputation (more than 32 windows)");
    if ((size_t)W * bases->n >= (1ull << 31)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "precomputed table would exceed 2^31 points");
    if (bases->pre) { cudaFree(bases->pre); bases->pre = nullptr; bases->pre_c = bases->pre_W = 0; }
    if (bases->n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaMalloc(&bases->pre, (size_t)W * bases->n * sizeof(Affine<F>)));
    k_bases_precompute<F><<<(unsigned)((bases->n + 63) / 64), 64, 0, ctx->stream>>>((const Affine<F> *)bases->points, bases->n, c, W, (Affine<F> *)bases->pre);
    B200ZK_CUDA(ctx, cudaGetLastError());
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bases->pre_c = c;
    bases->pre_W = W;
    return B200ZK_OK;
}

// ------------------------------------------------------------------------------------------------ fixed-base batch mul
// table[j][d-1] = d * 2^(8j) * P  (XYZZ), j < nwin, d in 1..255
template <class F>
__global__ void k_fixed_base_table(const Affine<F> *__restrict__ base, XYZZ<F> *__restrict__ table, uint32_t nwin) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwin) return;
    XYZZ<F> start = XYZZ<F>::from_affine(base[0]);
    for (uint32_t d = 0; d < 8 * j; d++) start.dbl();
    XYZZ<F> cur = start;
    for (uint32_t d = 1; d < 256; d++) {
        table[(size_t)j * 255 + d - 1] = cur;
        cur.add(start);
    }
}
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_mul(const XYZZ<F> *__restrict__ table, const uint32_t *__restrict__ scalars, size_t n,
                                                       uint32_t nwin, Affine<F> *__restrict__ out, uint8_t *__restrict__ out_inf) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> acc = XYZZ<F>::zero();
    for (uint32_t j = 0; j < nwin; j++) {
        uint32_t d = (scalars[8 * i + (j >> 2)] >> (8 * (j & 3))) & 0xff;
        if (d) acc.add(table[(size_t)j * 255 + d - 1]);
    }
    Affine<F> r;
    bool inf = acc.is_zero();
    if (inf) { r.x = F::zero(); r.y = F::one(); }
    else {
        F inv = (acc.zz * acc.zzz).inverse();
        r.x = acc.x * (inv * acc.zzz);
        r.y = acc.y * (inv * acc.zz);
    }
    out[i] = r;
    if (out_inf) out_inf[i] = inf ? 1 : 0;
}
template <class F>
static int msm_fixed_base_t(Ctx *ctx, const void *d_base, const void *d_scalars, size_t n, uint32_t bits, void *d_out, uint8_t *d_out_inf) {
    if (bits == 0 || bits > 256) return set_error(ctx, B200ZK_ERR_BAD_ARG, "scalar_bits must be in [1, 256]");
    uint32_t nwin = (bits + 7) / 8;
    int rc = ensure_scratch(ctx, &ctx->scratch2, &ctx->scratch2_bytes, (size_t)nwin * 255 * sizeof(XYZZ<F>));
    if (rc) return rc;
    XYZZ<F> *table = (XYZZ<F> *)ctx->scratch2;
    k_fixed_base_table<F><<<1, 32, 0, ctx->stream>>>((const Affine<F> *)d_base, table, nwin);
    if (n) k_fixed_base_mul<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(table, (const uint32_t *)d_scalars, n, nwin, (Affine<F> *)d_out, d_out_inf);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}
// ------------------------------------------------------------------------------------------------ small helpers
template <class F>
__global__ void k_into_affine(const Jacobian<F> *__restrict__ in, size_t n, Affine<F> *__restrict__ out, uint8_t *__restrict__ out_inf) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> a;
    bool ok = jacobian_to_affine(in[i], a);
    out[i] = a;
    if (out_inf) out_inf[i] = ok ? 0 : 1;
}
// Sum of n Jacobian points with the reference's add_assign (ec.rs:356-444): strided partials, then a serial fold.
template <class F>
__global__ void __launch_bounds__(128) k_sum_points(const char *__restrict__ in, size_t n, size_t stride, Jacobian<F> *__restrict__ partial, Jacobian<F> *__restrict__ out) {
    Jacobian<F> acc = Jacobian<F>::zero();
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) jacobian_add(acc, *reinterpret_cast<const Jacobian<F> *>(in + i * stride));
    partial[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Jacobian<F> t = Jacobian<F>::zero();
        for (uint32_t k = 0; k < blockDim.x; k++) jacobian_add(t, partial[k]);
        *out = t;
    }
}
template <class F>
static int msm_into_affine_t(Ctx *ctx, const void *d_jac, size_t n, void *d_out_xy, uint8_t *d_out_inf) {
    if (n == 0) return B200ZK_OK;
    k_into_affine<F><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const Jacobian<F> *)d_jac, n, (Affine<F> *)d_out_xy, d_out_inf);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}
template <class F>
static int msm_sum_points_t(Ctx *ctx, const void *d_jac_in, size_t n, void *d_jac_out, size_t stride) {
    int rc = ensure_scratch(ctx, &ctx->scratch2, &ctx->scratch2_bytes, 128 * sizeof(Jacobian<F>));
    if (rc) return rc;
    k_sum_points<F><<<1, 128, 0, ctx->stream>>>((const char *)d_jac_in, n, stride ? stride : sizeof(Jacobian<F>), (Jacobian<F> *)ctx->scratch2, (Jacobian<F> *)d_jac_out);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// one translation unit per group (msm_g1.cu / msm_g2.cu) instantiates these for its coordinate field
#define B200ZK_MSM_INSTANTIATE(SUFFIX, F)                                                                                          \
    int msm_run_##SUFFIX(Ctx *ctx, const Bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp, const uint8_t *d_density, \
                         void *d_out_jac, void *d_status_out, int window_bits, MsmBatch batch) {                                   \
        return msm_run_t<F>(ctx, bases, base_offset, d_scalars, n_exp, d_density, d_out_jac, d_status_out, window_bits, batch);    \
    }                                                                                                                              \
    int msm_fixed_base_##SUFFIX(Ctx *ctx, const void *d_base, const void *d_scalars, size_t n, uint32_t bits, void *d_out, uint8_t *d_out_inf) { \
        return msm_fixed_base_t<F>(ctx, d_base, d_scalars, n, bits, d_out, d_out_inf);                                             \
    }                                                                                                                              \
    int msm_into_affine_##SUFFIX(Ctx *ctx, const void *d_jac, size_t n, void *d_out_xy, uint8_t *d_out_inf) {                       \
        return msm_into_affine_t<F>(ctx, d_jac, n, d_out_xy, d_out_inf);                                                           \
    }                                                                                                                              \
    int msm_sum_points_##SUFFIX(Ctx *ctx, const void *d_jac_in, size_t n, void *d_jac_out, size_t stride) { return msm_sum_points_t<F>(ctx, d_jac_in, n, d_jac_out, stride); } \
    int msm_precompute_##SUFFIX(Ctx *ctx, Bases *bases, uint32_t c) { return msm_precompute_t<F>(ctx, bases, c); }                    \
    int msm_build_table_##SUFFIX(Ctx *ctx, const void *d_base_affine, void *d_table, uint32_t nwin) {                                 \
        k_fixed_base_table<F><<<1, 32, 0, ctx->stream>>>((const Affine<F> *)d_base_affine, (XYZZ<F> *)d_table, nwin);                 \
        B200ZK_CUDA(ctx, cudaGetLastError());                                                                                         \
        return B200ZK_OK;                                                                                                             \
    }

}  // namespace b200zk
