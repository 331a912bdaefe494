// Montgomery prime-field arithmetic for BLS12-381 Fr (8 x u32) and Fq (12 x u32) on sm_100a.
//
// Device counterpart of the reference's pairing::bls12_381::{fr,fq} (fr.rs:341-571, fq.rs:813-1123) and of
// its limb primitives adc/sbb/mac_with_carry (pairing/src/lib.rs:645-679).  The reference works on 64-bit
// limbs with u128 products; the B200 integer pipe is 32-bit (IMAD / IMAD.WIDE on the FMA pipe, IADD3 on
// the ALU pipe), so elements are kept as N little-endian u32 limbs -- the *same bytes* as the reference's
// [u64; N/2] -- and every operation returns the fully reduced canonical value in [0, p), exactly like the
// reference (fq.rs:1023-1031 `reduce`), which is what makes results bit-identical.
//
// Multiplication is an operand-scanning Montgomery product with the partial products split into an
// "even" and an "odd" accumulator so that each a_j*b_i (64-bit) lands on a 64-bit aligned pair and the
// whole row is one mad.lo.cc/madc.hi.cc carry chain (ptxas fuses each pair into IMAD.WIDE.U32 + carry).
// The per-row right shift by 32 bits is free: the accumulators swap roles every row.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200zk {

// -p^-1 mod 2^32 for Fr and Fq, read at run time (see mad_n_redc)
__device__ __constant__ uint32_t B200ZK_M0_RT[2] = {0xffffffffu, 0xfffcfffdu};

// ------------------------------------------------------------------------------------------------ PTX carry helpers
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// ------------------------------------------------------------------------------------------------ field parameters
// Constants as in fr.rs:4-55 / fq.rs:5-42, split into u32 limbs.
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = 0xffffffffu;  // -r^-1 mod 2^32 (low word of INV = 0xfffffffeffffffff)
    __device__ __host__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    __device__ __host__ static constexpr uint32_t one(int i) {  // R = 2^256 mod r
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    __device__ __host__ static constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
};
struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = 0xfffcfffdu;  // low word of INV = 0x89f3fffcfffcfffd
    __device__ __host__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    __device__ __host__ static constexpr uint32_t one(int i) {  // R = 2^384 mod q
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    __device__ __host__ static constexpr uint32_t r2(int i) {  // R^2 mod q
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
};

template <class P>
struct __align__(16) Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    __device__ __forceinline__ static Fp zero() { Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r; }
    __device__ __forceinline__ static Fp one() { Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r; }
    __device__ __forceinline__ static Fp r2() { Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
        return r; }
    __device__ __forceinline__ bool is_zero() const { uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i];
        return o == 0; }
    __device__ __forceinline__ bool operator==(const Fp &b) const { uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
        return o == 0; }
    __device__ __forceinline__ bool operator!=(const Fp &b) const { return !(*this == b); }

    // r = x - p if x >= p else x   (x < 2p)      -- fq.rs:1023-1031
    __device__ __forceinline__ static void final_sub(uint32_t *x) {
        uint32_t t[N];
        t[0] = sub_cc(x[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(x[i], P::mod(i));
        uint32_t borrow = subc(0, 0);  // 0xffffffff if x < p
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = borrow ? x[i] : t[i];
    }

    // fq.rs:813-823
    __device__ __forceinline__ friend Fp operator+(const Fp &a, const Fp &b) {
        Fp r;
        r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
        r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);
        final_sub(r.v);
        return r;
    }
    // fq.rs:825-834
    __device__ __forceinline__ friend Fp operator-(const Fp &a, const Fp &b) {
        Fp r;
        r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = subc(0, 0);
        // add p back when the subtraction wrapped
        r.v[0] = add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], P::mod(i) & borrow);
        r.v[N - 1] = addc(r.v[N - 1], P::mod(N - 1) & borrow);
        return r;
    }
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }  // fq.rs:836-842
    __device__ __forceinline__ Fp neg() const {                          // fq.rs:1004-1010
        Fp r;
        r.v[0] = sub_cc(P::mod(0), v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::mod(i), v[i]);
        r.v[N - 1] = subc(P::mod(N - 1), v[N - 1]);
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < N; i++) nz |= v[i];
        uint32_t mask = nz ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] &= mask;
        return r;
    }

    // ---- Montgomery product rows ---------------------------------------------------------------
    // acc[j], acc[j+1] = lo/hi(a[j] * bi), j = 0, 2, ..  (a points at the even- or odd-indexed limbs)
    __device__ __forceinline__ static void mul_n(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N; j += 2) { acc[j] = mul_lo(a[j], bi); acc[j + 1] = mul_hi(a[j], bi); }
    }
    // acc[j..j+1] += a[j] * bi, one carry chain; carry-out left in CC
    __device__ __forceinline__ static void cmad_n(uint32_t *acc, const uint32_t *a, uint32_t bi) {
        acc[0] = mad_lo_cc(a[0], bi, acc[0]);
        acc[1] = madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = madc_lo_cc(a[j], bi, acc[j]); acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]); }
    }
    // same with the modulus (immediates); `odd` selects limbs 1,3,5.. of p
    // t0 = the limb m was derived from (m = -t0 mod 2^32 for Fr, whose M0 is 2^32 - 1)
    template <int ODD>
    __device__ __forceinline__ static void cmad_mod(uint32_t *acc, uint32_t mi, uint32_t t0) {
        if (P::mod(0) == 1u && P::mod(1) == 0xffffffffu) {
            // Fr: the two low limbs of r are 1 and 2^32 - 1, so their products with m need no multiplier:
            //   m * 1 = m ;  m * (2^32 - 1) = (m - [m != 0]) * 2^32 + (2^32 - m)   (hi : lo),  and 2^32 - m = t0 (no negation)
            // c = [m != 0] = [t0 != 0] = min(t0, 1).  Even chain: t0 + m = c 2^32 exactly, so limb 0 (dropped by the shift) needs no
            // addition and c goes straight into limb 1.
            uint32_t c;
            asm("min.u32 %0, %1, 1;" : "=r"(c) : "r"(t0));
            if (ODD) {
                acc[0] = add_cc(acc[0], t0);
                acc[1] = addc_cc(acc[1], mi - c);
            } else {
                acc[0] = 0;
                acc[1] = add_cc(acc[1], c);
            }
        } else {
            acc[0] = mad_lo_cc(P::mod(ODD), mi, acc[0]);
            acc[1] = madc_hi_cc(P::mod(ODD), mi, acc[1]);
        }
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = madc_lo_cc(P::mod(j + ODD), mi, acc[j]); acc[j + 1] = madc_hi_cc(P::mod(j + ODD), mi, acc[j + 1]); }
    }
    // acc[j] = a[j]*bi + acc[j+2] (+ carry-in), i.e. accumulate while shifting acc down by two limbs (64 bits)
    __device__ __forceinline__ static void madc_n_rshift(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) { acc[j] = madc_lo_cc(a[j], bi, acc[j + 2]); acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 3]); }
        acc[N - 2] = madc_lo_cc(a[N - 2], bi, 0);
        acc[N - 1] = madc_hi(a[N - 2], bi, 0);
    }
    // m = t0 * (-p^-1) mod 2^32.  Fq: M0 comes from constant memory (as an immediate ptxas strength-reduces the m * p_k products
    // into slower forms).  Fr: M0 = 2^32 - 1, so m = -t0 -- an opaque ALU subtraction instead of a product on the multiplier pipe
    // (the wide products keep their IMAD.WIDE.X form: 112 per Fr product in the SASS).
    __device__ __forceinline__ static uint32_t mont_m(uint32_t t0) {
        uint32_t mi;
        if (P::N == 8) asm volatile("sub.u32 %0, 0, %1;" : "=r"(mi) : "r"(t0));
        else mi = t0 * B200ZK_M0_RT[1];
        return mi;
    }
    // One row: T += a*bi; T += m*p; T >>= 32 (the shift is implicit: the caller swaps `even` and `odd`).
    // Value represented: T = sum even[k] 2^(32k) + 2^32 * sum odd[k] 2^(32k).
    // Written as mad.lo.cc / madc.hi.cc pairs.  Inside the out-of-line product (mul_call) ptxas fuses every pair into
    // one IMAD.WIDE.U32.X with a predicate carry: ~300 multiplier-pipe instructions + ~80 ALU per Fq product, the
    // minimum for 32-bit limbs.  (Measured on B200: IMAD.WIDE/.HI issue at 32/clk/SM on the fmaheavy pipe, plain IMAD
    // at 64/clk/SM; when the same code is inlined into a large kernel ptxas sometimes splits pairs into IMAD +
    // IMAD.HI + 2 IADD3.X, which is why products go through the call by default.)
    __device__ __forceinline__ static void mad_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi, bool first) {
        if (first) {
            mul_n(odd, a + 1, bi);
            mul_n(even, a, bi);
        } else {
            even[0] = add_cc(even[0], odd[1]);   // stray limb of the previous row
            madc_n_rshift(odd, a + 1, bi);       // absorbs that carry (odd[0] is one limb above even[0])
            cmad_n(even, a, bi);
            odd[N - 1] = addc(odd[N - 1], 0);
        }
        // M0 comes from constant memory: for Fr it is 2^32 - 1 and ptxas would otherwise rewrite m = -t0 and the
        // m * p_k products into negated IMAD.X / IMAD.HI pairs (6 cycles) instead of one IMAD.WIDE.X (4 cycles)
        const uint32_t t0 = even[0];
        uint32_t mi = mont_m(t0);
        cmad_mod<1>(odd, mi, t0);
        cmad_mod<0>(even, mi, t0);
        odd[N - 1] = addc(odd[N - 1], 0);
    }

    // ---- plain M x M limb product by rows (M even): the row scheme of mad_n_redc without the reduction; row i finishes limb i
    template <int M>
    __device__ __forceinline__ static void mulw_row(uint32_t *base, uint32_t *up, const uint32_t *a, uint32_t bi) {
        base[0] = add_cc(base[0], up[1]);  // stray limb of the previous row
#pragma unroll
        for (int j = 0; j < M - 2; j += 2) { up[j] = madc_lo_cc(a[j + 1], bi, up[j + 2]); up[j + 1] = madc_hi_cc(a[j + 1], bi, up[j + 3]); }
        up[M - 2] = madc_lo_cc(a[M - 1], bi, 0);
        up[M - 1] = madc_hi(a[M - 1], bi, 0);
        base[0] = mad_lo_cc(a[0], bi, base[0]);
        base[1] = madc_hi_cc(a[0], bi, base[1]);
#pragma unroll
        for (int j = 2; j < M; j += 2) { base[j] = madc_lo_cc(a[j], bi, base[j]); base[j + 1] = madc_hi_cc(a[j], bi, base[j + 1]); }
        up[M - 1] = addc(up[M - 1], 0);
    }
    template <int M>
    __device__ __forceinline__ static void mul_rows(uint32_t *T, const uint32_t *a, const uint32_t *b) {
        uint32_t even[M], odd[M];
#pragma unroll
        for (int j = 0; j < M; j += 2) {
            odd[j] = mul_lo(a[j + 1], b[0]); odd[j + 1] = mul_hi(a[j + 1], b[0]);
            even[j] = mul_lo(a[j], b[0]); even[j + 1] = mul_hi(a[j], b[0]);
        }
        T[0] = even[0];
#pragma unroll
        for (int i = 1; i < M; i += 2) {
            mulw_row<M>(odd, even, a, b[i]);
            T[i] = odd[0];
            if (i + 1 < M) {
                mulw_row<M>(even, odd, a, b[i + 1]);
                T[i + 1] = even[0];
            }
        }
        // the last row (i = M - 1, odd) left `odd` as the base array and `even` one limb above it
        T[M] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 1; k < M - 1; k++) T[M + k] = addc_cc(even[k], odd[k + 1]);
        T[2 * M - 1] = addc(even[M - 1], 0);
    }
    // fq.rs:910-963 mul_assign + fq.rs:1040-1123 mont_reduce  ->  a*b*R^-1 mod p, canonical
    __device__ __forceinline__ friend Fp operator*(const Fp &a, const Fp &b) {
        return mul_call(a, b);
    }
    // Out-of-line product: a fully inlined mixed add is ~100 KB of straight-line SASS and ncu showed
    // `no_instruction` (instruction-cache misses) as the top stall of the bucket-accumulation kernel.  One shared
    // ~10 KB function body keeps the hot loop inside the instruction cache; arguments and the result travel in
    // registers (no stack traffic).
    static __device__ __noinline__ Fp mul_call(Fp a, Fp b) { return mul_inline(a, b); }
    // a*b - c*d with ONE Montgomery reduction (the Y3 of the point additions is such a difference: 156 multiplier instructions
    // less than two products): a*b + (p - c)*d through the fused two-term rows below.  (Round 1 formed both products as unreduced
    // 2N-limb values, T = a*b + (p^2 - c*d), and reduced T: the same multiplies, ~100 more additions and two 24-limb temporaries.)
    // Only for Fq (see dot_inline).
    static __device__ __noinline__ Fp mulsub_call(Fp a, Fp b, Fp c, Fp d) { return muladd2_inline(a, b, c.neg_raw(), d); }
    // a*b + c*d with ONE Montgomery reduction, for operands up to p (not only below it): T <= 2 p^2 < p R.  The lane-split Fq2
    // product (fq2.cuh fq2h_t) is such a sum on either lane -- a0 b0 + (p - a1) b1 and a0 b1 + a1 b0.
    static __device__ __noinline__ Fp muladd2_call(Fp a, Fp b, Fp c, Fp d) { return muladd2_inline(a, b, c, d); }
    // Fused like the single product (mad_n_redc): every row adds x_k * y_k[i] for ALL K terms to the even / odd accumulators, then
    // m p, then shifts -- no 2N-limb intermediates, no separate additions.  A row leaves T < x_0 + .. + x_(K-1) + p <= (K + 1) p,
    // and 2^32 (K + 1) p fits the accumulator pair (N + 1 limbs) while (K + 1) p < 2^(32 N): true for Fq up to K = 8
    // (q / 2^384 = 0.10); the result is < p (1 + K p / R) < 2p, one conditional subtraction away from canonical.
    template <int K>
    __device__ __forceinline__ static void madk_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *const (&x)[K], const uint32_t (&yi)[K], bool first) {
        if (first) {
            mul_n(odd, x[0] + 1, yi[0]);
            mul_n(even, x[0], yi[0]);
        } else {
            even[0] = add_cc(even[0], odd[1]);
            madc_n_rshift(odd, x[0] + 1, yi[0]);
            cmad_n(even, x[0], yi[0]);
            odd[N - 1] = addc(odd[N - 1], 0);
        }
#pragma unroll
        for (int k = 1; k < K; k++) {
            cmad_n(odd, x[k] + 1, yi[k]);  // (its carry out of the top limb is zero: the whole sum fits)
            cmad_n(even, x[k], yi[k]);
            odd[N - 1] = addc(odd[N - 1], 0);
        }
        const uint32_t t0 = even[0];
        uint32_t mi = mont_m(t0);
        cmad_mod<1>(odd, mi, t0);
        cmad_mod<0>(even, mi, t0);
        odd[N - 1] = addc(odd[N - 1], 0);
    }
    // sum_k x_k * y_k under one reduction; operands up to p (not only below it)
    template <int K>
    __device__ __forceinline__ static Fp dot_inline(const uint32_t *const (&x)[K], const uint32_t *const (&y)[K]) {
        static_assert(N == 12 && K <= 8, "the accumulator bound above is checked for Fq only");
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            uint32_t y0[K], y1[K];
#pragma unroll
            for (int k = 0; k < K; k++) { y0[k] = y[k][i]; y1[k] = y[k][i + 1]; }
            madk_n_redc<K>(even, odd, x, y0, i == 0);
            madk_n_redc<K>(odd, even, x, y1, false);
        }
        Fp r;
        r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(even[i], odd[i + 1]);
        r.v[N - 1] = addc(even[N - 1], 0);
        final_sub(r.v);
        return r;
    }
    __device__ __forceinline__ static Fp muladd2_inline(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
        const uint32_t *const x[2] = {a.v, c.v};
        const uint32_t *const y[2] = {b.v, d.v};
        return dot_inline<2>(x, y);
    }
    // p - a without the zero test: a value in (0, p] that stands for -a as an operand of muladd2_call
    __device__ __forceinline__ Fp neg_raw() const {
        Fp r;
        r.v[0] = sub_cc(P::mod(0), v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::mod(i), v[i]);
        r.v[N - 1] = subc(P::mod(N - 1), v[N - 1]);
        return r;
    }
    // Two independent products in one out-of-line body (the Fq2 squaring: one call instead of two)
    struct Pair { Fp x, y; };
    static __device__ __noinline__ Pair mul2_call(Fp a, Fp b, Fp c, Fp d) { return {mul_inline(a, b), mul_inline(c, d)}; }
    __device__ __forceinline__ static Fp mul_inline(const Fp &a, const Fp &b) { return mul_inline_t<true>(a, b); }
    // CANONICAL = false: no final subtraction.  For a < p, b < 2p the result is < p (2p / R + 1) < 2p when 2p^2 < pR, true for Fr
    // (r / 2^256 = 0.453): the NTT keeps its values in [0, 2r) between stages and pays one subtraction at the very end.
    // The operand that may exceed p must be b, the one scanned limb by limb: a row leaves T < a + p, and the even/odd
    // accumulator pair holds 2^32 (a + p) only while a + p < 2^(32N).
    template <bool CANONICAL>
    __device__ __forceinline__ static Fp mul_inline_t(const Fp &a, const Fp &b) {
        uint32_t even[N], odd[N];
        // a's limbs interleave: even-indexed at a.v[0], a.v[2].. ; mul_n/cmad_n step by 2 from the pointer given
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_n_redc(even, odd, a.v, b.v[i], i == 0);
            mad_n_redc(odd, even, a.v, b.v[i + 1], false);
        }
        // merge: result = even + (odd >> 32) with odd[0] == 0 ... (here roles are back to the original names)
        Fp r;
        r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(even[i], odd[i + 1]);
        r.v[N - 1] = addc(even[N - 1], 0);
        if (CANONICAL) final_sub(r.v);
        return r;
    }
    // ---- arithmetic on values in [0, 2p) (2p < 2^(32N) for both fields)
    __device__ __host__ static constexpr uint32_t mod2(int i) { return (P::mod(i) << 1) | (i ? P::mod(i - 1) >> 31 : 0u); }
    // (a + b) mod' 2p: a, b < 2p -> result < 2p.  The sum can pass 2^(32N): the carry joins the comparison.
    __device__ __forceinline__ static Fp add_2p(const Fp &a, const Fp &b) {
        Fp s;
        uint32_t t[N];
        s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
        const uint32_t carry = addc(0, 0);
        t[0] = sub_cc(s.v[0], mod2(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(s.v[i], mod2(i));
        const uint32_t keep = subc(carry, 0) >> 31;  // 1 iff carry:s < 2p
#pragma unroll
        for (int i = 0; i < N; i++) s.v[i] = keep ? s.v[i] : t[i];
        return s;
    }
    // (a - b) mod' 2p: a, b < 2p -> result < 2p
    __device__ __forceinline__ static Fp sub_2p(const Fp &a, const Fp &b) {
        Fp r;
        r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
        const uint32_t borrow = subc(0, 0);
        r.v[0] = add_cc(r.v[0], mod2(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], mod2(i) & borrow);
        r.v[N - 1] = addc(r.v[N - 1], mod2(N - 1) & borrow);
        return r;
    }
    // [0, 2p) -> [0, p)
    __device__ __forceinline__ Fp canonical() const { Fp r = *this; final_sub(r.v); return r; }
    // fq.rs:965-1002 square: the off-diagonal products a_i a_j (i < j) are formed once and doubled, 78 instead of 144
    // multiplier instructions for the product; then the same Montgomery rows as the product, without the a*b part.
    __device__ __forceinline__ Fp sqr() const {
        return sqr_call(*this);
    }
    static __device__ __noinline__ Fp sqr_call(Fp a) { return sqr_inline(a); }
    __device__ __forceinline__ static Fp sqr_inline(const Fp &a) {
        // E: products whose position i + j is even (64-bit aligned at even limbs); O: odd positions, O[k] sits at limb k + 1
        uint32_t E[2 * N], O[2 * N];
#pragma unroll
        for (int k = 0; k < 2 * N; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            if (i + 2 < N) {
#pragma unroll
                for (int j = i + 2; j < N; j += 2) {
                    E[i + j] = j == i + 2 ? mad_lo_cc(a.v[i], a.v[j], E[i + j]) : madc_lo_cc(a.v[i], a.v[j], E[i + j]);
                    E[i + j + 1] = madc_hi_cc(a.v[i], a.v[j], E[i + j + 1]);
                }
                // The chain ends in limb `last`.  Rows below reached at most limb `last` (and only as a high half), so the
                // carry can spill one limb further, into a limb that is still zero: one addition, not a ripple to the top.
                const int last = i + (i + 2 + ((N - 1 - (i + 2)) / 2) * 2) + 1;
                if (last + 1 <= 2 * N - 1) E[last + 1] = addc(E[last + 1], 0);
            }
            {
#pragma unroll
                for (int j = i + 1; j < N; j += 2) {
                    O[i + j - 1] = j == i + 1 ? mad_lo_cc(a.v[i], a.v[j], O[i + j - 1]) : madc_lo_cc(a.v[i], a.v[j], O[i + j - 1]);
                    O[i + j] = madc_hi_cc(a.v[i], a.v[j], O[i + j]);
                }
                const int last = i + (i + 1 + ((N - 1 - (i + 1)) / 2) * 2);
                if (last + 1 <= 2 * N - 1) O[last + 1] = addc(O[last + 1], 0);
            }
        }
        // U = E + (O << 32), T = 2 U + sum a_i^2 2^(64 i)
        uint32_t T[2 * N];
        T[0] = E[0];
        T[1] = add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) T[k] = addc_cc(E[k], O[k - 1]);
        T[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2]);
#pragma unroll
        for (int k = 2 * N - 1; k > 0; k--) T[k] = __funnelshift_l(T[k - 1], T[k], 1);
        T[0] <<= 1;
        T[0] = mad_lo_cc(a.v[0], a.v[0], T[0]);
        T[1] = madc_hi_cc(a.v[0], a.v[0], T[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) { T[2 * i] = madc_lo_cc(a.v[i], a.v[i], T[2 * i]); T[2 * i + 1] = madc_hi_cc(a.v[i], a.v[i], T[2 * i + 1]); }
        T[2 * N - 2] = madc_lo_cc(a.v[N - 1], a.v[N - 1], T[2 * N - 2]);
        T[2 * N - 1] = madc_hi(a.v[N - 1], a.v[N - 1], T[2 * N - 1]);
        return redc_wide(T);
    }
    // Montgomery reduction of an unreduced 2N-limb value T < 2 p^2: the rows of mad_n_redc without the a*b part on the low
    // half, then + high half; one conditional subtraction gives the canonical result
    __device__ __forceinline__ static Fp redc_wide(const uint32_t *T) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int k = 0; k < N; k++) { even[k] = T[k]; odd[k] = 0; }
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            redc_row(even, odd, i == 0);
            redc_row(odd, even, false);
        }
        Fp r;
        r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(even[i], odd[i + 1]);
        r.v[N - 1] = addc(even[N - 1], 0);
        r.v[0] = add_cc(r.v[0], T[N]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], T[N + i]);
        r.v[N - 1] = addc(r.v[N - 1], T[2 * N - 1]);
        final_sub(r.v);
        return r;
    }
    // one reduction row: T += m p; T >>= 32 (roles of even / odd swap at the caller)
    __device__ __forceinline__ static void redc_row(uint32_t *even, uint32_t *odd, bool first) {
        if (!first && !(P::mod(0) == 1u && P::mod(1) == 0xffffffffu)) {
            // the two-limb shift of `odd` rides in the addend of its m * p products (as madc_n_rshift does for a * b):
            // 12 carry additions less per row than shifting first and multiplying after
            even[0] = add_cc(even[0], odd[1]);
            const uint32_t mi = even[0] * B200ZK_M0_RT[P::N == 8 ? 0 : 1];  // a plain product: the carry flag survives it
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) { odd[j] = madc_lo_cc(P::mod(j + 1), mi, odd[j + 2]); odd[j + 1] = madc_hi_cc(P::mod(j + 1), mi, odd[j + 3]); }
            odd[N - 2] = madc_lo_cc(P::mod(N - 1), mi, 0);
            odd[N - 1] = madc_hi(P::mod(N - 1), mi, 0);
            cmad_mod<0>(even, mi, even[0]);
            odd[N - 1] = addc(odd[N - 1], 0);
            return;
        }
        if (!first) {
            even[0] = add_cc(even[0], odd[1]);  // stray limb; the carry is absorbed while odd shifts down by two limbs
#pragma unroll
            for (int j = 0; j < N - 2; j++) odd[j] = addc_cc(odd[j + 2], 0);
            odd[N - 2] = addc(0, 0);
            odd[N - 1] = 0;
        }
        const uint32_t t0 = even[0];
        uint32_t mi = mont_m(t0);
        cmad_mod<1>(odd, mi, t0);
        cmad_mod<0>(even, mi, t0);
        odd[N - 1] = addc(odd[N - 1], 0);
    }

    // Montgomery -> canonical (fr.rs:290-303 into_repr) and back (fr.rs:279-288 from_repr)
    __device__ __forceinline__ Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }
    __device__ __forceinline__ Fp to_mont() const { return *this * r2(); }

    // Field::pow (lib.rs:306-324) with a small exponent
    __device__ Fp pow(uint64_t e) const {
        Fp res = one();
        bool found = false;
        for (int i = 63; i >= 0; i--) {
            bool bit = (e >> i) & 1;
            if (found) res = res.sqr(); else found = bit;
            if (bit) res = res * *this;
        }
        return res;
    }
    // Inverse by the binary extended Euclid the reference uses (fq.rs:849-903; b starts at R^2 so the result is already in
    // Montgomery form).  Data-dependent loops, ~5x fewer cycles than the Fermat ladder for ONE thread: for the single-thread
    // tails (proof assembly); full warps keep `inverse()`, whose control flow is uniform.  Same canonical value.  0 -> 0.
    __device__ __noinline__ Fp inverse_binary() const {
        if (is_zero()) return zero();
        Fp u = *this, w, b = r2(), c = zero();
#pragma unroll
        for (int i = 0; i < N; i++) w.v[i] = P::mod(i);
        auto is_one = [](const Fp &x) {
            uint32_t rest = 0;
            for (int i = 1; i < N; i++) rest |= x.v[i];
            return rest == 0 && x.v[0] == 1u;
        };
        auto halve = [](Fp &x) {
            for (int i = 0; i < N - 1; i++) x.v[i] = __funnelshift_r(x.v[i], x.v[i + 1], 1);
            x.v[N - 1] >>= 1;
        };
        auto halve_mod = [&](Fp &x) {  // x / 2 mod p for x in [0, p): (x + p) / 2 when x is odd (x + p < 2^(32 N), no carry out)
            if (x.v[0] & 1u) {
                x.v[0] = add_cc(x.v[0], P::mod(0));
#pragma unroll
                for (int i = 1; i < N - 1; i++) x.v[i] = addc_cc(x.v[i], P::mod(i));
                x.v[N - 1] = addc(x.v[N - 1], P::mod(N - 1));
            }
            halve(x);
        };
        auto less = [](const Fp &x, const Fp &y) {
            for (int i = N - 1; i >= 0; i--)
                if (x.v[i] != y.v[i]) return x.v[i] < y.v[i];
            return false;
        };
        auto sub_raw = [](Fp &x, const Fp &y) {  // x -= y, x >= y
            x.v[0] = sub_cc(x.v[0], y.v[0]);
#pragma unroll
            for (int i = 1; i < N - 1; i++) x.v[i] = subc_cc(x.v[i], y.v[i]);
            x.v[N - 1] = subc(x.v[N - 1], y.v[N - 1]);
        };
        while (!is_one(u) && !is_one(w)) {
            while (!(u.v[0] & 1u)) { halve(u); halve_mod(b); }
            while (!(w.v[0] & 1u)) { halve(w); halve_mod(c); }
            if (less(w, u)) { sub_raw(u, w); b = b - c; } else { sub_raw(w, u); c = c - b; }
        }
        return is_one(u) ? b : c;
    }
    // inverse by Fermat (p-2).  Canonical, so it equals the reference's binary EEA (fq.rs:849-903).  0 -> 0.
    __device__ Fp inverse() const {
        uint32_t e[N];
        e[0] = sub_cc(P::mod(0), 2);  // p - 2 (Fr's low word is 1: the borrow must ripple)
#pragma unroll
        for (int i = 1; i < N - 1; i++) e[i] = subc_cc(P::mod(i), 0);
        e[N - 1] = subc(P::mod(N - 1), 0);
        Fp res = one();
        bool found = false;
        for (int i = 32 * N - 1; i >= 0; i--) {
            bool bit = (e[i >> 5] >> (i & 31)) & 1;
            if (found) res = res.sqr(); else found = bit;
            if (bit) res = res * *this;
        }
        return res;
    }
};

typedef Fp<FrParams> fr_t;
typedef Fp<FqParams> fq_t;

}  // namespace b200zk
