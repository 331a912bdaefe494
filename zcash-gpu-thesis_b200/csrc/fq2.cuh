// Fq2 = Fq[u]/(u^2+1) on the device: the coordinate field of G2.  Mirrors pairing::bls12_381::fq2
// (fq2.rs:84-98 square = 2 Fq mul, :118-132 mul = 3 Fq mul (Karatsuba), add/sub/double/negate componentwise).
// Results are canonical per component, hence bit-identical with the reference whatever formula is used.
#pragma once
#include "fp.cuh"

namespace b200zk {

struct fq2_t {
    fq_t c0, c1;
    __device__ __forceinline__ static fq2_t zero() { return {fq_t::zero(), fq_t::zero()}; }
    __device__ __forceinline__ static fq2_t one() { return {fq_t::one(), fq_t::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    __device__ __forceinline__ bool operator==(const fq2_t &b) const { return c0 == b.c0 && c1 == b.c1; }
    __device__ __forceinline__ bool operator!=(const fq2_t &b) const { return !(*this == b); }
    __device__ __forceinline__ friend fq2_t operator+(const fq2_t &a, const fq2_t &b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    __device__ __forceinline__ friend fq2_t operator-(const fq2_t &a, const fq2_t &b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    __device__ __forceinline__ fq2_t dbl() const { return {c0.dbl(), c1.dbl()}; }
    __device__ __forceinline__ fq2_t neg() const { return {c0.neg(), c1.neg()}; }
    // fq2.rs:118-132
    __device__ __noinline__ friend fq2_t operator*(const fq2_t &a, const fq2_t &b) {
#ifdef B200ZK_FQ2_FUSED
        // Karatsuba on unreduced 24-limb products: 3 wide products + 2 Montgomery reductions (744 multiplier instructions)
        // instead of 3 full products (900):  c0 = redc(a0 b0 + q^2 - a1 b1),  c1 = redc((a0+a1)(b0+b1) - a0 b0 - a1 b1)
        constexpr int N = 12;
        uint32_t W0[2 * N], W1[2 * N], W2[2 * N], sa[N], sb[N];
        fq_t::mul_rows<N>(W0, a.c0.v, b.c0.v);
        fq_t::mul_rows<N>(W1, a.c1.v, b.c1.v);
        sa[0] = add_cc(a.c0.v[0], a.c1.v[0]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) sa[k] = addc_cc(a.c0.v[k], a.c1.v[k]);
        sa[N - 1] = addc(a.c0.v[N - 1], a.c1.v[N - 1]);  // < 2q < 2^382: no carry out, no reduction needed
        sb[0] = add_cc(b.c0.v[0], b.c1.v[0]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) sb[k] = addc_cc(b.c0.v[k], b.c1.v[k]);
        sb[N - 1] = addc(b.c0.v[N - 1], b.c1.v[N - 1]);
        fq_t::mul_rows<N>(W2, sa, sb);  // < 4 q^2 < 2^764
        W2[0] = sub_cc(W2[0], W0[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) W2[k] = subc_cc(W2[k], W0[k]);
        W2[2 * N - 1] = subc(W2[2 * N - 1], W0[2 * N - 1]);
        W2[0] = sub_cc(W2[0], W1[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) W2[k] = subc_cc(W2[k], W1[k]);
        W2[2 * N - 1] = subc(W2[2 * N - 1], W1[2 * N - 1]);  // a0 b1 + a1 b0 < 2 q^2
        W1[0] = sub_cc(FqParams::mod_sq(0), W1[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) W1[k] = subc_cc(FqParams::mod_sq(k), W1[k]);
        W1[2 * N - 1] = subc(FqParams::mod_sq(2 * N - 1), W1[2 * N - 1]);
        W0[0] = add_cc(W0[0], W1[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) W0[k] = addc_cc(W0[k], W1[k]);
        W0[2 * N - 1] = addc(W0[2 * N - 1], W1[2 * N - 1]);  // a0 b0 - a1 b1 + q^2 in (0, 2 q^2)
        return {fq_t::redc_wide(W0), fq_t::redc_wide(W2)};
#endif
        // the three Karatsuba products in ONE out-of-line body (one call, three independent carry chains for ptxas to interleave):
        // 2.5 % on the G2 multiexp against three separate calls
        fq_t::Triple t = fq_t::mul3_call(a.c0, b.c0, a.c1, b.c1, a.c1 + a.c0, b.c0 + b.c1);
        return {t.x - t.y, t.z - t.x - t.y};
    }
    // fq2.rs:84-98
    __device__ __noinline__ fq2_t sqr() const {
        fq_t s = c0 + c1;
        fq_t d = c0 - c1;
        fq_t::Pair p = fq_t::mul2_call(c0, c1, d, s);
        return {p.y, p.x.dbl()};  // (c0-c1)(c0+c1) = c0^2 - c1^2 ;  2 c0 c1
    }
    __device__ fq2_t inverse_binary() const {  // as inverse(), with the serial Fq inversion (single-thread kernels)
        fq_t t = (c0.sqr() + c1.sqr()).inverse_binary();
        return {c0 * t, (c1 * t).neg()};
    }
    __device__ fq2_t inverse() const {  // fq2.rs:134-153
        fq_t t = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * t, (c1 * t).neg()};
    }
};

}  // namespace b200zk
