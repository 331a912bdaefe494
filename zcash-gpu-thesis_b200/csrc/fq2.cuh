// Fq2 = Fq[u]/(u^2+1) on the device: the coordinate field of G2.  Mirrors pairing::bls12_381::fq2
// (fq2.rs:84-98 square = 2 Fq mul, :118-132 mul = 3 Fq mul (Karatsuba), add/sub/double/negate componentwise).
// Results are canonical per component, hence bit-identical with the reference whatever formula is used: here the product is two
// two-term sums under one Montgomery reduction each (fp.cuh dot_inline), and the bucket accumulation splits an element over a
// lane pair (fq2h_t below).
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
    // out of line with the operands and the result in REGISTERS (by value): references would force both operands and the result
    // through local memory at every call (round 1: a 1552-byte stack frame and 97 M local loads in the G2 bucket accumulation)
    __device__ __forceinline__ friend fq2_t operator*(const fq2_t &a, const fq2_t &b) { return mul_call(a, b); }
    static __device__ __noinline__ fq2_t mul_call(fq2_t a, fq2_t b) {
        // two two-term sums, each under ONE reduction (fp.cuh dot_inline): a0 b0 + (q - a1) b1  and  a0 b1 + a1 b0 -- 2 x 432 wide
        // multiplies and no additions around them, against 3 x 300 + five additions for the reference's Karatsuba form
        // (fq2.rs:118-132; round 1's mul3_call): G2 2^22 62.3 -> 61.6 ms, the 61 300-point G2 multiexp of a Spend proof 2.58 -> 2.44 ms.
        return {fq_t::muladd2_inline(a.c0, b.c0, a.c1.neg_raw(), b.c1), fq_t::muladd2_inline(a.c0, b.c1, a.c1, b.c0)};
    }
    // fq2.rs:84-98
    __device__ __forceinline__ fq2_t sqr() const { return sqr_call(*this); }
    static __device__ __noinline__ fq2_t sqr_call(fq2_t a) {
        fq_t s = a.c0 + a.c1;
        fq_t d = a.c0 - a.c1;
        fq_t::Pair p = fq_t::mul2_call(a.c0, a.c1, d, s);
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

// ------------------------------------------------------------------------------------------------ one Fq2 element on a lane pair
// fq2h_t is ONE component of an Fq2 element: the even lane of a pair holds c0, the odd lane c1 of the same element, and both
// lanes run the same instruction stream.  Additions are componentwise (no traffic); a product needs the partner's operands
// (2 x 12 warp shuffles) and is, on either lane, a sum of two Fq products under ONE reduction:
//     even: c0 = a0 b0 + (q - a1) b1        odd: c1 = a0 b1 + a1 b0            (2 x 144 + 156 = 444 multiplier instructions per lane;
// the single-thread Karatsuba product is 3 x 300 = 900), a square is one Fq product per lane ((a0 + a1)(a0 - a1) | (a1 + a1) a0).
// An Fq2 point then needs half the registers per thread (the G2 bucket accumulation spilled at 255), and a dependent point
// addition takes half the time.  Same field values, canonical per component, as fq2_t.
// Every function must be called by both lanes of the pair together (the shuffles name exactly these two lanes).
struct fq2h_t {
    fq_t c;
    __device__ __forceinline__ static unsigned role() { return threadIdx.x & 1u; }
    __device__ __forceinline__ static unsigned pair_mask() { return 3u << (threadIdx.x & 30u); }
    __device__ __forceinline__ static fq_t partner(const fq_t &v) {
        fq_t r;
#pragma unroll
        for (int i = 0; i < fq_t::N; i++) r.v[i] = __shfl_xor_sync(pair_mask(), v.v[i], 1);
        return r;
    }
    __device__ __forceinline__ static fq_t pick(bool first, const fq_t &a, const fq_t &b) {
        fq_t r;
#pragma unroll
        for (int i = 0; i < fq_t::N; i++) r.v[i] = first ? a.v[i] : b.v[i];
        return r;
    }
    __device__ __forceinline__ static fq2h_t zero() { return {fq_t::zero()}; }
    __device__ __forceinline__ static fq2h_t one() { return {pick(role() == 0, fq_t::one(), fq_t::zero())}; }
    __device__ __forceinline__ bool is_zero() const {
        const int z = c.is_zero();
        return z & __shfl_xor_sync(pair_mask(), z, 1);
    }
    __device__ __forceinline__ friend fq2h_t operator+(const fq2h_t &a, const fq2h_t &b) { return {a.c + b.c}; }
    __device__ __forceinline__ friend fq2h_t operator-(const fq2h_t &a, const fq2h_t &b) { return {a.c - b.c}; }
    __device__ __forceinline__ fq2h_t dbl() const { return {c.dbl()}; }
    __device__ __forceinline__ fq2h_t neg() const { return {c.neg()}; }
    // The exchange, the operand selection and the two-term product form ONE out-of-line body: a call passes two Fq values in and one
    // out (the inlined form put 24 shuffles + 36 selects + the marshalling of four operands at every one of ~20 call sites: 136 KB of
    // kernel, and ptxas moved ~2000 register copies to the multiplier pipe as IMAD.MOV).
    static __device__ __noinline__ fq_t mul_call(fq_t a, fq_t b) {
        const bool odd = role() != 0;
        const fq_t ap = partner(a), bp = partner(b);
        // X * b_mine + Y * b_partner:  even (X, Y) = (a0, q - a1),  odd (X, Y) = (a0, a1) with b_mine = b1, b_partner = b0
        const fq_t X = pick(odd, ap, a);
        const fq_t Y = pick(odd, a, ap.neg_raw());
        return fq_t::muladd2_inline(X, b, Y, bp);
    }
    static __device__ __noinline__ fq_t sqr_call(fq_t a) {
        const bool odd = role() != 0;
        const fq_t ap = partner(a);
        const fq_t P = a + pick(odd, a, ap);      // even: a0 + a1      odd: a1 + a1
        const fq_t Q = pick(odd, ap, a - ap);     // even: a0 - a1      odd: a0
        return fq_t::mul_inline(P, Q);
    }
    __device__ __forceinline__ friend fq2h_t operator*(const fq2h_t &a, const fq2h_t &b) { return {mul_call(a.c, b.c)}; }
    __device__ __forceinline__ fq2h_t sqr() const { return {sqr_call(c)}; }
};

}  // namespace b200zk
