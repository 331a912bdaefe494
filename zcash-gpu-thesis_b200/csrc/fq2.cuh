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
    // out of line with the operands and the result in REGISTERS (by value): references would force both operands and the result
    // through local memory at every call (round 1: a 1552-byte stack frame and 97 M local loads in the G2 bucket accumulation)
    __device__ __forceinline__ friend fq2_t operator*(const fq2_t &a, const fq2_t &b) { return mul_call(a, b); }
    static __device__ __noinline__ fq2_t mul_call(fq2_t a, fq2_t b) {
        // the three Karatsuba products in ONE out-of-line body (one call, three independent carry chains for ptxas to interleave):
        // 2.5 % on the G2 multiexp against three separate calls
        fq_t::Triple t = fq_t::mul3_call(a.c0, b.c0, a.c1, b.c1, a.c1 + a.c0, b.c0 + b.c1);
        return {t.x - t.y, t.z - t.x - t.y};
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

}  // namespace b200zk
