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
        // a0 b0 and a1 b1 in one out-of-line body (two interleaved carry chains): 2 % on the G2 multiexp
        fq_t::Pair p = fq_t::mul2_call(a.c0, b.c0, a.c1, b.c1);
        const fq_t &aa = p.x, &bb = p.y;
        fq_t o = b.c0 + b.c1;
        fq_t c1 = (a.c1 + a.c0) * o;
        c1 = c1 - aa - bb;
        return {aa - bb, c1};
    }
    // fq2.rs:84-98
    __device__ __noinline__ fq2_t sqr() const {
        fq_t s = c0 + c1;
        fq_t d = c0 - c1;
        fq_t::Pair p = fq_t::mul2_call(c0, c1, d, s);
        return {p.y, p.x.dbl()};  // (c0-c1)(c0+c1) = c0^2 - c1^2 ;  2 c0 c1
    }
    __device__ fq2_t inverse() const {  // fq2.rs:134-153
        fq_t t = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * t, (c1 * t).neg()};
    }
};

}  // namespace b200zk
