// BLS12-381 short-Weierstrass group law on the device, generic over the coordinate field F (fq_t -> G1,
// fq2_t -> G2).  Replaces the reference's CurveProjective::{double, add_assign, add_assign_mixed, into_affine}
// (pairing/src/bls12_381/ec.rs:296-354, 356-444, 446-526, 586-619).
//
// The reference accumulates in Jacobian coordinates (madd-2007-bl, 7M+4S).  Buckets here use the extended
// XYZZ form (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; madd-2008-s, 8M+2S) which saves a field multiplication per
// mixed add and has no 2*Y1*J style doublings.  Group results are identical; only the projective
// representative differs, and parity is defined on the affine value (the reference's own PartialEq
// cross-multiplies, ec.rs:45-85).  The C ABI hands back a Jacobian triple (X*ZZ, Y*ZZZ, ZZ).
#pragma once
#include "fp.cuh"
#include "fq2.cuh"

namespace b200zk {

template <class F>
struct Affine {  // x, y in Montgomery form; the reference keeps `infinity: bool` beside them (ec.rs:13-18)
    F x, y;
};

template <class F>
struct Jacobian {  // ec.rs:20-24; identity <=> z == 0, canonical zero = (0, 1, 0) (ec.rs:224-230)
    F x, y, z;
    __device__ __forceinline__ static Jacobian zero() { return {F::zero(), F::one(), F::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return z.is_zero(); }
};

// a*b - c*d: for Fq one fused out-of-line body with a single reduction (fp.cuh mulsub_call), otherwise two products
template <class F>
__device__ __forceinline__ F mul_sub(const F &a, const F &b, const F &c, const F &d) { return a * b - c * d; }
#ifndef B200ZK_NO_FUSED_MULSUB  // measured: G1 MSM 2^24 81.1 -> 79.7 ms
template <>
__device__ __forceinline__ fq_t mul_sub<fq_t>(const fq_t &a, const fq_t &b, const fq_t &c, const fq_t &d) { return fq_t::mulsub_call(a, b, c, d); }
#endif

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    __device__ __forceinline__ static XYZZ zero() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ from_affine(const Affine<F> &p) { return {p.x, p.y, F::one(), F::one()}; }

    // mdbl-2008-s-1: 2 * (affine p)
    __device__ __forceinline__ void set_double_affine(const Affine<F> &p) {
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        y = m * (s - x3) - w * p.y;
        x = x3;
        zz = v;
        zzz = w;
    }
    // dbl-2008-s-1
    __device__ __forceinline__ void dbl() {
        if (is_zero()) return;
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        y = mul_sub(m, s - x3, w, y);
        x = x3;
        zz = v * zz;
        zzz = w * zzz;
    }
    // madd-2008-s with the exceptional cases of ec.rs:446-526 (self = identity, equal points, opposite points).
    // NEG: add -p instead (signed-digit buckets).
    __device__ __forceinline__ void add_mixed(const Affine<F> &p_in, bool neg) {
        Affine<F> p = p_in;
        if (neg) p.y = p.y.neg();
        if (is_zero()) { *this = from_affine(p); return; }
        F u2 = p.x * zz;
        F s2 = p.y * zzz;
        F pp_ = u2 - x;
        F r = s2 - y;
        if (pp_.is_zero()) {
            if (r.is_zero()) set_double_affine(p); else *this = zero();
            return;
        }
        F pp = pp_.sqr();
        F ppp = pp_ * pp;
        F q = x * pp;
        F x3 = r.sqr() - ppp - q.dbl();
        y = mul_sub(r, q - x3, y, ppp);
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }
    // add-2008-s, all exceptional cases handled
    __device__ __forceinline__ void add(const XYZZ &o) {
        if (o.is_zero()) return;
        if (is_zero()) { *this = o; return; }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F pp_ = u2 - u1;
        F r = s2 - s1;
        if (pp_.is_zero()) {
            if (r.is_zero()) dbl(); else *this = zero();
            return;
        }
        F pp = pp_.sqr();
        F ppp = pp_ * pp;
        F q = u1 * pp;
        F x3 = r.sqr() - ppp - q.dbl();
        y = mul_sub(r, q - x3, s1, ppp);
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
    __device__ __forceinline__ Jacobian<F> to_jacobian() const {
        if (is_zero()) return Jacobian<F>::zero();
        return {x * zz, y * zzz, zz};
    }
    __device__ __forceinline__ static XYZZ from_jacobian(const Jacobian<F> &j) {
        if (j.is_zero()) return zero();
        F zz = j.z.sqr();
        return {j.x, j.y, zz, zz * j.z};
    }
};

// ec.rs:586-619 into_affine (one inversion).  Returns false for the identity (x = 0, y = 1 like ec.rs:158-164).
template <class F>
__device__ inline bool jacobian_to_affine(const Jacobian<F> &p, Affine<F> &out) {
    if (p.is_zero()) { out.x = F::zero(); out.y = F::one(); return false; }
    F zi = p.z.inverse();
    F zi2 = zi.sqr();
    out.x = p.x * zi2;
    out.y = p.y * (zi2 * zi);
    return true;
}

// the same for a kernel in which ONE thread normalises one point (the proof assembly): serial binary-Euclid inversion
template <class F>
__device__ inline bool jacobian_to_affine_serial(const Jacobian<F> &p, Affine<F> &out) {
    if (p.is_zero()) { out.x = F::zero(); out.y = F::one(); return false; }
    F zi = p.z.inverse_binary();
    F zi2 = zi.sqr();
    out.x = p.x * zi2;
    out.y = p.y * (zi2 * zi);
    return true;
}

// Jacobian ops exactly as the reference does them -- used by the point-op parity kernels and the final
// (tiny) cross-GPU / window sums where a Jacobian result is the API's output type.
template <class F>
__device__ inline void jacobian_double(Jacobian<F> &p) {  // ec.rs:296-354 dbl-2009-l
    if (p.is_zero()) return;
    F a = p.x.sqr();
    F b = p.y.sqr();
    F c = b.sqr();
    F d = ((p.x + b).sqr() - a - c).dbl();
    F e = a.dbl() + a;
    F f = e.sqr();
    p.z = (p.z * p.y).dbl();
    p.x = f - d - d;
    p.y = (d - p.x) * e - c.dbl().dbl().dbl();
}
template <class F>
__device__ inline void jacobian_add(Jacobian<F> &p, const Jacobian<F> &o) {  // ec.rs:356-444 add-2007-bl
    if (p.is_zero()) { p = o; return; }
    if (o.is_zero()) return;
    F z1z1 = p.z.sqr();
    F z2z2 = o.z.sqr();
    F u1 = p.x * z2z2;
    F u2 = o.x * z1z1;
    F s1 = p.y * o.z * z2z2;
    F s2 = o.y * p.z * z1z1;
    if (u1 == u2 && s1 == s2) { jacobian_double(p); return; }
    F h = u2 - u1;
    F i = h.dbl().sqr();
    F j = h * i;
    F r = (s2 - s1).dbl();
    F v = u1 * i;
    F x3 = r.sqr() - j - v - v;
    p.y = (v - x3) * r - (s1 * j).dbl();
    p.x = x3;
    p.z = ((p.z + o.z).sqr() - z1z1 - z2z2) * h;
}
template <class F>
__device__ inline void jacobian_add_mixed(Jacobian<F> &p, const Affine<F> &o, bool o_inf) {  // ec.rs:446-526
    if (o_inf) return;
    if (p.is_zero()) { p.x = o.x; p.y = o.y; p.z = F::one(); return; }
    F z1z1 = p.z.sqr();
    F u2 = o.x * z1z1;
    F s2 = o.y * p.z * z1z1;
    if (p.x == u2 && p.y == s2) { jacobian_double(p); return; }
    F h = u2 - p.x;
    F hh = h.sqr();
    F i = hh.dbl().dbl();
    F j = h * i;
    F r = (s2 - p.y).dbl();
    F v = p.x * i;
    F x3 = r.sqr() - j - v - v;
    F y3 = (v - x3) * r - (j * p.y).dbl();
    p.z = (p.z + h).sqr() - z1z1 - hh;
    p.x = x3;
    p.y = y3;
}

typedef Affine<fq_t> g1_affine_t;
typedef Affine<fq2_t> g2_affine_t;
typedef XYZZ<fq_t> g1_xyzz_t;
typedef XYZZ<fq2_t> g2_xyzz_t;
typedef Jacobian<fq_t> g1_jac_t;
typedef Jacobian<fq2_t> g2_jac_t;

}  // namespace b200zk
