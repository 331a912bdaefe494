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
template <>  // measured: G1 MSM 2^24 81.1 -> 79.7 ms
__device__ __forceinline__ fq_t mul_sub<fq_t>(const fq_t &a, const fq_t &b, const fq_t &c, const fq_t &d) { return fq_t::mulsub_call(a, b, c, d); }

// (lane-split Fq2, fq2h_t: the generic form.  With the row-fused products a four-product a*b - c*d under one reduction
// (dot_inline<4>: 132 multiplier instructions less per addition) measures the same as two products and a subtraction, 61.6 ms at
// 2^22 -- its 24 KB body pushes the hot path further out of the instruction cache (ncu: no_instruction 0.7 per issue) -- so the
// smaller code stays.
// Before the products were fused row by row (fp.cuh dot_inline), the four-product form lost outright:  as unreduced 2N-limb products added
// up and reduced once it was measured twice -- as two calls handing a 24-limb intermediate over, and as one out-of-line body: 156
// multiplier instructions less per addition, but the row-wise unreduced products and their 24-limb accumulations cost more than the
// reduction they save: G2 2^22 68.0 / 66.9 vs 63.95 ms.)
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
        Affine<F> p;
        p.x = p_in.x;
        p.y = neg ? p_in.y.neg() : p_in.y;
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

// ------------------------------------------------------------------------------------------------ two lanes per point
// A dependent point addition in ONE thread is a ~10 us (G1) / ~25 us (G2) chain of field products, and everything that is not
// the big bucket accumulation (the chains of a small multiexp, the bucket reduction, the window combination, the scalar
// multiplications of the proof assembly) is bound by that latency, not by throughput.  PairXYZZ splits one XYZZ point over two
// adjacent lanes -- lane 0 holds (X, ZZ), lane 1 holds (Y, ZZZ) -- so that the formulas' independent products run side by side
// and the halves meet through warp shuffles:
//     madd-2008-s   6M + 2S + fused Y3   ->  5 product slots   (U2|S2, PP|RR, PPP|Q, ZZ3|ZZZ3, Y3)
//     add-2008-s   10M + 2S + fused Y3   ->  7 product slots
//     dbl-2008-s-1  4M + 3S + fused Y3   ->  4 product slots
// Both lanes execute every instruction (SIMT); values that only one lane needs are garbage on the other and never used.
// All 32 lanes of the warp must call these functions together (full-mask shuffles): idle pairs pass the identity / inactive.
// The exceptional cases of ec.rs:446-526 (equal or opposite points) are rare: when any pair of the warp hits one, both of its
// lanes gather the whole point and run the single-thread formula.
__device__ __forceinline__ fq_t lane_select(bool c, const fq_t &x, const fq_t &y) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < fq_t::N; i++) r.v[i] = c ? x.v[i] : y.v[i];
    return r;
}
__device__ __forceinline__ fq2_t lane_select(bool c, const fq2_t &x, const fq2_t &y) { return {lane_select(c, x.c0, y.c0), lane_select(c, x.c1, y.c1)}; }
__device__ __forceinline__ fq_t lane_swap(const fq_t &x) {  // the partner lane's value
    fq_t r;
#pragma unroll
    for (int i = 0; i < fq_t::N; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], 1);
    return r;
}
__device__ __forceinline__ fq2_t lane_swap(const fq2_t &x) { return {lane_swap(x.c0), lane_swap(x.c1)}; }

// the exceptional cases go through out-of-line single-thread code on a point in local memory: nothing of it is inlined into
// (or keeps registers alive across) the two-lane fast path
template <class F>
__device__ __noinline__ void xyzz_add_mixed_slow(XYZZ<F> *acc, const Affine<F> *p) { acc->add_mixed(*p, false); }
template <class F>
__device__ __noinline__ void xyzz_add_slow(XYZZ<F> *acc, const XYZZ<F> *o) { acc->add(*o); }

template <class F>
struct PairXYZZ {
    F a, b;  // role 0: (X, ZZ); role 1: (Y, ZZZ)
    __device__ __forceinline__ static bool role() { return threadIdx.x & 1u; }
    __device__ __forceinline__ static PairXYZZ zero() { return {F::zero(), F::zero()}; }
    // identity <=> ZZ == 0 (known to lane 0, told to lane 1)
    __device__ __forceinline__ bool is_zero() const { return __shfl_sync(0xffffffffu, (int)b.is_zero(), (threadIdx.x & 31u) & ~1u) != 0; }
    __device__ __forceinline__ static PairXYZZ load(const XYZZ<F> *p) { return role() ? PairXYZZ{p->y, p->zzz} : PairXYZZ{p->x, p->zz}; }
    __device__ __forceinline__ void store(XYZZ<F> *p) const {
        if (role()) { p->y = a; p->zzz = b; } else { p->x = a; p->zz = b; }
    }
    // both lanes get the whole point (slow paths only)
    __device__ __forceinline__ XYZZ<F> gather() const {
        const F pa = lane_swap(a), pb = lane_swap(b);
        return role() ? XYZZ<F>{pa, a, pb, b} : XYZZ<F>{a, pa, b, pb};
    }
    __device__ __forceinline__ static PairXYZZ split(const XYZZ<F> &p) { return role() ? PairXYZZ{p.y, p.zzz} : PairXYZZ{p.x, p.zz}; }
    __device__ __forceinline__ static PairXYZZ from_affine_half(const F &c2) { return {c2, F::one()}; }

    // this += (x2, y2) affine; c2 = this lane's coordinate of the point (x2 on lane 0; y2, already negated for a negative
    // digit, on lane 1); active == false leaves the accumulator alone (an idle pair)
    __device__ __forceinline__ void add_mixed(const F &c2, bool active) {
        const bool r1 = role();
        const unsigned lane = threadIdx.x & 31u;
        const bool self_zero = is_zero();
        const F t = c2 * b;                              // U2 = x2 ZZ1          | S2 = y2 ZZZ1
        const F d = t - a;                               // P = U2 - X1          | R = S2 - Y1
        const bool p_zero = __shfl_sync(0xffffffffu, (int)d.is_zero(), lane & ~1u) != 0;
        const bool special = active && !self_zero && p_zero;  // same x: doubling or P + (-P) (ec.rs:466-470)
        const F e = d.sqr();                             // PP                   | RR
        const F pe = lane_swap(e), pa = lane_swap(a);
        const F f = lane_select(r1, pa, d) * lane_select(r1, pe, e);  // PPP = P PP   | Q = X1 PP
        const F pf = lane_swap(f);
        const F g = b * lane_select(r1, pf, e);          // ZZ3 = ZZ1 PP         | ZZZ3 = ZZZ1 PPP
        const F x3 = e - pf - f.dbl();                   //                      | X3 = RR - PPP - 2Q
        const F px3 = lane_swap(x3);
        const F y3 = mul_sub(d, f - x3, a, pf);          //                      | Y3 = R (Q - X3) - Y1 PPP
        F na = lane_select(r1, y3, px3), nb = g;
        if (__any_sync(0xffffffffu, special)) {
            XYZZ<F> whole = gather();
            const F pc2 = lane_swap(c2);
            if (special) {
                const Affine<F> p = r1 ? Affine<F>{pc2, c2} : Affine<F>{c2, pc2};
                xyzz_add_mixed_slow(&whole, &p);
                const PairXYZZ h = split(whole);
                na = h.a;
                nb = h.b;
            }
        }
        const bool take_pt = active && self_zero;        // identity + P = P
        a = lane_select(take_pt, c2, lane_select(active, na, a));
        b = lane_select(take_pt, F::one(), lane_select(active, nb, b));
    }

    // this += o (add-2008-s), all exceptional cases handled
    __device__ __forceinline__ void add(const PairXYZZ &o) {
        const bool r1 = role();
        const unsigned lane = threadIdx.x & 31u;
        const bool self_zero = is_zero(), other_zero = o.is_zero();
        const F t1 = a * o.b;                            // U1 = X1 ZZ2          | S1 = Y1 ZZZ2
        const F t2 = o.a * b;                            // U2 = X2 ZZ1          | S2 = Y2 ZZZ1
        const F d = t2 - t1;                             // P                    | R
        const bool p_zero = __shfl_sync(0xffffffffu, (int)d.is_zero(), lane & ~1u) != 0;
        const bool special = !self_zero && !other_zero && p_zero;
        const F e = d.sqr();                             // PP                   | RR
        const F zz = b * o.b;                            // ZZ1 ZZ2              | ZZZ1 ZZZ2
        const F pe = lane_swap(e), pt1 = lane_swap(t1);
        const F f = lane_select(r1, pt1, d) * lane_select(r1, pe, e);  // PPP = P PP  | Q = U1 PP
        const F pf = lane_swap(f);
        const F g = zz * lane_select(r1, pf, e);         // ZZ3                  | ZZZ3
        const F x3 = e - pf - f.dbl();                   //                      | X3
        const F px3 = lane_swap(x3);
        const F y3 = mul_sub(d, f - x3, t1, pf);         //                      | Y3 = R (Q - X3) - S1 PPP
        F na = lane_select(r1, y3, px3), nb = g;
        if (__any_sync(0xffffffffu, special)) {
            XYZZ<F> whole = gather();
            const XYZZ<F> other = o.gather();
            if (special) {
                xyzz_add_slow(&whole, &other);
                const PairXYZZ h = split(whole);
                na = h.a;
                nb = h.b;
            }
        }
        // other = identity: keep this; this = identity: take other
        a = lane_select(other_zero, a, lane_select(self_zero, o.a, na));
        b = lane_select(other_zero, b, lane_select(self_zero, o.b, nb));
    }

    // this = 2 this (dbl-2008-s-1); the identity (all zero) stays the identity by the formulas themselves
    __device__ __forceinline__ void dbl() {
        const bool r1 = role();
        const F u = a.dbl();                             //                      | U = 2 Y1
        const F q = lane_select(r1, u, a).sqr();         // XX = X1^2            | V = U^2
        const F pq = lane_swap(q);                       // V                    | XX
        const F w = lane_select(r1, u, a) * lane_select(r1, q, pq);   // S = X1 V    | W = U V
        const F mm = q.dbl() + q;                        // M = 3 XX             |
        const F pmm = lane_swap(mm);                     //                      | M
        const F r3 = lane_select(r1, w, mm) * lane_select(r1, b, mm);  // M^2        | ZZZ3 = W ZZZ1
        const F x3 = r3 - w.dbl();                       // X3 = M^2 - 2 S       |
        const F px3 = lane_swap(x3), pw = lane_swap(w);  //                      | X3, S
        // lane 0: ZZ3 = V ZZ1 - 0 ; lane 1: Y3 = M (S - X3) - W Y1   (one fused two-product body for both)
        const F z = F::zero();
        const F y3 = mul_sub(lane_select(r1, pmm, pq), lane_select(r1, pw - px3, b), lane_select(r1, w, z), lane_select(r1, a, z));
        a = lane_select(r1, y3, x3);
        b = lane_select(r1, r3, y3);
    }

    // Jacobian (X ZZ, Y ZZZ, ZZ) for the C ABI: lane 0 returns X ZZ, lane 1 returns Y ZZZ; *zz_out = ZZ on both lanes
    __device__ __forceinline__ F to_jacobian_half(F *zz_out) const {
        const F pb = lane_swap(b);
        *zz_out = role() ? pb : b;
        return a * b;
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
