// CPU ORACLE / CPU BASELINE -- test infrastructure only (see oracle/__init__.py).
//
// A C++17 restatement of the reference's CPU path for the Groth16 prover numerics, with
// 64-bit limbs and unsigned __int128 multiply-accumulate (== the reference's `u128-support`
// feature, pairing/src/lib.rs:645-679) and a thread pool that mirrors bellman's Worker
// (bellman/src/multicore.rs:13-82).  The reference itself is Rust and cannot be built in this
// image (no rustc/cargo), so this port doubles as the timed CPU baseline ("kind": "port").
//
// Follows, function by function:
//   fq.rs:813-1123 / fr.rs:341-571   add/sub/double/negate/mul/square/mont_reduce   -> Fp<P>
//   fq2.rs:84-140                    Fq2 mul (3 Fq mul) / square (2 Fq mul)         -> Fp2
//   ec.rs:296-526, 586-619           double / add_assign / add_assign_mixed / into_affine -> Jac<F>
//   multiexp.rs:140-335              multiexp / multiexp_inner (one pool task per window)
//   domain.rs:48-189, 261-374        EvaluationDomain ops, best_fft / serial_fft / parallel_fft
//   prover.rs:256-287                the H-polynomial block
// Nothing in the product (zcash-gpu-thesis_b200/) links or calls this file.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <future>
#include <queue>
#include <memory>
#include <algorithm>

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------------------------------------
// limb primitives, pairing/src/lib.rs:645-679
static inline u64 adc(u64 a, u64 b, u64 &carry) { u128 t = (u128)a + b + carry; carry = (u64)(t >> 64); return (u64)t; }
static inline u64 sbb(u64 a, u64 b, u64 &borrow) { u128 t = ((u128)1 << 64) + a - b - borrow; borrow = (t >> 64) == 0 ? 1 : 0; return (u64)t; }
static inline u64 mac_with_carry(u64 a, u64 b, u64 c, u64 &carry) { u128 t = (u128)a + (u128)b * c + carry; carry = (u64)(t >> 64); return (u64)t; }

// ------------------------------------------------------------------------------------------------
struct FrP {
    static constexpr int N = 4;
    static constexpr u64 MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
    static constexpr u64 R[4] = {0x1fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};
    static constexpr u64 R2[4] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x5d314967254398full, 0x748d9d99f59ff11ull};
    static constexpr u64 INV = 0xfffffffeffffffffull;
};
struct FqP {
    static constexpr int N = 6;
    static constexpr u64 MOD[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
    static constexpr u64 R[6] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};
    static constexpr u64 R2[6] = {0xf4df1f341c341746ull, 0xa76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull};
    static constexpr u64 INV = 0x89f3fffcfffcfffdull;
};

template <class P>
struct Fp {
    static constexpr int N = P::N;
    u64 v[N];

    static Fp zero() { Fp r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
    static Fp one() { Fp r; for (int i = 0; i < N; i++) r.v[i] = P::R[i]; return r; }
    bool is_zero() const { u64 o = 0; for (int i = 0; i < N; i++) o |= v[i]; return o == 0; }
    bool operator==(const Fp &b) const { u64 o = 0; for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i]; return o == 0; }
    bool operator!=(const Fp &b) const { return !(*this == b); }

    static bool geq_mod(const u64 *a) {
        for (int i = N - 1; i >= 0; i--) { if (a[i] > P::MOD[i]) return true; if (a[i] < P::MOD[i]) return false; }
        return true;
    }
    bool is_valid() const { return !geq_mod(v); }
    // fq.rs:1023-1031 `reduce`
    void reduce() { if (geq_mod(v)) { u64 b = 0; for (int i = 0; i < N; i++) v[i] = sbb(v[i], P::MOD[i], b); } }
    void add_assign(const Fp &o) { u64 c = 0; for (int i = 0; i < N; i++) v[i] = adc(v[i], o.v[i], c); reduce(); }
    void dbl() { u64 last = 0; for (int i = 0; i < N; i++) { u64 t = v[i] >> 63; v[i] = (v[i] << 1) | last; last = t; } reduce(); }
    void sub_assign(const Fp &o) {
        // fq.rs:825-834: if other > self, add the modulus first
        bool lt = false;
        for (int i = N - 1; i >= 0; i--) { if (v[i] < o.v[i]) { lt = true; break; } if (v[i] > o.v[i]) break; }
        if (lt) { u64 c = 0; for (int i = 0; i < N; i++) v[i] = adc(v[i], P::MOD[i], c); }
        u64 b = 0; for (int i = 0; i < N; i++) v[i] = sbb(v[i], o.v[i], b);
    }
    void negate() { if (!is_zero()) { u64 b = 0; u64 t[N]; for (int i = 0; i < N; i++) t[i] = sbb(P::MOD[i], v[i], b); for (int i = 0; i < N; i++) v[i] = t[i]; } }
    // fq.rs:1040-1123 mont_reduce (HAC 14.32)
    void mont_reduce(u64 *r) {
        u64 carry2 = 0;
        for (int i = 0; i < N; i++) {
            u64 k = r[i] * P::INV;
            u64 carry = 0;
            mac_with_carry(r[i], k, P::MOD[0], carry);
            for (int j = 1; j < N; j++) r[i + j] = mac_with_carry(r[i + j], k, P::MOD[j], carry);
            r[i + N] = adc(r[i + N], carry2, carry);
            carry2 = carry;
        }
        for (int i = 0; i < N; i++) v[i] = r[N + i];
        reduce();
    }
    // fq.rs:910-963 mul_assign
    void mul_assign(const Fp &o) {
        u64 r[2 * N];
        for (int i = 0; i < 2 * N; i++) r[i] = 0;
        for (int i = 0; i < N; i++) {
            u64 carry = 0;
            for (int j = 0; j < N; j++) r[i + j] = mac_with_carry(r[i + j], v[i], o.v[j], carry);
            r[i + N] = carry;
        }
        mont_reduce(r);
    }
    void square() { Fp t = *this; mul_assign(t); }
    // Montgomery -> canonical, fr.rs:290-303
    void into_repr(u64 *out) const { u64 r[2 * N]; for (int i = 0; i < N; i++) { r[i] = v[i]; r[i + N] = 0; } Fp t; t.mont_reduce(r); for (int i = 0; i < N; i++) out[i] = t.v[i]; }
    static Fp from_repr(const u64 *in) { Fp t; for (int i = 0; i < N; i++) t.v[i] = in[i]; Fp r2; for (int i = 0; i < N; i++) r2.v[i] = P::R2[i]; t.mul_assign(r2); return t; }
    static Fp from_u64(u64 x) { u64 t[N] = {0}; t[0] = x; return from_repr(t); }
    // Field::pow, lib.rs:306-324 (MSB-first square and multiply)
    Fp pow(const u64 *e, int nlimbs) const {
        Fp res = one();
        bool found = false;
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            bool bit = (e[i / 64] >> (i % 64)) & 1;
            if (found) res.square(); else found = bit;
            if (bit) res.mul_assign(*this);
        }
        return res;
    }
    Fp pow64(u64 e) const { return pow(&e, 1); }
    // inverse by Fermat (canonical result equals the reference's binary EEA, fq.rs:849-903)
    Fp inverse() const {
        u64 e[N]; u64 b = 0; for (int i = 0; i < N; i++) e[i] = sbb(P::MOD[i], i == 0 ? 2 : 0, b);
        return pow(e, N);
    }
};
typedef Fp<FrP> Fr;
typedef Fp<FqP> Fq;

// fq2.rs
struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2 &b) const { return c0 == b.c0 && c1 == b.c1; }
    bool operator!=(const Fq2 &b) const { return !(*this == b); }
    void add_assign(const Fq2 &o) { c0.add_assign(o.c0); c1.add_assign(o.c1); }
    void sub_assign(const Fq2 &o) { c0.sub_assign(o.c0); c1.sub_assign(o.c1); }
    void dbl() { c0.dbl(); c1.dbl(); }
    void negate() { c0.negate(); c1.negate(); }
    void square() {  // fq2.rs:84-98
        Fq ab = c0; ab.mul_assign(c1);
        Fq c0c1 = c0; c0c1.add_assign(c1);
        Fq t = c1; t.negate(); t.add_assign(c0); t.mul_assign(c0c1); t.sub_assign(ab);
        c1 = ab; c1.add_assign(ab);
        t.add_assign(ab); c0 = t;
    }
    void mul_assign(const Fq2 &o) {  // fq2.rs:118-132
        Fq aa = c0; aa.mul_assign(o.c0);
        Fq bb = c1; bb.mul_assign(o.c1);
        Fq t = o.c0; t.add_assign(o.c1);
        c1.add_assign(c0); c1.mul_assign(t); c1.sub_assign(aa); c1.sub_assign(bb);
        c0 = aa; c0.sub_assign(bb);
    }
    Fq2 inverse() const {  // fq2.rs:134-153
        Fq t1 = c1; t1.square(); Fq t0 = c0; t0.square(); t0.add_assign(t1);
        Fq ti = t0.inverse();
        Fq2 r; r.c0 = c0; r.c0.mul_assign(ti); r.c1 = c1; r.c1.mul_assign(ti); r.c1.negate();
        return r;
    }
};

// ------------------------------------------------------------------------------------------------
// ec.rs: Jacobian points over F (Fq for G1, Fq2 for G2)
template <class F>
struct Aff { F x, y; bool inf; };

template <class F>
struct Jac {
    F x, y, z;
    static Jac zero() { return {F::zero(), F::one(), F::zero()}; }
    bool is_zero() const { return z.is_zero(); }
    void dbl() {  // ec.rs:296-354
        if (is_zero()) return;
        F a = x; a.square();
        F b = y; b.square();
        F c = b; c.square();
        F d = x; d.add_assign(b); d.square(); d.sub_assign(a); d.sub_assign(c); d.dbl();
        F e = a; e.dbl(); e.add_assign(a);
        F f = e; f.square();
        z.mul_assign(y); z.dbl();
        x = f; x.sub_assign(d); x.sub_assign(d);
        y = d; y.sub_assign(x); y.mul_assign(e);
        c.dbl(); c.dbl(); c.dbl();
        y.sub_assign(c);
    }
    void add_assign(const Jac &o) {  // ec.rs:356-444
        if (is_zero()) { *this = o; return; }
        if (o.is_zero()) return;
        F z1z1 = z; z1z1.square();
        F z2z2 = o.z; z2z2.square();
        F u1 = x; u1.mul_assign(z2z2);
        F u2 = o.x; u2.mul_assign(z1z1);
        F s1 = y; s1.mul_assign(o.z); s1.mul_assign(z2z2);
        F s2 = o.y; s2.mul_assign(z); s2.mul_assign(z1z1);
        if (u1 == u2 && s1 == s2) { dbl(); return; }
        F h = u2; h.sub_assign(u1);
        F i = h; i.dbl(); i.square();
        F j = h; j.mul_assign(i);
        F r = s2; r.sub_assign(s1); r.dbl();
        F v = u1; v.mul_assign(i);
        x = r; x.square(); x.sub_assign(j); x.sub_assign(v); x.sub_assign(v);
        y = v; y.sub_assign(x); y.mul_assign(r);
        s1.mul_assign(j); s1.dbl();
        y.sub_assign(s1);
        z.add_assign(o.z); z.square(); z.sub_assign(z1z1); z.sub_assign(z2z2); z.mul_assign(h);
    }
    void add_assign_mixed(const Aff<F> &o) {  // ec.rs:446-526
        if (o.inf) return;
        if (is_zero()) { x = o.x; y = o.y; z = F::one(); return; }
        F z1z1 = z; z1z1.square();
        F u2 = o.x; u2.mul_assign(z1z1);
        F s2 = o.y; s2.mul_assign(z); s2.mul_assign(z1z1);
        if (x == u2 && y == s2) { dbl(); return; }
        F h = u2; h.sub_assign(x);
        F hh = h; hh.square();
        F i = hh; i.dbl(); i.dbl();
        F j = h; j.mul_assign(i);
        F r = s2; r.sub_assign(y); r.dbl();
        F v = x; v.mul_assign(i);
        x = r; x.square(); x.sub_assign(j); x.sub_assign(v); x.sub_assign(v);
        j.mul_assign(y); j.dbl();
        y = v; y.sub_assign(x); y.mul_assign(r); y.sub_assign(j);
        z.add_assign(h); z.square(); z.sub_assign(z1z1); z.sub_assign(hh);
    }
    Aff<F> into_affine() const {  // ec.rs:586-619
        if (is_zero()) return {F::zero(), F::one(), true};
        F zi = z.inverse();
        F zi2 = zi; zi2.square();
        Aff<F> r; r.inf = false;
        r.x = x; r.x.mul_assign(zi2);
        zi2.mul_assign(zi);
        r.y = y; r.y.mul_assign(zi2);
        return r;
    }
};

// ec.rs:87-99 mul_bits (MSB first)
template <class F>
static Jac<F> affine_mul(const Aff<F> &p, const u64 *k, int nlimbs) {
    Jac<F> res = Jac<F>::zero();
    bool found = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        bool bit = (k[i / 64] >> (i % 64)) & 1;
        if (found) res.dbl(); else found = bit;
        if (bit) res.add_assign_mixed(p);
    }
    return res;
}

// ------------------------------------------------------------------------------------------------
// bellman::multicore::Worker (multicore.rs): a fixed pool + futures
class Worker {
public:
    explicit Worker(int cpus) : cpus_(cpus < 1 ? 1 : cpus), stop_(false) {
        for (int i = 0; i < cpus_; i++) threads_.emplace_back([this] { run(); });
    }
    ~Worker() {
        { std::unique_lock<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int cpus() const { return cpus_; }
    int log_num_cpus() const { int l = 0; while ((1 << (l + 1)) <= cpus_) l++; return l; }  // multicore.rs:15-21
    template <class Fn>
    auto compute(Fn fn) -> std::future<decltype(fn())> {
        auto task = std::make_shared<std::packaged_task<decltype(fn())()>>(std::move(fn));
        auto fut = task->get_future();
        { std::unique_lock<std::mutex> l(m_); q_.push([task] { (*task)(); }); }
        cv_.notify_one();
        return fut;
    }
    // Worker::scope: chunk = elements / cpus (multicore.rs:51-67); runs f(chunk_index, begin, end) on scoped threads
    template <class Fn>
    void scope(size_t elements, Fn fn) const {
        size_t chunk = elements < (size_t)cpus_ ? 1 : elements / cpus_;
        std::vector<std::thread> ts;
        size_t idx = 0;
        for (size_t b = 0; b < elements; b += chunk, idx++) {
            size_t e = std::min(elements, b + chunk);
            ts.emplace_back([=] { fn(idx, b, e); });
        }
        for (auto &t : ts) t.join();
    }
private:
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                job = std::move(q_.front()); q_.pop();
            }
            job();
        }
    }
    int cpus_;
    bool stop_;
    std::vector<std::thread> threads_;
    std::queue<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_;
};

enum { CREF_OK = 0, CREF_UNEXPECTED_IDENTITY = 1, CREF_UNEXPECTED_EOF = 2, CREF_DEGREE_TOO_LARGE = 3, CREF_BAD_ARG = 4 };

// multiexp.rs:140-233, one region
template <class F>
static int multiexp_region(const Aff<F> *bases, size_t nbases, size_t base_offset, const uint8_t *density,
                           const u64 *exps, size_t nexp, uint32_t skip, uint32_t c, bool handle_trivial, Jac<F> &out) {
    Jac<F> acc = Jac<F>::zero();
    size_t idx = base_offset;
    std::vector<Jac<F>> buckets((size_t(1) << c) - 1, Jac<F>::zero());
    for (size_t i = 0; i < nexp; i++) {
        if (density && !density[i]) continue;
        const u64 *e = exps + 4 * i;
        bool is_zero = (e[0] | e[1] | e[2] | e[3]) == 0;
        bool is_one = e[0] == 1 && (e[1] | e[2] | e[3]) == 0;
        if (nbases <= idx) return CREF_UNEXPECTED_EOF;  // both skip() and add_assign_mixed() check this first
        if (is_zero) { idx++; continue; }
        if (is_one) {
            if (handle_trivial) { if (bases[idx].inf) return CREF_UNEXPECTED_IDENTITY; acc.add_assign_mixed(bases[idx]); }
            idx++; continue;
        }
        // exp.shr(skip); exp.as_ref()[0] % (1 << c)
        uint32_t limb = skip / 64, sh = skip % 64;
        u64 w = limb < 4 ? e[limb] >> sh : 0;
        if (sh && limb + 1 < 4) w |= e[limb + 1] << (64 - sh);
        u64 digit = w & ((u64(1) << c) - 1);
        if (digit != 0) { if (bases[idx].inf) return CREF_UNEXPECTED_IDENTITY; buckets[digit - 1].add_assign_mixed(bases[idx]); }
        idx++;
    }
    Jac<F> running = Jac<F>::zero();
    for (size_t b = buckets.size(); b-- > 0;) { running.add_assign(buckets[b]); acc.add_assign(running); }
    out = acc;
    return CREF_OK;
}

static uint32_t window_size(size_t n) {  // multiexp.rs:296-300
    if (n < 32) return 3;
    return (uint32_t)std::ceil(std::log((double)(uint32_t)n));
}

template <class F>
static int multiexp(Worker &pool, const Aff<F> *bases, size_t nbases, size_t base_offset, const uint8_t *density,
                    const u64 *exps, size_t nexp, Jac<F> &out) {
    uint32_t c = window_size(nexp);
    struct Res { int st; Jac<F> p; };
    std::vector<std::future<Res>> futs;
    for (uint32_t skip = 0; skip < 255; skip += c) {
        bool ht = skip == 0;
        futs.push_back(pool.compute([=]() { Res r; r.st = multiexp_region<F>(bases, nbases, base_offset, density, exps, nexp, skip, c, ht, r.p); return r; }));
    }
    std::vector<Res> rs;
    for (auto &f : futs) rs.push_back(f.get());
    for (auto &r : rs) if (r.st != CREF_OK) return r.st;  // Join polls the lowest region first
    Jac<F> acc = rs.back().p;
    for (size_t w = rs.size() - 1; w-- > 0;) {
        for (uint32_t k = 0; k < c; k++) acc.dbl();
        acc.add_assign(rs[w].p);
    }
    out = acc;
    return CREF_OK;
}

// ------------------------------------------------------------------------------------------------
// domain.rs
static uint32_t bitreverse(uint32_t n, uint32_t l) { uint32_t r = 0; for (uint32_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; } return r; }

static void serial_fft(Fr *a, const Fr &omega, uint32_t log_n) {  // domain.rs:272-315
    uint32_t n = 1u << log_n;
    for (uint32_t k = 0; k < n; k++) { uint32_t rk = bitreverse(k, log_n); if (k < rk) std::swap(a[rk], a[k]); }
    uint32_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        Fr w_m = omega.pow64(n / (2 * m));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            Fr w = Fr::one();
            for (uint32_t j = 0; j < m; j++) {
                Fr t = a[k + j + m]; t.mul_assign(w);
                Fr tmp = a[k + j]; tmp.sub_assign(t);
                a[k + j + m] = tmp;
                a[k + j].add_assign(t);
                w.mul_assign(w_m);
            }
        }
        m *= 2;
    }
}

static void parallel_fft(Fr *a, const Worker &worker, const Fr &omega, uint32_t log_n, uint32_t log_cpus) {  // domain.rs:317-374
    uint32_t num_cpus = 1u << log_cpus, log_new_n = log_n - log_cpus;
    std::vector<std::vector<Fr>> tmp(num_cpus, std::vector<Fr>(size_t(1) << log_new_n, Fr::zero()));
    Fr new_omega = omega.pow64(num_cpus);
    {
        std::vector<std::thread> ts;
        for (uint32_t j = 0; j < num_cpus; j++) {
            ts.emplace_back([&, j] {
                Fr *t = tmp[j].data();
                Fr omega_j = omega.pow64(j);
                Fr omega_step = omega.pow64((u64)j << log_new_n);
                Fr elt = Fr::one();
                for (uint32_t i = 0; i < (1u << log_new_n); i++) {
                    for (uint32_t s = 0; s < num_cpus; s++) {
                        uint32_t idx = (i + (s << log_new_n)) % (1u << log_n);
                        Fr x = a[idx]; x.mul_assign(elt);
                        t[i].add_assign(x);
                        elt.mul_assign(omega_step);
                    }
                    elt.mul_assign(omega_j);
                }
                serial_fft(t, new_omega, log_new_n);
            });
        }
        for (auto &t : ts) t.join();
    }
    uint32_t mask = (1u << log_cpus) - 1;
    worker.scope(size_t(1) << log_n, [&](size_t, size_t b, size_t e) {
        for (size_t idx = b; idx < e; idx++) a[idx] = tmp[idx & mask][idx >> log_cpus];
    });
}

static void best_fft(Fr *a, const Worker &worker, const Fr &omega, uint32_t log_n) {  // domain.rs:261-270
    uint32_t log_cpus = worker.log_num_cpus();
    if (log_n <= log_cpus) serial_fft(a, omega, log_n); else parallel_fft(a, worker, omega, log_n, log_cpus);
}

struct Domain {  // domain.rs:26-81
    Fr *coeffs; size_t m; uint32_t exp; Fr omega, omegainv, geninv, minv;
    bool init(Fr *c, uint32_t log_m) {
        if (log_m >= 32) return false;
        coeffs = c; exp = log_m; m = size_t(1) << log_m;
        static const u64 ROOT[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0xaf53ae352a31e64ull, 0x5bf3adda19e9b27bull};  // fr.rs:50-55
        for (int i = 0; i < 4; i++) omega.v[i] = ROOT[i];
        for (uint32_t i = exp; i < 32; i++) omega.square();
        omegainv = omega.inverse();
        geninv = Fr::from_u64(7).inverse();
        minv = Fr::from_u64(m).inverse();
        return true;
    }
    void scale(const Worker &w, const Fr &s) { w.scope(m, [&](size_t, size_t b, size_t e) { for (size_t i = b; i < e; i++) coeffs[i].mul_assign(s); }); }
    void fft(const Worker &w) { best_fft(coeffs, w, omega, exp); }
    void ifft(const Worker &w) { best_fft(coeffs, w, omegainv, exp); scale(w, minv); }
    void distribute_powers(const Worker &w, const Fr &g) {  // domain.rs:105-118
        w.scope(m, [&](size_t, size_t b, size_t e) { Fr u = g.pow64(b); for (size_t i = b; i < e; i++) { coeffs[i].mul_assign(u); u.mul_assign(g); } });
    }
    void coset_fft(const Worker &w) { distribute_powers(w, Fr::from_u64(7)); fft(w); }
    void icoset_fft(const Worker &w) { ifft(w); distribute_powers(w, geninv); }
    void divide_by_z_on_coset(const Worker &w) { Fr z = Fr::from_u64(7).pow64(m); z.sub_assign(Fr::one()); scale(w, z.inverse()); }
    void mul_assign(const Worker &w, const Domain &o) { w.scope(m, [&](size_t, size_t b, size_t e) { for (size_t i = b; i < e; i++) coeffs[i].mul_assign(o.coeffs[i]); }); }
    void sub_assign(const Worker &w, const Domain &o) { w.scope(m, [&](size_t, size_t b, size_t e) { for (size_t i = b; i < e; i++) coeffs[i].sub_assign(o.coeffs[i]); }); }
};

// ------------------------------------------------------------------------------------------------
// C ABI for ctypes (tests/, bench.py cpu_baseline only)
static Worker *get_worker(int threads) {
    static std::mutex m; static std::unique_ptr<Worker> w; static int cur = 0;
    std::unique_lock<std::mutex> l(m);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (!w || cur != threads) { w.reset(new Worker(threads)); cur = threads; }
    return w.get();
}

template <class F> static void load_aff(std::vector<Aff<F>> &out, const u64 *xy, const uint8_t *inf, size_t n) {
    out.resize(n);
    for (size_t i = 0; i < n; i++) { memcpy((void *)&out[i].x, xy + i * (2 * sizeof(F) / 8), 2 * sizeof(F)); out[i].inf = inf ? inf[i] != 0 : false; }
}

extern "C" {
int cref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// op: 0 add 1 sub 2 mul 3 square 4 double 5 negate 6 into_repr 7 from_repr 8 inverse
#define VEC_OP(NAME, T)                                                                                     \
    void NAME(int op, const u64 *a, const u64 *b, u64 *out, size_t n) {                                     \
        for (size_t i = 0; i < n; i++) {                                                                    \
            T x, y; memcpy(x.v, a + i * T::N, sizeof(x.v)); if (b) memcpy(y.v, b + i * T::N, sizeof(y.v));   \
            switch (op) {                                                                                   \
            case 0: x.add_assign(y); break; case 1: x.sub_assign(y); break; case 2: x.mul_assign(y); break; \
            case 3: x.square(); break; case 4: x.dbl(); break; case 5: x.negate(); break;                   \
            case 6: { u64 t[T::N]; x.into_repr(t); memcpy(x.v, t, sizeof(t)); } break;                      \
            case 7: x = T::from_repr(x.v); break; case 8: x = x.inverse(); break; }                         \
            memcpy(out + i * T::N, x.v, sizeof(x.v));                                                       \
        }                                                                                                   \
    }
VEC_OP(cref_fr_vec, Fr)
VEC_OP(cref_fq_vec, Fq)

// out[i] = scalars[i] * base as affine x||y (Montgomery); inf_out[i] = 1 for the identity
void cref_g1_scalar_muls(const u64 *base_xy, const u64 *scalars, size_t n, u64 *out_xy, uint8_t *inf_out, int threads) {
    Aff<Fq> b; memcpy((void *)&b.x, base_xy, 96); b.inf = false;
    get_worker(threads)->scope(n, [&](size_t, size_t s, size_t e) {
        for (size_t i = s; i < e; i++) { Aff<Fq> r = affine_mul(b, scalars + 4 * i, 4).into_affine(); memcpy(out_xy + 12 * i, (void *)&r.x, 96); if (inf_out) inf_out[i] = r.inf; }
    });
}
void cref_g2_scalar_muls(const u64 *base_xy, const u64 *scalars, size_t n, u64 *out_xy, uint8_t *inf_out, int threads) {
    Aff<Fq2> b; memcpy((void *)&b.x, base_xy, 192); b.inf = false;
    get_worker(threads)->scope(n, [&](size_t, size_t s, size_t e) {
        for (size_t i = s; i < e; i++) { Aff<Fq2> r = affine_mul(b, scalars + 4 * i, 4).into_affine(); memcpy(out_xy + 24 * i, (void *)&r.x, 192); if (inf_out) inf_out[i] = r.inf; }
    });
}

int cref_g1_multiexp(const u64 *bases_xy, const uint8_t *inf, size_t nbases, size_t base_offset, const u64 *scalars, size_t nexp,
                     const uint8_t *density, u64 *out_jac, int threads) {
    std::vector<Aff<Fq>> b; load_aff(b, bases_xy, inf, nbases);
    Jac<Fq> r; int st = multiexp<Fq>(*get_worker(threads), b.data(), nbases, base_offset, density, scalars, nexp, r);
    if (st == CREF_OK) memcpy(out_jac, (void *)&r, 144);
    return st;
}
int cref_g2_multiexp(const u64 *bases_xy, const uint8_t *inf, size_t nbases, size_t base_offset, const u64 *scalars, size_t nexp,
                     const uint8_t *density, u64 *out_jac, int threads) {
    std::vector<Aff<Fq2>> b; load_aff(b, bases_xy, inf, nbases);
    Jac<Fq2> r; int st = multiexp<Fq2>(*get_worker(threads), b.data(), nbases, base_offset, density, scalars, nexp, r);
    if (st == CREF_OK) memcpy(out_jac, (void *)&r, 288);
    return st;
}
// ---- resident base vectors (an `Arc<Vec<G::Affine>>` the caller keeps, multiexp.rs:34-40): load once, time only multiexp()
struct BasesG1 { std::vector<Aff<Fq>> v; };
struct BasesG2 { std::vector<Aff<Fq2>> v; };
void *cref_g1_bases_load(const u64 *xy, const uint8_t *inf, size_t n) { auto *b = new BasesG1(); load_aff(b->v, xy, inf, n); return b; }
void *cref_g2_bases_load(const u64 *xy, const uint8_t *inf, size_t n) { auto *b = new BasesG2(); load_aff(b->v, xy, inf, n); return b; }
void cref_g1_bases_free(void *h) { delete (BasesG1 *)h; }
void cref_g2_bases_free(void *h) { delete (BasesG2 *)h; }
int cref_g1_multiexp_h(const void *h, size_t base_offset, const u64 *scalars, size_t nexp, const uint8_t *density, u64 *out_jac, int threads) {
    const auto &b = ((const BasesG1 *)h)->v;
    Jac<Fq> r; int st = multiexp<Fq>(*get_worker(threads), b.data(), b.size(), base_offset, density, scalars, nexp, r);
    if (st == CREF_OK) memcpy(out_jac, (void *)&r, 144);
    return st;
}
int cref_g2_multiexp_h(const void *h, size_t base_offset, const u64 *scalars, size_t nexp, const uint8_t *density, u64 *out_jac, int threads) {
    const auto &b = ((const BasesG2 *)h)->v;
    Jac<Fq2> r; int st = multiexp<Fq2>(*get_worker(threads), b.data(), b.size(), base_offset, density, scalars, nexp, r);
    if (st == CREF_OK) memcpy(out_jac, (void *)&r, 288);
    return st;
}
// Synthetic base vector for the CPU arm of the bench (input preparation, not the timed path): P_i = start + i * step as affine
// points, straight into a resident vector.  Chunks run on the pool; each normalises its Jacobian walk with one inversion
// (batch_normalization, ec.rs:246-294).  Returns the handle; `first_xy` (optional) receives the first min(n, n_first) points.
void *cref_g1_bases_walk(const u64 *start_xy, const u64 *step_xy, size_t n, u64 *first_xy, size_t n_first, int threads) {
    Aff<Fq> s0, st; memcpy((void *)&s0.x, start_xy, 96); s0.inf = false; memcpy((void *)&st.x, step_xy, 96); st.inf = false;
    auto *out = new BasesG1(); out->v.resize(n);
    get_worker(threads)->scope(n, [&](size_t, size_t b, size_t e) {
        u64 k[4] = {b, 0, 0, 0};
        Jac<Fq> cur = affine_mul(st, k, 4);  // b * step
        cur.add_assign_mixed(s0);
        const size_t B = 1024;
        std::vector<Jac<Fq>> blk(B); std::vector<Fq> pre(B);
        for (size_t i0 = b; i0 < e; i0 += B) {
            size_t cnt = std::min(B, e - i0);
            Fq run = Fq::one();
            for (size_t j = 0; j < cnt; j++) { blk[j] = cur; pre[j] = run; run.mul_assign(cur.z); cur.add_assign_mixed(st); }
            Fq inv = run.inverse();
            for (size_t j = cnt; j-- > 0;) {
                Fq zi = inv; zi.mul_assign(pre[j]); inv.mul_assign(blk[j].z);
                Fq zi2 = zi; zi2.square();
                Aff<Fq> a; a.inf = false; a.x = blk[j].x; a.x.mul_assign(zi2); zi2.mul_assign(zi); a.y = blk[j].y; a.y.mul_assign(zi2);
                out->v[i0 + j] = a;
            }
        }
    });
    for (size_t i = 0; i < std::min(n, n_first); i++) memcpy(first_xy + 12 * i, (void *)&out->v[i].x, 96);
    return out;
}
// Jacobian (Montgomery limbs) -> affine x||y, returns 1 if infinity
int cref_g1_into_affine(const u64 *jac, u64 *out_xy) { Jac<Fq> p; memcpy((void *)&p, jac, 144); Aff<Fq> a = p.into_affine(); memcpy(out_xy, (void *)&a.x, 96); return a.inf; }
int cref_g2_into_affine(const u64 *jac, u64 *out_xy) { Jac<Fq2> p; memcpy((void *)&p, jac, 288); Aff<Fq2> a = p.into_affine(); memcpy(out_xy, (void *)&a.x, 192); return a.inf; }
// generic point ops for device parity tests: op 0 = double, 1 = add (jac+jac), 2 = add_mixed (jac + affine xy, inf flag in b_inf)
void cref_g1_point_op(int op, const u64 *a_jac, const u64 *b, int b_inf, u64 *out_jac) {
    Jac<Fq> p; memcpy((void *)&p, a_jac, 144);
    if (op == 0) p.dbl();
    else if (op == 1) { Jac<Fq> q; memcpy((void *)&q, b, 144); p.add_assign(q); }
    else { Aff<Fq> q; memcpy((void *)&q.x, b, 96); q.inf = b_inf; p.add_assign_mixed(q); }
    memcpy(out_jac, (void *)&p, 144);
}
void cref_g2_point_op(int op, const u64 *a_jac, const u64 *b, int b_inf, u64 *out_jac) {
    Jac<Fq2> p; memcpy((void *)&p, a_jac, 288);
    if (op == 0) p.dbl();
    else if (op == 1) { Jac<Fq2> q; memcpy((void *)&q, b, 288); p.add_assign(q); }
    else { Aff<Fq2> q; memcpy((void *)&q.x, b, 192); q.inf = b_inf; p.add_assign_mixed(q); }
    memcpy(out_jac, (void *)&p, 288);
}

// kind: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft.  coeffs: m x 4 Montgomery limbs, in place. threads<=0: all cores.
// serial != 0 forces serial_fft (log_cpus = 0).
int cref_fft(u64 *coeffs, uint32_t log_m, int kind, int threads, int serial) {
    Domain d; if (!d.init((Fr *)coeffs, log_m)) return CREF_DEGREE_TOO_LARGE;
    Worker *w = get_worker(serial ? 1 : threads);
    switch (kind) { case 0: d.fft(*w); break; case 1: d.ifft(*w); break; case 2: d.coset_fft(*w); break; case 3: d.icoset_fft(*w); break; default: return CREF_BAD_ARG; }
    return CREF_OK;
}
// explicit parallel_fft with a chosen log_cpus (domain.rs:464-494 parallel_fft_consistency)
int cref_parallel_fft(u64 *coeffs, uint32_t log_m, uint32_t log_cpus, int threads) {
    Domain d; if (!d.init((Fr *)coeffs, log_m)) return CREF_DEGREE_TOO_LARGE;
    if (log_cpus > log_m) return CREF_BAD_ARG;
    parallel_fft((Fr *)coeffs, *get_worker(threads), d.omega, log_m, log_cpus);
    return CREF_OK;
}
// prover.rs:256-287: a,b,c evaluation vectors (m x 4 Montgomery, clobbered) -> out (m-1) x 4 canonical limbs
int cref_h_poly(u64 *a, u64 *b, u64 *c, uint32_t log_m, u64 *out, int threads) {
    Worker *w = get_worker(threads);
    Domain da, db, dc;
    if (!da.init((Fr *)a, log_m) || !db.init((Fr *)b, log_m) || !dc.init((Fr *)c, log_m)) return CREF_DEGREE_TOO_LARGE;
    da.ifft(*w); da.coset_fft(*w);
    db.ifft(*w); db.coset_fft(*w);
    dc.ifft(*w); dc.coset_fft(*w);
    da.mul_assign(*w, db); da.sub_assign(*w, dc); da.divide_by_z_on_coset(*w); da.icoset_fft(*w);
    for (size_t i = 0; i + 1 < da.m; i++) da.coeffs[i].into_repr(out + 4 * i);
    return CREF_OK;
}
}  // extern "C"
