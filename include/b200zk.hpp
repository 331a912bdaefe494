// b200zk.hpp -- C++ host-side mirror of the reference's operator interface for the prover hot path, on top of the
// C ABI (b200zk.h).  The reference is compiled code (Rust); this header is what a compiled host links against when it
// cannot use the Rust `-sys` binding of INTEGRATION.md.  Same names, argument meaning and error behaviour as
//
//   bellman::multicore::Worker                                 bellman/src/multicore.rs:13-49
//   bellman::multiexp::{multiexp, FullDensity, DensityTracker}   bellman/src/multiexp.rs:70-138, 285-335
//   bellman::domain::EvaluationDomain                           bellman/src/domain.rs:26-189
//   bellman::groth16::{Parameters, Proof, create_proof}          bellman/src/groth16/{mod.rs:27-53,215-238, prover.rs:205-364}
//   bellman::SynthesisError                                     bellman/src/lib.rs:171-188
//
// Header only; every arithmetic operation happens on the GPU inside libb200zk.so (there is no CPU fallback).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200zk.h"

namespace b200zk {

// bellman/src/lib.rs:171-188
struct SynthesisError : std::runtime_error {
    enum Kind { UnexpectedIdentity = 1, IoErrorUnexpectedEof = 2, PolynomialDegreeTooLarge = 3, BadArgument = 4, Device = 5, Nccl = 6 } kind;
    SynthesisError(Kind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};

typedef std::array<uint64_t, 4> FrRepr;         // canonical scalar (PrimeField::into_repr)
typedef std::array<uint64_t, 4> Fr;             // Montgomery form, as in Rust memory
typedef std::array<uint64_t, 12> G1Affine;      // x || y (infinity carried separately)
typedef std::array<uint64_t, 24> G2Affine;
typedef std::array<uint64_t, 18> G1Projective;  // Jacobian X, Y, Z
typedef std::array<uint64_t, 36> G2Projective;

// bellman::multicore::Worker: here one GPU + one CUDA stream
class Worker {
public:
    explicit Worker(int device = 0) {
        if (b200zk_init(device, &ctx_) != B200ZK_OK) throw SynthesisError(SynthesisError::Device, "b200zk_init failed: no usable CUDA device (no CPU fallback)");
    }
    ~Worker() { b200zk_destroy(ctx_); }
    Worker(const Worker &) = delete;
    Worker &operator=(const Worker &) = delete;
    b200zk_ctx *ctx() const { return ctx_; }
    void check(int st) const {
        if (st == B200ZK_OK) return;
        std::string m = b200zk_last_error(ctx_);
        throw SynthesisError(static_cast<SynthesisError::Kind>(st), m.empty() ? "b200zk error" : m);
    }
private:
    b200zk_ctx *ctx_ = nullptr;
};

// multiexp.rs:78-97 / 99-138
struct FullDensity {};
struct DensityTracker {
    std::vector<uint8_t> bv;  // bit-vec expanded to bytes at the boundary
    size_t total_density = 0;
    void add_element() { bv.push_back(0); }
    void inc(size_t idx) { if (!bv.at(idx)) { bv[idx] = 1; total_density++; } }
    size_t get_total_density() const { return total_density; }
};

// The Arc<Vec<G>> of a SourceBuilder, resident in HBM (multiexp.rs:34-68)
template <int GROUP>
class Bases {
public:
    typedef typename std::conditional<GROUP == B200ZK_G1, G1Affine, G2Affine>::type Affine;
    Bases(const Worker &w, const std::vector<Affine> &points, const std::vector<uint8_t> *infinity = nullptr) : w_(w) {
        w.check(b200zk_bases_upload(w.ctx(), GROUP, points.data(), points.size(), sizeof(Affine), infinity ? infinity->data() : nullptr, 1, &h_));
    }
    ~Bases() { b200zk_bases_free(h_); }
    Bases(const Bases &) = delete;
    size_t len() const { return b200zk_bases_len(h_); }
    void precompute(int window_bits = 0) { w_.check(b200zk_bases_precompute(w_.ctx(), h_, window_bits)); }
    b200zk_bases *handle() const { return h_; }
private:
    const Worker &w_;
    b200zk_bases *h_ = nullptr;
};
typedef Bases<B200ZK_G1> G1Bases;
typedef Bases<B200ZK_G2> G2Bases;

// bellman::multiexp::multiexp (multiexp.rs:285-335): bases = (vector, offset) like the (Arc<Vec<G>>, usize) SourceBuilder
template <int GROUP>
typename std::conditional<GROUP == B200ZK_G1, G1Projective, G2Projective>::type
multiexp(const Worker &pool, std::pair<const Bases<GROUP> *, size_t> bases, const DensityTracker *density_map, const std::vector<FrRepr> &exponents) {
    if (density_map && density_map->bv.size() != exponents.size()) throw std::logic_error("query_size == exponents.len()");  // multiexp.rs:306
    typename std::conditional<GROUP == B200ZK_G1, G1Projective, G2Projective>::type out{};
    pool.check(b200zk_multiexp(pool.ctx(), bases.first->handle(), bases.second, exponents.empty() ? nullptr : exponents[0].data(), exponents.size(),
                               density_map ? density_map->bv.data() : nullptr, out.data()));
    return out;
}
template <int GROUP>
auto multiexp(const Worker &pool, std::pair<const Bases<GROUP> *, size_t> bases, FullDensity, const std::vector<FrRepr> &exponents) {
    return multiexp<GROUP>(pool, bases, static_cast<const DensityTracker *>(nullptr), exponents);
}

// The future multiexp() returns in the reference (`Box<Future<Item = G::Projective, Error = SynthesisError>>`): submit now,
// wait() later; several may be in flight per Worker (prover.rs:289-318, 339-354).  The exponent vector must outlive wait().
template <int GROUP>
class MultiexpFuture {
public:
    typedef typename std::conditional<GROUP == B200ZK_G1, G1Projective, G2Projective>::type Projective;
    MultiexpFuture(const Worker &pool, std::pair<const Bases<GROUP> *, size_t> bases, const DensityTracker *density_map, const std::vector<FrRepr> &exponents)
        : pool_(pool) {
        if (density_map && density_map->bv.size() != exponents.size()) throw std::logic_error("query_size == exponents.len()");
        pool.check(b200zk_multiexp_async(pool.ctx(), bases.first->handle(), bases.second, exponents.empty() ? nullptr : exponents[0].data(),
                                         exponents.size(), density_map ? density_map->bv.data() : nullptr, &job_));
    }
    MultiexpFuture(MultiexpFuture &&o) noexcept : pool_(o.pool_), job_(o.job_) { o.job_ = nullptr; }
    ~MultiexpFuture() { if (job_) b200zk_job_wait(job_, nullptr); }
    Projective wait() {
        Projective out{};
        b200zk_job *j = job_;
        job_ = nullptr;
        pool_.check(b200zk_job_wait(j, out.data()));
        return out;
    }
private:
    const Worker &pool_;
    b200zk_job *job_ = nullptr;
};

// bellman::domain::EvaluationDomain<E, Scalar<E>> with the coefficients resident on the GPU
class EvaluationDomain {
public:
    // domain.rs:48-81
    static EvaluationDomain from_coeffs(const Worker &w, std::vector<Fr> coeffs) {
        size_t m = 1;
        uint32_t exp = 0;
        while (m < coeffs.size()) {
            m *= 2;
            exp++;
            if (exp >= 32) throw SynthesisError(SynthesisError::PolynomialDegreeTooLarge, "PolynomialDegreeTooLarge");
        }
        coeffs.resize(m, Fr{0, 0, 0, 0});
        EvaluationDomain d(w, m, exp);
        w.check(b200zk_dev_alloc(w.ctx(), m * 32, &d.buf_));
        w.check(b200zk_h2d(w.ctx(), d.buf_, coeffs.data(), m * 32));
        w.check(b200zk_sync(w.ctx()));
        return d;
    }
    EvaluationDomain(EvaluationDomain &&o) noexcept : w_(o.w_), m_(o.m_), exp_(o.exp_), buf_(o.buf_) { o.buf_ = nullptr; }
    ~EvaluationDomain() { if (buf_) b200zk_dev_free(w_.ctx(), buf_); }
    size_t len() const { return m_; }
    std::vector<Fr> into_coeffs() const {
        std::vector<Fr> out(m_);
        w_.check(b200zk_d2h(w_.ctx(), out.data(), buf_, m_ * 32));
        return out;
    }
    void fft(const Worker &) { ntt(B200ZK_FFT); }                // domain.rs:83-86
    void ifft(const Worker &) { ntt(B200ZK_IFFT); }              // domain.rs:88-103
    void coset_fft(const Worker &) { ntt(B200ZK_COSET_FFT); }    // domain.rs:120-124
    void icoset_fft(const Worker &) { ntt(B200ZK_ICOSET_FFT); }  // domain.rs:126-132
    void distribute_powers(const Worker &, const Fr &g) { w_.check(b200zk_distribute_powers_dev(w_.ctx(), buf_, m_, g.data())); }
    Fr z(const Fr &tau) const { Fr out; w_.check(b200zk_domain_z(w_.ctx(), tau.data(), exp_, out.data())); return out; }
    void divide_by_z_on_coset(const Worker &) { w_.check(b200zk_divide_by_z_on_coset_dev(w_.ctx(), buf_, exp_)); }
    void mul_assign(const Worker &, const EvaluationDomain &o) { vec(B200ZK_OP_MUL, o); }  // domain.rs:162-175
    void sub_assign(const Worker &, const EvaluationDomain &o) { vec(B200ZK_OP_SUB, o); }  // domain.rs:178-189
private:
    EvaluationDomain(const Worker &w, size_t m, uint32_t exp) : w_(w), m_(m), exp_(exp) {}
    void ntt(int kind) { w_.check(b200zk_ntt_dev(w_.ctx(), buf_, exp_, kind)); }
    void vec(int op, const EvaluationDomain &o) {
        if (o.m_ != m_) throw std::logic_error("assert_eq!(self.coeffs.len(), other.coeffs.len())");
        w_.check(b200zk_field_vec_dev(w_.ctx(), B200ZK_FR, op, buf_, o.buf_, buf_, m_));
    }
    const Worker &w_;
    size_t m_;
    uint32_t exp_;
    void *buf_ = nullptr;
};

// groth16::Parameters resident in HBM (groth16/mod.rs:215-238) = the ParameterSource of create_proof
struct VerifyingKeyPoints {
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, delta_g2;
    std::array<uint8_t, 5> infinity{};  // the `infinity` flags of alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2 (delta at infinity -> UnexpectedIdentity, prover.rs:320-324)
};
class Parameters {
public:
    Parameters(const Worker &w, const std::vector<G1Affine> &h, const std::vector<G1Affine> &l, const std::vector<G1Affine> &a,
               const std::vector<G1Affine> &b_g1, const std::vector<G2Affine> &b_g2, const VerifyingKeyPoints &vk, bool precompute = true)
        : h_(w, h), l_(w, l), a_(w, a), b1_(w, b_g1), b2_(w, b_g2) {
        if (precompute) { h_.precompute(); l_.precompute(); a_.precompute(); b1_.precompute(); b2_.precompute(); }
        w.check(b200zk_crs_create(w.ctx(), h_.handle(), l_.handle(), a_.handle(), b1_.handle(), b2_.handle(), vk.alpha_g1.data(), vk.beta_g1.data(),
                                  vk.beta_g2.data(), vk.delta_g1.data(), vk.delta_g2.data(), vk.infinity.data(), &crs_));
    }
    ~Parameters() { b200zk_crs_free(crs_); }
    Parameters(const Parameters &) = delete;
    const b200zk_crs *handle() const { return crs_; }
private:
    G1Bases h_, l_, a_, b1_;
    G2Bases b2_;
    b200zk_crs *crs_ = nullptr;
};

// groth16::Proof (groth16/mod.rs:27-32), affine Montgomery coordinates + infinity flags
struct Proof { G1Affine a; G2Affine b; G1Affine c; std::array<uint8_t, 3> infinity; };

// What ProvingAssignment holds after synthesis (prover.rs:84-99)
struct ProvingAssignment {
    DensityTracker a_aux_density, b_input_density, b_aux_density;
    std::vector<Fr> a, b, c;                               // evaluations (Montgomery)
    std::vector<FrRepr> input_assignment, aux_assignment;  // into_repr() of the assignments (prover.rs:290-291)
};

// groth16::create_proof after circuit synthesis (prover.rs:249-364)
inline Proof create_proof(const Worker &w, const Parameters &params, const ProvingAssignment &p, const FrRepr &r, const FrRepr &s) {
    if (p.a.empty() || p.b.size() != p.a.size() || p.c.size() != p.a.size() || p.input_assignment.empty() ||
        p.a_aux_density.bv.size() != p.aux_assignment.size() || p.b_aux_density.bv.size() != p.aux_assignment.size() ||
        p.b_input_density.bv.size() != p.input_assignment.size())
        throw std::invalid_argument("ProvingAssignment: a, b, c need one evaluation per constraint and the density maps one entry per variable");
    Proof out{};
    w.check(b200zk_groth16_prove(w.ctx(), params.handle(), p.a[0].data(), p.b[0].data(), p.c[0].data(), p.a.size(), p.input_assignment[0].data(),
                                 p.input_assignment.size(), p.aux_assignment.empty() ? nullptr : p.aux_assignment[0].data(), p.aux_assignment.size(),
                                 p.a_aux_density.bv.data(), p.b_input_density.bv.data(), p.b_aux_density.bv.data(), r.data(), s.data(), out.a.data(),
                                 out.b.data(), out.c.data(), out.infinity.data()));
    return out;
}

// A run of create_proof calls over one CRS and one circuit (the spend proofs of one transaction, rustzcash.rs:1375): proved
// `lockstep` at a time by b200zk_groth16_prove_batch (0 = the library default of 8).  rs[i] = (r, s) of proof i.
inline std::vector<Proof> create_proofs(const Worker &w, const Parameters &params, const std::vector<const ProvingAssignment *> &ps,
                                        const std::vector<std::pair<FrRepr, FrRepr>> &rs, int lockstep = 0) {
    const size_t n = ps.size();
    std::vector<Proof> out(n);
    if (n == 0) return out;
    if (rs.size() != n) throw std::invalid_argument("one (r, s) pair per proof");
    std::vector<b200zk_prove_input> rows(n);
    for (size_t i = 0; i < n; i++) {
        const ProvingAssignment &p = *ps[i];
        if (p.a.size() != ps[0]->a.size() || p.input_assignment.size() != ps[0]->input_assignment.size() ||
            p.aux_assignment.size() != ps[0]->aux_assignment.size())
            throw std::invalid_argument("the proofs of a batch must come from the same circuit");
        rows[i] = b200zk_prove_input{p.a[0].data(), p.b[0].data(), p.c[0].data(), p.input_assignment[0].data(),
                                     p.aux_assignment.empty() ? nullptr : p.aux_assignment[0].data(), p.a_aux_density.bv.data(),
                                     p.b_input_density.bv.data(), p.b_aux_density.bv.data(), rs[i].first.data(), rs[i].second.data()};
    }
    std::vector<uint64_t> a(12 * n), b(24 * n), c(12 * n);
    std::vector<uint8_t> inf(3 * n);
    w.check(b200zk_groth16_prove_batch(w.ctx(), params.handle(), rows.data(), n, ps[0]->a.size(), ps[0]->input_assignment.size(),
                                       ps[0]->aux_assignment.size(), lockstep, a.data(), b.data(), c.data(), inf.data()));
    for (size_t i = 0; i < n; i++) {
        std::copy_n(a.begin() + 12 * i, 12, out[i].a.begin());
        std::copy_n(b.begin() + 24 * i, 24, out[i].b.begin());
        std::copy_n(c.begin() + 12 * i, 12, out[i].c.begin());
        std::copy_n(inf.begin() + 3 * i, 3, out[i].infinity.begin());
    }
    return out;
}

}  // namespace b200zk
