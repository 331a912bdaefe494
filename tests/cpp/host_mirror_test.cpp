// Exercises include/b200zk.hpp (the C++ mirror of bellman's Worker / multiexp / EvaluationDomain) against properties the
// reference's own tests use: fft_composition (domain.rs:426-461) and small exact multiexps.  Built and run by
// tests/test_cpp_mirror.py on the GPU box.
#include <cstdio>
#include <cstdlib>

#include "b200zk.hpp"

using namespace b200zk;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } \
    } while (0)

// fq.rs:81-136 G1 generator, Montgomery limbs
static const G1Affine G1_GEN = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull,
                                0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
// G2 generator (fq.rs:138-264 constants), Montgomery limbs x.c0 x.c1 y.c0 y.c1
static const G2Affine G2_GEN = {0xf5f28fa202940a10ull, 0xb3f5fb2687b4961aull, 0xa1a893b53e2ae580ull, 0x9894999d1a3caee9ull, 0x6f67b7631863366bull, 0x58191924350bcd7ull,
                                0xa5a9c0759e23f606ull, 0xaaa0c59dbccd60c3ull, 0x3bb17e18e2867806ull, 0x1b1ab6cc8541b367ull, 0xc2b6ed0ef2158547ull, 0x11922a097360edf3ull,
                                0x4c730af860494c4aull, 0x597cfa1f5e369c5aull, 0xe7e6856caa0a635aull, 0xbbefb5e96e0d495full, 0x7d3a975f0ef25a2ull, 0x83fd8e7e80dae5ull,
                                0xadc0fc92df64b05dull, 0x18aa270a2b1461dcull, 0x86adac6a3be4eba0ull, 0x79495c4ec93da33aull, 0xe7175850a43ccaedull, 0xb2bc2a163de1bf2ull};
// fr.rs:18-24 R = one in Montgomery form
static const Fr FR_ONE = {0x1fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};

static G1Affine to_affine(const Worker &w, const G1Projective &p, bool *inf) {
    G1Affine out;
    uint8_t f = 0;
    w.check(b200zk_into_affine(w.ctx(), B200ZK_G1, p.data(), 1, out.data(), &f));
    *inf = f != 0;
    return out;
}

int main() {
    Worker w(0);
    // ---- EvaluationDomain: compositions are the identity (domain.rs:426-461)
    std::vector<Fr> v(1000);
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (auto &e : v) {
        for (int i = 0; i < 4; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; e[i] = x; }
        e[3] &= 0x3fffffffffffffffull;  // < r: a valid Montgomery residue
    }
    auto d = EvaluationDomain::from_coeffs(w, v);
    CHECK(d.len() == 1024);
    std::vector<Fr> padded = d.into_coeffs();
    d.ifft(w); d.fft(w);
    CHECK(d.into_coeffs() == padded);
    d.coset_fft(w); d.icoset_fft(w);
    CHECK(d.into_coeffs() == padded);
    // mul_assign by the all-ones vector is the identity; sub_assign of itself gives zero
    auto ones = EvaluationDomain::from_coeffs(w, std::vector<Fr>(1024, FR_ONE));
    d.mul_assign(w, ones);
    CHECK(d.into_coeffs() == padded);
    auto d2 = EvaluationDomain::from_coeffs(w, padded);
    d.sub_assign(w, d2);
    for (auto &e : d.into_coeffs()) CHECK(e == (Fr{0, 0, 0, 0}));
    // z(one) = 1^m - 1 = 0
    CHECK(ones.z(FR_ONE) == (Fr{0, 0, 0, 0}));

    // ---- multiexp: 1 * G == G, 0 * G == identity, 2G + 3G == 5G
    std::vector<G1Affine> pts(2, G1_GEN);
    G1Bases bases(w, pts);
    bool inf;
    auto r1 = multiexp<B200ZK_G1>(w, {&bases, 0}, FullDensity(), std::vector<FrRepr>{FrRepr{1, 0, 0, 0}});
    CHECK(to_affine(w, r1, &inf) == G1_GEN && !inf);
    auto r0 = multiexp<B200ZK_G1>(w, {&bases, 0}, FullDensity(), std::vector<FrRepr>{FrRepr{0, 0, 0, 0}});
    to_affine(w, r0, &inf);
    CHECK(inf);
    auto r23 = multiexp<B200ZK_G1>(w, {&bases, 0}, FullDensity(), std::vector<FrRepr>{FrRepr{2, 0, 0, 0}, FrRepr{3, 0, 0, 0}});
    auto r5 = multiexp<B200ZK_G1>(w, {&bases, 1}, FullDensity(), std::vector<FrRepr>{FrRepr{5, 0, 0, 0}});
    bool i1, i2;
    CHECK(to_affine(w, r23, &i1) == to_affine(w, r5, &i2) && !i1 && !i2);
    // density map: only the second exponent consumes a base
    DensityTracker dt;
    dt.add_element(); dt.add_element(); dt.inc(1);
    auto rd = multiexp<B200ZK_G1>(w, {&bases, 0}, &dt, std::vector<FrRepr>{FrRepr{7, 0, 0, 0}, FrRepr{5, 0, 0, 0}});
    CHECK(to_affine(w, rd, &i1) == to_affine(w, r5, &i2));
    // bases exhausted -> IoError(UnexpectedEof) (multiexp.rs:44-46)
    try {
        multiexp<B200ZK_G1>(w, {&bases, 1}, FullDensity(), std::vector<FrRepr>{FrRepr{2, 0, 0, 0}, FrRepr{3, 0, 0, 0}});
        CHECK(false);
    } catch (const SynthesisError &e) {
        CHECK(e.kind == SynthesisError::IoErrorUnexpectedEof);
    }
    // futures: two multiexps in flight, waited for later (prover.rs:289-318, 339-354)
    {
        std::vector<FrRepr> e1{FrRepr{2, 0, 0, 0}, FrRepr{3, 0, 0, 0}}, e2{FrRepr{5, 0, 0, 0}};
        MultiexpFuture<B200ZK_G1> f1(w, {&bases, 0}, nullptr, e1), f2(w, {&bases, 1}, nullptr, e2);
        auto a1 = to_affine(w, f1.wait(), &i1);
        auto a2 = to_affine(w, f2.wait(), &i2);
        CHECK(a1 == a2 && !i1 && !i2);
    }
    // precomputed tables do not change the result
    bases.precompute(8);
    auto r23p = multiexp<B200ZK_G1>(w, {&bases, 0}, FullDensity(), std::vector<FrRepr>{FrRepr{2, 0, 0, 0}, FrRepr{3, 0, 0, 0}});
    CHECK(to_affine(w, r23p, &i1) == to_affine(w, r5, &i2));
    // ---- create_proof and the lock-step batch agree (the CRS is made of generator points: any group elements do for this)
    {
        const size_t n_con = 5, n_in = 2, n_aux = 3;
        VerifyingKeyPoints vk{G1_GEN, G1_GEN, G1_GEN, G2_GEN, G2_GEN};
        Parameters params(w, std::vector<G1Affine>(7, G1_GEN), std::vector<G1Affine>(n_aux, G1_GEN), std::vector<G1Affine>(n_in + n_aux, G1_GEN),
                          std::vector<G1Affine>(n_in + n_aux, G1_GEN), std::vector<G2Affine>(n_in + n_aux, G2_GEN), vk);
        auto make = [&](uint64_t seed) {
            ProvingAssignment p;
            for (size_t i = 0; i < n_con; i++) {
                p.a.push_back(Fr{seed + i, 1, 2, 3});
                p.b.push_back(Fr{seed * 3 + i, 5, 1, 2});
                p.c.push_back(Fr{seed * 7 + i, 9, 4, 1});
            }
            for (size_t i = 0; i < n_in; i++) { p.input_assignment.push_back(FrRepr{seed + 11 * i + 1, 0, 0, 0}); p.b_input_density.add_element(); }
            for (size_t i = 0; i < n_aux; i++) {
                p.aux_assignment.push_back(FrRepr{seed * 5 + i, i, 0, 0});
                p.a_aux_density.add_element();
                p.b_aux_density.add_element();
            }
            p.a_aux_density.inc(0); p.a_aux_density.inc(2); p.b_aux_density.inc(1); p.b_input_density.inc(0);
            return p;
        };
        ProvingAssignment p1 = make(17), p2 = make(40), p3 = make(99);
        std::vector<std::pair<FrRepr, FrRepr>> rs{{FrRepr{5, 0, 0, 1}, FrRepr{7, 3, 0, 0}}, {FrRepr{6, 0, 0, 1}, FrRepr{8, 3, 0, 0}}, {FrRepr{9, 1, 0, 0}, FrRepr{2, 2, 2, 0}}};
        std::vector<Proof> batch = create_proofs(w, params, {&p1, &p2, &p3}, rs, 2);
        const ProvingAssignment *all[3] = {&p1, &p2, &p3};
        for (int i = 0; i < 3; i++) {
            Proof one = create_proof(w, params, *all[i], rs[i].first, rs[i].second);
            CHECK(one.a == batch[i].a && one.b == batch[i].b && one.c == batch[i].c && one.infinity == batch[i].infinity);
        }
        CHECK(!(batch[0].a == batch[1].a));
    }
    std::printf("host mirror OK\n");
    return 0;
}
