// groth16::verify_proof on the device (bellman/src/groth16/verifier.rs:18-66) and the BLS12-381 pairing under it
// (pairing/src/bls12_381/mod.rs:40-160 miller_loop / final_exponentiation; fq6.rs, fq12.rs for the tower).
//
// The prover's hot path never needs a pairing; this closes the prove -> verify loop of the outer FFI
// (librustzcash/src/rustzcash.rs:1556-1601 verifies every proof it has just made) without a CPU stall, as a BATCH verifier: one
// thread block per proof, many proofs per launch.
//
// Tower: Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = u + 1 (so w^6 = xi).  Tower elements live in
// local memory and every operation is an out-of-line function on pointers: a verification is a long chain of dependent field
// operations in a few threads, the code must stay small, not fast per thread.
// Miller loop: affine steps on the twist E'(Fq2): y^2 = x^3 + 4 xi.  With psi(x, y) = (x / w^2, y / w^3) the tangent / chord through
// psi(T) evaluated at P = (xp, yp) in G1 and scaled by w^3 (an element of the proper subfield Fq4, killed by the final
// exponentiation) is the sparse element  (lambda xt - yt) - (lambda xp) v + yp v w,  lambda = the slope on the twist.
// e(P, Q) after the final exponentiation is canonical, so it equals the reference's value bit for bit (the RELIC known-answer
// test of pairing/src/bls12_381/tests/mod.rs:5-53 is part of the parity suite).
#include <vector>

#include "ec.cuh"
#include "internal.h"
#include "pairing_consts.cuh"

namespace b200zk {

struct fq6_t { fq2_t c0, c1, c2; };
struct fq12_t { fq6_t c0, c1; };
static constexpr uint64_t BLS_X = 0xd201000000010000ull;  // mod.rs:24-25; the curve parameter is -BLS_X

__device__ __forceinline__ fq2_t mul_by_xi(const fq2_t &a) { return {a.c0 - a.c1, a.c0 + a.c1}; }  // (c0 + c1 u)(1 + u), fq2.rs:51-58
__device__ __forceinline__ fq2_t fq2_frob(const fq2_t &a, int k) { return (k & 1) ? fq2_t{a.c0, a.c1.neg()} : a; }  // fq2.rs:155-158
__device__ __forceinline__ fq2_t frob_const(int k, int m) {
    fq2_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) { r.c0.v[i] = PAIRING_FROB[k - 1][m][i]; r.c1.v[i] = PAIRING_FROB[k - 1][m][12 + i]; }
    return r;
}

// ---- Fq6 (fq6.rs:  add / sub / neg componentwise, mul :255-290, mul_by_nonresidue :32-36, inverse :300-340)
__device__ __noinline__ void f6_add(fq6_t *r, const fq6_t *a, const fq6_t *b) { r->c0 = a->c0 + b->c0; r->c1 = a->c1 + b->c1; r->c2 = a->c2 + b->c2; }
__device__ __noinline__ void f6_sub(fq6_t *r, const fq6_t *a, const fq6_t *b) { r->c0 = a->c0 - b->c0; r->c1 = a->c1 - b->c1; r->c2 = a->c2 - b->c2; }
__device__ __noinline__ void f6_neg(fq6_t *r, const fq6_t *a) { r->c0 = a->c0.neg(); r->c1 = a->c1.neg(); r->c2 = a->c2.neg(); }
__device__ __noinline__ void f6_mul(fq6_t *r, const fq6_t *a, const fq6_t *b) {
    // (a0 + a1 v + a2 v^2)(b0 + b1 v + b2 v^2), v^3 = xi
    const fq2_t c0 = a->c0 * b->c0 + mul_by_xi(a->c1 * b->c2 + a->c2 * b->c1);
    const fq2_t c1 = a->c0 * b->c1 + a->c1 * b->c0 + mul_by_xi(a->c2 * b->c2);
    const fq2_t c2 = a->c0 * b->c2 + a->c1 * b->c1 + a->c2 * b->c0;
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
__device__ __noinline__ void f6_mul_by_v(fq6_t *r, const fq6_t *a) {
    const fq2_t t = mul_by_xi(a->c2);
    r->c2 = a->c1; r->c1 = a->c0; r->c0 = t;
}
__device__ __noinline__ void f6_inv(fq6_t *r, const fq6_t *a) {
    const fq2_t t0 = a->c0.sqr() - mul_by_xi(a->c1 * a->c2);
    const fq2_t t1 = mul_by_xi(a->c2.sqr()) - a->c0 * a->c1;
    const fq2_t t2 = a->c1.sqr() - a->c0 * a->c2;
    const fq2_t d = a->c0 * t0 + mul_by_xi(a->c2 * t1 + a->c1 * t2);
    const fq2_t di = d.inverse_binary();
    r->c0 = t0 * di; r->c1 = t1 * di; r->c2 = t2 * di;
}

// ---- Fq12 (fq12.rs: mul :120-135, conjugate :37-39, inverse :137-155, frobenius_map :41-56)
__device__ __noinline__ void f12_mul(fq12_t *r, const fq12_t *a, const fq12_t *b) {
    fq6_t aa, bb, t, u;
    f6_mul(&aa, &a->c0, &b->c0);
    f6_mul(&bb, &a->c1, &b->c1);
    f6_mul(&t, &a->c0, &b->c1);
    f6_mul(&u, &a->c1, &b->c0);
    f6_add(&r->c1, &t, &u);
    f6_mul_by_v(&bb, &bb);
    f6_add(&r->c0, &aa, &bb);
}
__device__ __noinline__ void f12_conj(fq12_t *r, const fq12_t *a) { r->c0 = a->c0; f6_neg(&r->c1, &a->c1); }
__device__ __noinline__ void f12_inv(fq12_t *r, const fq12_t *a) {
    fq6_t t0, t1;
    f6_mul(&t0, &a->c0, &a->c0);
    f6_mul(&t1, &a->c1, &a->c1);
    f6_mul_by_v(&t1, &t1);
    f6_sub(&t0, &t0, &t1);
    f6_inv(&t1, &t0);
    f6_mul(&r->c0, &a->c0, &t1);
    f6_mul(&t0, &a->c1, &t1);
    f6_neg(&r->c1, &t0);
}
__device__ void f12_one(fq12_t *r) {
    const fq2_t z = fq2_t::zero();
    r->c0 = {fq2_t::one(), z, z};
    r->c1 = {z, z, z};
}
__device__ bool f12_eq(const fq12_t *a, const fq12_t *b) {
    return a->c0.c0 == b->c0.c0 && a->c0.c1 == b->c0.c1 && a->c0.c2 == b->c0.c2 && a->c1.c0 == b->c1.c0 && a->c1.c1 == b->c1.c1 && a->c1.c2 == b->c1.c2;
}
// a^(q^k), k = 1..3: coefficient of w^m (m = 2 i + j for c_j.c_i) -> conj^k(coefficient) * xi^(m (q^k - 1) / 6)
__device__ __noinline__ void f12_frobenius(fq12_t *r, const fq12_t *a, int k) {
    r->c0.c0 = fq2_frob(a->c0.c0, k);
    r->c0.c1 = fq2_frob(a->c0.c1, k) * frob_const(k, 2);
    r->c0.c2 = fq2_frob(a->c0.c2, k) * frob_const(k, 4);
    r->c1.c0 = fq2_frob(a->c1.c0, k) * frob_const(k, 1);
    r->c1.c1 = fq2_frob(a->c1.c1, k) * frob_const(k, 3);
    r->c1.c2 = fq2_frob(a->c1.c2, k) * frob_const(k, 5);
}
// f^x for the 64-bit curve parameter, then conjugated because the parameter is negative (mod.rs:118-125 exp_by_x)
__device__ __noinline__ void f12_exp_by_x(fq12_t *r, const fq12_t *f, uint64_t x) {
    fq12_t acc, base = *f;
    f12_one(&acc);
    bool found = false;
    for (int i = 63; i >= 0; i--) {
        if (found) f12_mul(&acc, &acc, &acc);
        if ((x >> i) & 1) {
            found = true;
            f12_mul(&acc, &acc, &base);
        }
    }
    f12_conj(r, &acc);
}

// ---- Miller loop (mod.rs:40-101), affine on the twist; P, Q affine and not the identity
__device__ void line_eval(fq12_t *l, const fq2_t &lam, const fq2_t &xt, const fq2_t &yt, const fq_t &xp, const fq_t &yp) {
    const fq2_t z = fq2_t::zero();
    l->c0.c0 = lam * xt - yt;
    l->c0.c1 = fq2_t{lam.c0 * xp, lam.c1 * xp}.neg();
    l->c0.c2 = z;
    l->c1.c0 = z;
    l->c1.c1 = {yp, fq_t::zero()};
    l->c1.c2 = z;
}
__device__ __noinline__ void miller_loop(fq12_t *f, const g1_affine_t *p, bool p_inf, const g2_affine_t *q, bool q_inf) {
    f12_one(f);
    if (p_inf || q_inf) return;  // mod.rs:49-56: pairs with an identity contribute 1
    const fq_t xp = p->x, yp = p->y;
    const fq2_t xq = q->x, yq = q->y;
    fq2_t xt = xq, yt = yq;
    fq12_t l;
    for (int i = 62; i >= 0; i--) {  // from the bit below the top one of BLS_X
        fq2_t xx = xt.sqr();
        fq2_t lam = (xx.dbl() + xx) * yt.dbl().inverse_binary();
        line_eval(&l, lam, xt, yt, xp, yp);
        f12_mul(f, f, f);
        f12_mul(f, f, &l);
        fq2_t x3 = lam.sqr() - xt.dbl();
        yt = lam * (xt - x3) - yt;
        xt = x3;
        if ((BLS_X >> i) & 1) {
            lam = (yq - yt) * (xq - xt).inverse_binary();
            line_eval(&l, lam, xt, yt, xp, yp);
            f12_mul(f, f, &l);
            x3 = lam.sqr() - xt - xq;
            yt = lam * (xt - x3) - yt;
            xt = x3;
        }
    }
    f12_conj(f, f);  // the parameter is negative
}
// mod.rs:103-160: the easy part f^((q^6 - 1)(q^2 + 1)), then the reference's chain for the hard part
__device__ __noinline__ void final_exponentiation(fq12_t *out, const fq12_t *in) {
    fq12_t r, f1, f2, y0, y1, y2, y3;
    f12_conj(&f1, in);
    f12_inv(&f2, in);
    f12_mul(&r, &f1, &f2);
    f2 = r;
    f12_frobenius(&r, &r, 2);
    f12_mul(&r, &r, &f2);
    f12_mul(&y0, &r, &r);
    f12_exp_by_x(&y1, &y0, BLS_X);
    f12_exp_by_x(&y2, &y1, BLS_X >> 1);
    f12_conj(&y3, &r);
    f12_mul(&y1, &y1, &y3);
    f12_conj(&y1, &y1);
    f12_mul(&y1, &y1, &y2);
    f12_exp_by_x(&y2, &y1, BLS_X);
    f12_exp_by_x(&y3, &y2, BLS_X);
    f12_conj(&y1, &y1);
    f12_mul(&y3, &y3, &y1);
    f12_conj(&y1, &y1);
    f12_frobenius(&y1, &y1, 3);
    f12_frobenius(&y2, &y2, 2);
    f12_mul(&y1, &y1, &y2);
    f12_exp_by_x(&y2, &y3, BLS_X);
    f12_mul(&y2, &y2, &y0);
    f12_mul(&y2, &y2, &r);
    f12_mul(&y1, &y1, &y2);
    f12_frobenius(&y2, &y3, 1);
    f12_mul(out, &y1, &y2);
}

// Engine::pairing (pairing/src/lib.rs:86-96) for n independent pairs: one thread each (the parity kernel)
__global__ void __launch_bounds__(32) k_pairing(const g1_affine_t *p, const uint8_t *p_inf, const g2_affine_t *q, const uint8_t *q_inf, size_t n, fq12_t *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq12_t f;
    miller_loop(&f, p + i, p_inf && p_inf[i], q + i, q_inf && q_inf[i]);
    final_exponentiation(out + i, &f);
}

// verify_proof (verifier.rs:35-66) for a batch: block = one proof.
//   acc = ic[0] + sum_i input_i * ic[i + 1]                                    (threads share the scalar multiplications)
//   FE( ML(A, B) * ML(acc, -gamma) * ML(C, -delta) ) == e(alpha, beta)         (three Miller loops on three threads)
struct VerifyArgs {
    const g1_affine_t *ic;        // n_inputs + 1
    size_t n_inputs;
    const g2_affine_t *neg_gamma_delta;  // [0] = -gamma_g2, [1] = -delta_g2
    const fq12_t *alpha_beta;     // e(alpha_g1, beta_g2)
    const g1_affine_t *proof_a, *proof_c;  // n proofs
    const g2_affine_t *proof_b;
    const uint8_t *proof_inf;     // 3 flags per proof (a, b, c at infinity)
    const uint32_t *inputs;       // n x n_inputs canonical FrRepr (8 words)
    uint8_t *ok;                  // n results
};
__global__ void __launch_bounds__(32) k_verify(VerifyArgs g) {
    __shared__ g1_xyzz_t terms[32];
    __shared__ fq12_t ml[3];
    __shared__ g1_affine_t acc_aff;
    __shared__ bool acc_inf;
    const size_t k = blockIdx.x;
    const uint32_t t = threadIdx.x;
    // input_i * ic[i + 1] by double-and-add (ec.rs:87-99), inputs strided over the threads
    g1_xyzz_t mine = g1_xyzz_t::zero();
    for (size_t i = t; i < g.n_inputs; i += 32) {
        const uint32_t *s = g.inputs + 8 * (k * g.n_inputs + i);
        g1_xyzz_t prod = g1_xyzz_t::zero();
        bool found = false;
        for (int b = 255; b >= 0; b--) {
            const bool bit = (s[b >> 5] >> (b & 31)) & 1;
            if (found) prod.dbl(); else found = bit;
            if (bit) prod.add_mixed(g.ic[i + 1], false);
        }
        mine.add(prod);
    }
    terms[t] = mine;
    __syncthreads();
    if (t == 0) {
        g1_xyzz_t acc = g1_xyzz_t::from_affine(g.ic[0]);
        for (int i = 0; i < 32; i++) acc.add(terms[i]);
        acc_inf = !jacobian_to_affine_serial(acc.to_jacobian(), acc_aff);
    }
    __syncthreads();
    if (t < 3) {
        const uint8_t *inf = g.proof_inf + 3 * k;
        if (t == 0) miller_loop(&ml[0], g.proof_a + k, inf[0] != 0, g.proof_b + k, inf[1] != 0);
        else if (t == 1) miller_loop(&ml[1], &acc_aff, acc_inf, g.neg_gamma_delta + 0, false);
        else miller_loop(&ml[2], g.proof_c + k, inf[2] != 0, g.neg_gamma_delta + 1, false);
    }
    __syncthreads();
    if (t == 0) {
        fq12_t f, e;
        f12_mul(&f, &ml[0], &ml[1]);
        f12_mul(&f, &f, &ml[2]);
        final_exponentiation(&e, &f);
        g.ok[k] = f12_eq(&e, g.alpha_beta) ? 1 : 0;
    }
}
__global__ void k_negate_g2(const g2_affine_t *in, g2_affine_t *out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = {in[i].x, in[i].y.neg()};
}

}  // namespace b200zk

using namespace b200zk;

// prepare_verifying_key (verifier.rs:18-33): e(alpha, beta), -gamma, -delta and ic, resident on the device
struct b200zk_pvk {
    b200zk_ctx *ctx;
    size_t n_ic;
    void *ic = nullptr, *neg_gd = nullptr, *alpha_beta = nullptr;
};

extern "C" {

int b200zk_pairing(b200zk_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint64_t *out_fq12) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (n && (!g1_xy || !g2_xy || !out_fq12)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (n == 0) return B200ZK_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_q = al(n * 96), o_pi = o_q + al(n * 192), o_qi = o_pi + al(n), o_out = o_qi + al(n), total = o_out + n * sizeof(fq12_t);
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, total);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, g1_xy, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_q, g2_xy, n * 192, cudaMemcpyHostToDevice, ctx->stream));
    if (g1_inf) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_pi, g1_inf, n, cudaMemcpyHostToDevice, ctx->stream));
    if (g2_inf) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_qi, g2_inf, n, cudaMemcpyHostToDevice, ctx->stream));
    k_pairing<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>((const g1_affine_t *)s, g1_inf ? (const uint8_t *)(s + o_pi) : nullptr, (const g2_affine_t *)(s + o_q),
                                                                g2_inf ? (const uint8_t *)(s + o_qi) : nullptr, n, (fq12_t *)(s + o_out));
    ctx->launches++;
    B200ZK_CUDA(ctx, cudaGetLastError());
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_fq12, s + o_out, n * sizeof(fq12_t), cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_prepare_verifying_key(b200zk_ctx *ctx, const uint64_t alpha_g1[12], const uint64_t beta_g2[24], const uint64_t gamma_g2[24], const uint64_t delta_g2[24],
                                 const uint64_t *ic, size_t n_ic, b200zk_pvk **out) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (!alpha_g1 || !beta_g2 || !gamma_g2 || !delta_g2 || !ic || n_ic == 0 || !out) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument / empty ic");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_pvk *p = new b200zk_pvk();
    p->ctx = ctx;
    p->n_ic = n_ic;
    uint64_t ab[72];
    int rc = b200zk_pairing(ctx, alpha_g1, nullptr, beta_g2, nullptr, 1, ab);  // alpha_g1_beta_g2 (verifier.rs:27)
    if (!rc && (cudaMalloc(&p->ic, n_ic * 96) != cudaSuccess || cudaMalloc(&p->neg_gd, 2 * 192) != cudaSuccess || cudaMalloc(&p->alpha_beta, sizeof(fq12_t)) != cudaSuccess))
        rc = set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(pvk) failed");
    if (!rc) {
        uint64_t gd[48];
        memcpy(gd, gamma_g2, 192);
        memcpy(gd + 24, delta_g2, 192);
        rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, 512);
        if (!rc) {
            cudaMemcpyAsync(ctx->scratch, gd, 384, cudaMemcpyHostToDevice, ctx->stream);
            k_negate_g2<<<1, 32, 0, ctx->stream>>>((const g2_affine_t *)ctx->scratch, (g2_affine_t *)p->neg_gd, 2);  // verifier.rs:28-31
            cudaMemcpyAsync(p->ic, ic, n_ic * 96, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(p->alpha_beta, ab, sizeof(fq12_t), cudaMemcpyHostToDevice, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, B200ZK_ERR_CUDA, "pvk upload failed");
        }
    }
    if (rc) { b200zk_pvk_free(p); return rc; }
    *out = p;
    return B200ZK_OK;
}

void b200zk_pvk_free(b200zk_pvk *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaFree(p->ic);
    cudaFree(p->neg_gd);
    cudaFree(p->alpha_beta);
    delete p;
}

int b200zk_verify_proofs(b200zk_ctx *ctx, const b200zk_pvk *pvk, const uint64_t *proofs_a, const uint64_t *proofs_b, const uint64_t *proofs_c, const uint8_t *inf_flags,
                         const uint64_t *public_inputs, size_t n_inputs, size_t n_proofs, uint8_t *ok) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (!pvk || (n_proofs && (!proofs_a || !proofs_b || !proofs_c || !ok)) || (n_proofs && n_inputs && !public_inputs)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (n_inputs + 1 != pvk->n_ic) return set_error(ctx, B200ZK_ERR_BAD_ARG, "MalformedVerifyingKey: public_inputs.len() + 1 != ic.len() (verifier.rs:41-43)");
    if (pvk->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "the prepared key lives on another device");
    if (n_proofs == 0) return B200ZK_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t n = n_proofs;
    const size_t o_b = al(n * 96), o_c = o_b + al(n * 192), o_f = o_c + al(n * 96), o_in = o_f + al(3 * n), o_ok = o_in + al(n * n_inputs * 32 + 1), total = o_ok + al(n);
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, total);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, proofs_a, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_b, proofs_b, n * 192, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_c, proofs_c, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    if (inf_flags) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_f, inf_flags, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    else B200ZK_CUDA(ctx, cudaMemsetAsync(s + o_f, 0, 3 * n, ctx->stream));
    if (n_inputs) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_in, public_inputs, n * n_inputs * 32, cudaMemcpyHostToDevice, ctx->stream));
    VerifyArgs g{(const g1_affine_t *)pvk->ic, n_inputs, (const g2_affine_t *)pvk->neg_gd, (const fq12_t *)pvk->alpha_beta, (const g1_affine_t *)s,
                 (const g1_affine_t *)(s + o_c), (const g2_affine_t *)(s + o_b), (const uint8_t *)(s + o_f), (const uint32_t *)(s + o_in), (uint8_t *)(s + o_ok)};
    k_verify<<<(unsigned)n, 32, 0, ctx->stream>>>(g);
    ctx->launches++;
    B200ZK_CUDA(ctx, cudaGetLastError());
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ok, s + o_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

}  // extern "C"
