// The numeric pipeline of groth16::create_proof (bellman/src/groth16/prover.rs:249-364) on the device:
// H polynomial (7 NTTs + pointwise, prover.rs:256-287) -> the multiexps of prover.rs:289-318 -> proof assembly
// (prover.rs:326-363).  Circuit synthesis (prover.rs:212-234) stays on the host: the caller passes the a/b/c
// evaluation vectors, the input / aux assignments and the three density maps that ProvingAssignment collects.
//
// Assembly identities (same group elements as prover.rs:326-354, fewer scalar multiplications):
//   g_a = r*delta_g1 + alpha_g1 + (a_inputs + a_aux)
//   g_b = s*delta_g2 + beta_g2  + (b2_inputs + b2_aux)
//   g_c = s*g_a + r*(beta_g1 + b1_inputs + b1_aux) + h + l
//        [ = rs*delta_g1 + s*alpha_g1 + r*beta_g1 + s*a_answer + r*b1_answer + h + l ]
// r*delta_g1 and s*delta_g2 use per-CRS window tables (32 windows x 255 multiples, built once at upload, summed by a
// warp tree); the two variable-base products are MSB-first double-and-add with the reference's Jacobian formulas.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ec.cuh"
#include "internal.h"

namespace b200zk {

// The reference issues a_inputs / a_aux (and the B pairs) as separate multiexps over the *same* base vector with the aux
// cursor starting right after the inputs (groth16/mod.rs:456-481); only their sums are used (prover.rs:339-347), so each
// pair is one multiexp over inputs ++ aux with the concatenated density map.

// All assembly kernels run on lane PAIRS (ec.cuh PairXYZZ: lane 0 holds (X, ZZ), lane 1 holds (Y, ZZZ)): they are chains of
// dependent point operations in a handful of threads, i.e. pure latency, and the pair halves it.
// scal[16 k ..] = r_k (8 words), s_k (8 words), canonical FrRepr.

// sum of the 32 pairs' points of a 64-thread block (shared-memory tree); the result is valid in pair 0
template <class F>
__device__ PairXYZZ<F> pair_block_sum32(PairXYZZ<F> acc, XYZZ<F> *sm) {
    const uint32_t j = threadIdx.x >> 1;
    acc.store(&sm[j]);
    __syncthreads();
    for (uint32_t stride = 16; stride > 0; stride >>= 1) {
        const bool on = j < stride;
        PairXYZZ<F> x = on ? PairXYZZ<F>::load(&sm[j]) : PairXYZZ<F>::zero();
        const PairXYZZ<F> y = on ? PairXYZZ<F>::load(&sm[j + stride]) : PairXYZZ<F>::zero();
        x.add(y);
        __syncthreads();
        if (on) x.store(&sm[j]);
        __syncthreads();
    }
    return PairXYZZ<F>::load(&sm[0]);
}
// T = scalar * (the point whose 8-bit window table is `table`: 32 windows x 255 multiples, built once per CRS): pair j looks up
// window j, the block adds the 32 entries
template <class F>
__device__ PairXYZZ<F> pair_table_mul(const XYZZ<F> *table, const uint32_t *scalar, XYZZ<F> *sm) {
    const uint32_t j = threadIdx.x >> 1;  // 64 threads, one 8-bit window per pair
    const uint32_t d = (scalar[j >> 2] >> (8 * (j & 3))) & 0xff;
    const PairXYZZ<F> e = d ? PairXYZZ<F>::load(table + (size_t)j * 255 + d - 1) : PairXYZZ<F>::zero();
    return pair_block_sum32<F>(e, sm);
}
// Jacobian (X, Y, Z) -> the pair's halves of (X, Y, Z^2, Z^3)
template <class F>
__device__ PairXYZZ<F> pair_from_jacobian(const Jacobian<F> *p, bool mine) {
    const bool r1 = PairXYZZ<F>::role();
    const F z = mine ? p->z : F::zero();
    const F zz = z.sqr();
    const F zzz = zz * z;
    PairXYZZ<F> out;
    out.a = mine ? (r1 ? p->y : p->x) : F::zero();
    out.b = r1 ? zzz : zz;
    return out;
}
// affine (x, y) = (X / ZZ, Y / ZZZ): each lane inverts its own denominator (the reference's binary Euclid, fq.rs:849-903:
// data-dependent loops, but no shuffle inside) -- lane 0 returns x, lane 1 returns y
template <class F>
__device__ F pair_to_affine_half(const PairXYZZ<F> &p) { return p.a * p.b.inverse_binary(); }

// CurveProjective::mul_assign (ec.rs:528-552) on a lane pair: 4-bit fixed windows, MSB first.  `tab` (shared memory, 15 entries)
// receives 1 P .. 15 P.  The scalar is uniform over the block, so the control flow is too.
template <class F>
__device__ PairXYZZ<F> pair_scalar_mul(const PairXYZZ<F> &p, const uint32_t *k, XYZZ<F> *tab) {
    const bool mine = threadIdx.x < 2;
    PairXYZZ<F> cur = p;
    if (mine) cur.store(&tab[0]);
    for (int i = 1; i < 15; i++) {  // i P -> (i + 1) P
        if (i & 1) { PairXYZZ<F> h = mine ? PairXYZZ<F>::load(&tab[(i - 1) / 2]) : PairXYZZ<F>::zero(); h.dbl(); cur = h; }  // 2 m P = dbl(m P)
        else cur.add(p);
        if (mine) cur.store(&tab[i]);
    }
    PairXYZZ<F> acc = PairXYZZ<F>::zero();
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { acc.dbl(); acc.dbl(); acc.dbl(); acc.dbl(); }
        const uint32_t d = (k[w >> 3] >> (4 * (w & 7))) & 0xf;
        if (d) {
            const PairXYZZ<F> e = mine ? PairXYZZ<F>::load(&tab[d - 1]) : PairXYZZ<F>::zero();  // each lane reads back the halves it wrote itself
            acc.add(e);
        }
    }
    return acc;
}

// The assembly is cut along the multiexps it consumes, so each piece runs on the lane of its multiexp as soon as that one
// is done:
//   k_proof_a  (after A):    g_a = r*delta_g1 (table) + alpha_g1 + A;  proof.a = affine(g_a);  sga = s * g_a
//   k_proof_b1 (after B-G1): rb1 = r * (beta_g1 + B1)
//   k_proof_b  (after B-G2): proof.b = affine(s*delta_g2 (table) + beta_g2 + B2)
//   k_proof_c  (after all):  proof.c = affine(sga + rb1 + H + L)
// One 64-thread block per proof of the batch (blockIdx.x = k); every result array is indexed by k.
__global__ void __launch_bounds__(64) k_proof_a(const g1_xyzz_t *table_d1, const uint32_t *scal, const g1_affine_t *vk_g1 /* alpha, beta, delta */,
                                               bool alpha_inf, const g1_jac_t *res_a, g1_xyzz_t *sga, g1_affine_t *proof_a, uint8_t *inf_flags) {
    __shared__ g1_xyzz_t sm[32];
    const uint32_t k = blockIdx.x;
    const bool r1 = PairXYZZ<fq_t>::role();
    scal += 16 * k;
    PairXYZZ<fq_t> t = pair_table_mul<fq_t>(table_d1, scal, sm);
    if (threadIdx.x >= 32) return;  // one warp goes on (its first pair does the work, the others run along on the identity)
    const bool mine = threadIdx.x < 2;
    if (!mine) t = PairXYZZ<fq_t>::zero();
    const fq_t alpha_half = r1 ? vk_g1[0].y : vk_g1[0].x;
    t.add_mixed(alpha_half, mine && !alpha_inf);  // add_assign_mixed skips an identity operand (ec.rs:447-449)
    t.add(pair_from_jacobian<fq_t>(res_a + k, mine));
    const bool zero = t.is_zero();
    const fq_t h = pair_to_affine_half<fq_t>(t);
    if (threadIdx.x == 0) { proof_a[k].x = zero ? fq_t::zero() : h; inf_flags[4 * k + 0] = zero ? 1 : 0; }
    if (threadIdx.x == 1) proof_a[k].y = zero ? fq_t::one() : h;
    __syncwarp();
    const PairXYZZ<fq_t> prod = pair_scalar_mul<fq_t>(t, scal + 8, sm);
    if (mine) prod.store(&sga[k]);
}

__global__ void __launch_bounds__(32) k_proof_b1(const uint32_t *scal, const g1_affine_t *vk_g1, bool beta_inf, const g1_jac_t *res_b1, g1_xyzz_t *rb1) {
    __shared__ g1_xyzz_t sm[16];
    const uint32_t k = blockIdx.x;
    const bool r1 = PairXYZZ<fq_t>::role(), mine = threadIdx.x < 2;
    PairXYZZ<fq_t> b1 = PairXYZZ<fq_t>::zero();
    const fq_t beta_half = r1 ? vk_g1[1].y : vk_g1[1].x;
    b1.add_mixed(beta_half, mine && !beta_inf);
    b1.add(pair_from_jacobian<fq_t>(res_b1 + k, mine));
    const PairXYZZ<fq_t> prod = pair_scalar_mul<fq_t>(b1, scal + 16 * k, sm);
    if (mine) prod.store(&rb1[k]);
}

__global__ void __launch_bounds__(64) k_proof_b(const g2_xyzz_t *table_d2, const uint32_t *scal, const g2_affine_t *vk_g2 /* beta, delta */,
                                               bool beta_inf, const g2_jac_t *res_b2, g2_affine_t *proof_b, uint8_t *inf_flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    g2_xyzz_t *sm = reinterpret_cast<g2_xyzz_t *>(smem_raw);
    const uint32_t k = blockIdx.x;
    const bool r1 = PairXYZZ<fq2_t>::role();
    PairXYZZ<fq2_t> t = pair_table_mul<fq2_t>(table_d2, scal + 16 * k + 8, sm);
    if (threadIdx.x >= 32) return;
    const bool mine = threadIdx.x < 2;
    if (!mine) t = PairXYZZ<fq2_t>::zero();
    const fq2_t beta_half = r1 ? vk_g2[0].y : vk_g2[0].x;
    t.add_mixed(beta_half, mine && !beta_inf);
    t.add(pair_from_jacobian<fq2_t>(res_b2 + k, mine));
    const bool zero = t.is_zero();
    const fq2_t h = pair_to_affine_half<fq2_t>(t);
    if (threadIdx.x == 0) { proof_b[k].x = zero ? fq2_t::zero() : h; inf_flags[4 * k + 1] = zero ? 1 : 0; }
    if (threadIdx.x == 1) proof_b[k].y = zero ? fq2_t::one() : h;
}

__global__ void __launch_bounds__(32) k_proof_c(const g1_xyzz_t *sga, const g1_xyzz_t *rb1, const g1_jac_t *res_h, const g1_jac_t *res_l,
                                               g1_affine_t *proof_c, uint8_t *inf_flags) {
    const uint32_t k = blockIdx.x;
    const bool mine = threadIdx.x < 2;
    PairXYZZ<fq_t> c = mine ? PairXYZZ<fq_t>::load(&sga[k]) : PairXYZZ<fq_t>::zero();
    c.add(mine ? PairXYZZ<fq_t>::load(&rb1[k]) : PairXYZZ<fq_t>::zero());
    c.add(pair_from_jacobian<fq_t>(res_h + k, mine));
    c.add(pair_from_jacobian<fq_t>(res_l + k, mine));
    const bool zero = c.is_zero();
    const fq_t h = pair_to_affine_half<fq_t>(c);
    if (threadIdx.x == 0) { proof_c[k].x = zero ? fq_t::zero() : h; inf_flags[4 * k + 2] = zero ? 1 : 0; }
    if (threadIdx.x == 1) proof_c[k].y = zero ? fq_t::one() : h;
}

static inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

// B200ZK_PROVE_TRACE=1: device timeline of one create_proof (events on the main stream and the lanes), printed to stderr
struct ProveTrace {
    bool on = false;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    ProveTrace() { if (const char *e = getenv("B200ZK_PROVE_TRACE")) on = e[0] == '1'; }
    void mark(const char *name, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        marks.emplace_back(name, ev);
    }
    void report() {
        if (!on || marks.empty()) return;
        for (auto &m : marks) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[0].second, m.second);
            fprintf(stderr, "[prove] %-22s %8.3f ms\n", m.first, ms);
        }
        for (auto &m : marks) cudaEventDestroy(m.second);
    }
};

// K proofs over one CRS in lock-step (all of the same circuit: equal n_constraints / n_inputs / n_aux).  The five multiexps
// of the K proofs run as five batched multiexps (K bucket sets each, msm_impl.cuh) so that a batch fills the machine where a
// single 10^5-point multiexp cannot; K = 1 is groth16::create_proof as is.
int groth16_prove_batch(Ctx *ctx, const Crs *crs, const ProveArgs *args, uint32_t K, uint64_t *proof_a, uint64_t *proof_b, uint64_t *proof_c,
                        uint8_t *inf_flags) {
    cudaStream_t st = ctx->stream;
    if (crs->subverted) return set_error(ctx, B200ZK_ERR_UNEXPECTED_IDENTITY, "delta_g1 / delta_g2 is the identity (subversion check, prover.rs:320-324)");
    if (K == 0 || K > 256) return set_error(ctx, B200ZK_ERR_BAD_ARG, "batch must be in [1, 256]");
    const ProveArgs &g = args[0];
    for (uint32_t k = 1; k < K; k++)
        if (args[k].n_constraints != g.n_constraints || args[k].n_inputs != g.n_inputs || args[k].n_aux != g.n_aux)
            return set_error(ctx, B200ZK_ERR_BAD_ARG, "the proofs of a batch must come from the same circuit (equal sizes)");
    // EvaluationDomain::from_coeffs (domain.rs:48-81)
    size_t m = 1;
    uint32_t log_m = 0;
    while (m < g.n_constraints) {
        m *= 2;
        log_m++;
        if (log_m >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "PolynomialDegreeTooLarge");
    }
    const size_t vec = m * 32;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    // every per-proof array is K consecutive copies (stride = the size of one)
    size_t o_a = take(K * vec), o_b = take(K * vec), o_c = take(K * vec), o_h = take(K * vec);
    const size_t n_all = g.n_inputs + g.n_aux;
    size_t o_in = take(K * n_all * 32);                      // inputs ++ aux, contiguous
    size_t o_da = take(K * n_all), o_db = take(K * n_all);   // [1..1] ++ a_aux_density ; b_input_density ++ b_aux_density
    size_t o_scal = take(K * 64), o_st = take(5 * K * 4);
    size_t o_rh = take(K * 144), o_rl = take(K * 144), o_ra = take(K * 144), o_rb1 = take(K * 144), o_rb2 = take(K * 288);
    size_t o_sga = take(K * 192), o_rb1r = take(K * 192), o_pa = take(K * 96), o_pb = take(K * 192), o_pc = take(K * 96), o_inf = take(K * 4);
    int rc = ensure_scratch(ctx, &ctx->scratch3, &ctx->scratch3_bytes, off);
    if (rc) return rc;
    char *w = (char *)ctx->scratch3;
    // ---- inputs to HBM (zero padding of a, b, c up to m as from_coeffs does)
    ProveTrace trace;
    trace.mark("start", st);
    auto up = [&](size_t o, const void *src, size_t bytes) { return bytes ? cudaMemcpyAsync(w + o, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    const size_t used = g.n_constraints * 32;
    // all fills first: once the lanes run, a fill kernel on this stream would queue behind their multiexps and hold up the
    // copies that follow it (a batch of 8 spent 6 ms uploading 76 MB that way)
    for (uint32_t k = 0; k < K && vec > used; k++) {
        B200ZK_CUDA(ctx, cudaMemsetAsync(w + o_a + k * vec + used, 0, vec - used, st));
        B200ZK_CUDA(ctx, cudaMemsetAsync(w + o_b + k * vec + used, 0, vec - used, st));
        B200ZK_CUDA(ctx, cudaMemsetAsync(w + o_c + k * vec + used, 0, vec - used, st));
    }
    for (uint32_t k = 0; k < K; k++) {
        const ProveArgs &p = args[k];
        const size_t ok = k * n_all;
        B200ZK_CUDA(ctx, up(o_in + ok * 32, p.inputs, p.n_inputs * 32));
        B200ZK_CUDA(ctx, up(o_in + (ok + p.n_inputs) * 32, p.aux, p.n_aux * 32));
        if (p.n_inputs) B200ZK_CUDA(ctx, cudaMemsetAsync(w + o_da + ok, 1, p.n_inputs, st));  // inputs have full density in A (prover.rs:148-151)
        B200ZK_CUDA(ctx, up(o_da + ok + p.n_inputs, p.a_aux_density, p.n_aux));
        B200ZK_CUDA(ctx, up(o_db + ok, p.b_input_density, p.n_inputs));
        B200ZK_CUDA(ctx, up(o_db + ok + p.n_inputs, p.b_aux_density, p.n_aux));
        B200ZK_CUDA(ctx, up(o_scal + k * 64, p.r, 32));
        B200ZK_CUDA(ctx, up(o_scal + k * 64 + 32, p.s, 32));
    }
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));  // the assignments and densities are in HBM: lanes may start
    trace.mark("assignment uploaded", st);
    for (uint32_t k = 0; k < K; k++) {
        const ProveArgs &p = args[k];
        B200ZK_CUDA(ctx, up(o_a + k * vec, p.a, used));
        B200ZK_CUDA(ctx, up(o_b + k * vec, p.b, used));
        B200ZK_CUDA(ctx, up(o_c + k * vec, p.c, used));
    }
    trace.mark("a, b, c uploaded", st);
    // ---- H polynomials (prover.rs:256-287)
    if (o_b == o_a + K * vec && o_c == o_b + K * vec) {  // (always: the three blocks are carved back to back) all K proofs per launch
        // (measured, no gain: the H block on a highest-priority stream of its own -- its launches wait for resident multiexp
        // blocks to retire, not for pending ones)
        if ((rc = ntt_h_poly_batch(ctx, w + o_a, log_m, w + o_h, K))) return rc;
    } else {
        for (uint32_t k = 0; k < K; k++)
            if ((rc = ntt_h_poly(ctx, w + o_a + k * vec, w + o_b + k * vec, w + o_c + k * vec, log_m, w + o_h + k * vec))) return rc;
    }
    trace.mark("h polynomial", st);
    // ---- the multiexps (prover.rs:289-318), each over the whole batch
    uint32_t *stw = (uint32_t *)(w + o_st);
    const uint8_t *da = (const uint8_t *)(w + o_da), *db = (const uint8_t *)(w + o_db);
    struct Job { const Bases *b; size_t src; size_t n; const uint8_t *d; size_t out; MsmBatch batch; };
    const int n_jobs = 5;
    Job jobs[n_jobs] = {
        {crs->h, o_h, m - 1, nullptr, o_rh, {K, m, 0}},
        {crs->l, o_in + g.n_inputs * 32, g.n_aux, nullptr, o_rl, {K, n_all, 0}},
        {crs->a, o_in, n_all, da, o_ra, {K, n_all, n_all}},
        {crs->b_g1, o_in, n_all, db, o_rb1, {K, n_all, n_all}},
        {crs->b_g2, o_in, n_all, db, o_rb2, {K, n_all, n_all}},
    };
    // The H multiexp needs the H polynomials; the other four only need the assignments, so they run beside it on the context's
    // lanes (own stream and workspaces each) and join before the last piece.  B200ZK_PROVE_LANES=0 keeps everything on one stream.
    int n_lanes = n_jobs - 1;
    if (const char *e = getenv("B200ZK_PROVE_LANES")) n_lanes = std::max(0, std::min(n_jobs - 1, atoi(e)));
    if (n_lanes && (rc = ctx_lanes(ctx, n_lanes))) return rc;
    const g1_affine_t *vk1 = (const g1_affine_t *)crs->vk;
    const g2_affine_t *vk2 = (const g2_affine_t *)((const char *)crs->vk + 3 * 96);
    g1_xyzz_t *sga = (g1_xyzz_t *)(w + o_sga), *rb1 = (g1_xyzz_t *)(w + o_rb1r);
    uint8_t *dinf = (uint8_t *)(w + o_inf);
    const uint32_t *scal = (const uint32_t *)(w + o_scal);
    // enqueue order = how early a lane gets its kernels: A first (a 255-bit scalar multiplication follows it), then the G2
    // multiexp (the longest chain), B-G1 (another scalar multiplication), L, and H last (it waits for the H block anyway)
    static const int order[n_jobs] = {2, 4, 3, 1, 0};
    for (int oi = 0; oi < n_jobs; oi++) {
        const int j = order[oi];
        // job 0 (H) stays on the main stream; job j >= 1 goes to lane (j - 1) mod n_lanes
        Ctx *on = (j == 0 || n_lanes == 0) ? ctx : ctx->lanes[(j - 1) % n_lanes];
        std::unique_lock<std::recursive_mutex> lk;
        if (on != ctx) {
            lk = std::unique_lock<std::recursive_mutex>(on->mu);
            B200ZK_CUDA(ctx, cudaStreamWaitEvent(on->stream, ctx->ev_fork, 0));
        }
        const unsigned long long before = on->launches;
        if ((rc = msm_run(on, jobs[j].b, 0, w + jobs[j].src, jobs[j].n, jobs[j].d, w + jobs[j].out, stw + (size_t)j * K, 0, jobs[j].batch))) {
            for (Ctx *lane : ctx->lanes) cudaStreamSynchronize(lane->stream);  // other lanes still read this call's workspace
            cudaStreamSynchronize(st);
            return on == ctx ? rc : set_error(ctx, rc, on->last_error);
        }
        static const char *const msm_names[] = {"multiexp H", "multiexp L", "multiexp A", "multiexp B-G1", "multiexp B-G2"};
        static const char *const piece_names[] = {"", "", "piece a (g_a, s*g_a)", "piece b1 (r*B1)", "piece b (g_b)"};
        trace.mark(msm_names[j], on->stream);
        // the piece of the assembly (prover.rs:326-363) that only needs this multiexp
        if (j == 2)
            k_proof_a<<<K, 64, 0, on->stream>>>((const g1_xyzz_t *)crs->table_delta_g1, scal, vk1, crs->vk_inf[0] != 0, (const g1_jac_t *)(w + o_ra), sga, (g1_affine_t *)(w + o_pa), dinf);
        if (j == 3) k_proof_b1<<<K, 32, 0, on->stream>>>(scal, vk1, crs->vk_inf[1] != 0, (const g1_jac_t *)(w + o_rb1), rb1);
        if (j == 4)
            k_proof_b<<<K, 64, 32 * sizeof(g2_xyzz_t), on->stream>>>((const g2_xyzz_t *)crs->table_delta_g2, scal, vk2, crs->vk_inf[2] != 0, (const g2_jac_t *)(w + o_rb2),
                                                                    (g2_affine_t *)(w + o_pb), dinf);
        if (j >= 2) { on->launches++; trace.mark(piece_names[j], on->stream); }
        if (on != ctx) ctx->launches += on->launches - before;
    }
    for (int l = 0; l < n_lanes; l++) {
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->lanes[l]->ev_join, ctx->lanes[l]->stream));
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->lanes[l]->ev_join, 0));
    }
    trace.mark("join", st);
    k_proof_c<<<K, 32, 0, st>>>(sga, rb1, (const g1_jac_t *)(w + o_rh), (const g1_jac_t *)(w + o_rl), (g1_affine_t *)(w + o_pc), dinf);
    ctx->launches++;
    trace.mark("piece c", st);
    B200ZK_CUDA(ctx, cudaGetLastError());
    std::vector<uint32_t> status(5 * (size_t)K);
    std::vector<uint8_t> inf(4 * (size_t)K);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(proof_a, w + o_pa, K * 96, cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(proof_b, w + o_pb, K * 192, cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(proof_c, w + o_pc, K * 96, cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(inf.data(), dinf, inf.size(), cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(status.data(), stw, status.size() * 4, cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
    trace.report();
    for (uint32_t k = 0; k < K; k++) {
        for (int j = 0; j < n_jobs; j++) {
            const uint32_t code = status[(size_t)j * K + k];
            const std::string which = K > 1 ? " (proof " + std::to_string(k) + " of the batch)" : "";
            if (code == B200ZK_ERR_UNEXPECTED_IDENTITY) return set_error(ctx, code, "UnexpectedIdentity in multiexp" + which);
            if (code == B200ZK_ERR_UNEXPECTED_EOF) return set_error(ctx, code, "IoError(UnexpectedEof) in multiexp" + which);
        }
        if (inf_flags) { inf_flags[3 * k] = inf[4 * k]; inf_flags[3 * k + 1] = inf[4 * k + 1]; inf_flags[3 * k + 2] = inf[4 * k + 2]; }
    }
    return B200ZK_OK;
}

int groth16_prove(Ctx *ctx, const Crs *crs, const ProveArgs &g, uint64_t *proof_a, uint64_t *proof_b, uint64_t *proof_c, uint8_t *inf_flags) {
    return groth16_prove_batch(ctx, crs, &g, 1, proof_a, proof_b, proof_c, inf_flags);
}

}  // namespace b200zk
