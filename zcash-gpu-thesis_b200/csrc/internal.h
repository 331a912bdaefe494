// Internal declarations shared by the .cu translation units of libb200zk.so (not part of the C ABI).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200zk.h"

namespace b200zk {

struct Ctx;

// ---- error plumbing: every C-ABI entry returns an int status and records a message in the ctx
int set_error(Ctx *ctx, int code, const std::string &msg);
#define B200ZK_CUDA(ctx, call)                                                                                  \
    do {                                                                                                        \
        cudaError_t e__ = (call);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return b200zk::set_error(ctx, B200ZK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// ---- cached per-domain-size NTT tables (device memory, Montgomery form)
struct NttTables {
    uint32_t log_n = 0;
    void *tw = nullptr;        // omega^k,      k < n/2
    void *tw_inv = nullptr;    // omega^-k,     k < n/2
    void *g_pow = nullptr;     // g^i,          i < n   (coset_fft pre-scale; built on first use)
    void *gi_pow = nullptr;    // g^-i / n,     i < n   (icoset_fft post-scale)
    void *consts = nullptr;    // fr_t[8]: omega, omega_inv, n_inv, g, g_inv, z_inv (1/(g^n - 1)), ...
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;
    std::string last_error;
    std::recursive_mutex mu;  // every C-ABI entry holds it: calls on one context from several host threads are serialised
    std::map<uint32_t, NttTables> ntt_tables;
    // reusable scratch (grown on demand, stream-ordered use only)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *scratch2 = nullptr;
    size_t scratch2_bytes = 0;
    void *scratch3 = nullptr;  // groth16 prover workspace
    size_t scratch3_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // NCCL (loaded lazily, only when a communicator is requested)
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    void *gather_buf = nullptr;  // (world + 1) records of REC_BYTES: [0] = this rank's send record, [1..] = the gathered ones
    void *small_slot = nullptr;  // 256 B staging for scalar constants
    cudaStream_t copy_stream = nullptr;  // H2D of the next job's exponents while the current one computes
    struct JobSlot { void *dev = nullptr; size_t bytes = 0, o_res = 0; void *host_res = nullptr; cudaEvent_t copied = nullptr, done = nullptr; bool busy = false; };
    JobSlot slots[4];
    int window_override = 0;  // b200zk_set_msm_window
    int ntt_large_from = 20;  // log2 size from which a transform uses the radix-4 kernels (k_ntt_pass4) (B200ZK_NTT_LARGE_FROM)
    unsigned long long launches = 0;  // kernels launched by this context (b200zk_launch_count)
    bool prof_on = false;        // bracket the dominant MSM kernel with events
    // lanes: child contexts on the same device (own stream + workspaces) on which one create_proof runs its independent
    // multiexps side by side, as prover.rs:289-318 keeps its futures in flight together; created on first use
    std::vector<Ctx *> lanes;
    bool combine_smem_opt_in[2] = {false, false};  // k_msm_combine_big<G1/G2> allowed its dynamic shared memory on this device
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
};

int ensure_scratch(Ctx *ctx, void **buf, size_t *cur, size_t bytes);
int ctx_lanes(Ctx *ctx, int n);  // make sure ctx->lanes holds n children; B200ZK_OK or an error recorded in ctx

struct Bases {
    Ctx *ctx;
    int group;          // 1 = G1 (96 B/point), 2 = G2 (192 B/point)
    size_t n;
    void *points;       // device, x||y Montgomery
    uint8_t *infinity;  // device, n bytes, or nullptr when no base is the identity
    void *pre = nullptr;  // optional: pre[w * n + i] = 2^(c w) * points[i], w < pre_W (b200zk_bases_precompute)
    uint32_t pre_c = 0, pre_W = 0;
};

// ---- launchers implemented in the .cu files (all enqueue on ctx->stream, no implicit sync)
// field_kernels.cu
int launch_field_vec(Ctx *ctx, int field, int op, const void *a, const void *b, void *out, size_t n);
int launch_fr_scale(Ctx *ctx, void *a, const void *scalar_dev, size_t n);
int launch_fr_spmv(Ctx *ctx, const void *row_ptr, const void *col, const void *val, const void *x, size_t n_rows, void *y);
int launch_point_op(Ctx *ctx, int group, int op, const void *a, const void *b, const uint8_t *b_inf, void *out, size_t n);
int launch_microbench(Ctx *ctx, int kind, int iters, int blocks, int threads, void *out);
// ntt.cu
int ntt_get_tables(Ctx *ctx, uint32_t log_n, NttTables **out);
uint32_t ntt_plan(uint32_t log_n, int large_from, int sm_count, uint32_t batch, uint32_t *B, uint32_t *Q, int *radix4);  // host logic: passes of a transform
int ntt_run(Ctx *ctx, void *d_coeffs, uint32_t log_n, int kind);
int ntt_run_batch(Ctx *ctx, void *d_coeffs, uint32_t log_n, int kind, uint32_t batch, size_t stride);  // vector v at d_coeffs + v * stride elements
int ntt_h_poly_batch(Ctx *ctx, void *d_abc, uint32_t log_n, void *d_out_repr, uint32_t K);            // a_0..a_(K-1) | b_0.. | c_0.. contiguous
int ntt_distribute_powers(Ctx *ctx, void *d_coeffs, size_t n, const void *d_g);
void ntt_free_all_tables(Ctx *ctx);
int ntt_divide_by_z_on_coset(Ctx *ctx, void *d_coeffs, uint32_t log_n);
int ntt_domain_z(Ctx *ctx, const void *d_tau, uint32_t log_n, void *d_out);
int ntt_h_poly(Ctx *ctx, void *d_a, void *d_b, void *d_c, uint32_t log_n, void *d_out_repr);
int ntt_h_poly_front(Ctx *ctx, void *d_v, uint32_t log_n);  // ifft + coset_fft of one of a, b, c (prover.rs:257-265)
int ntt_h_poly_tail(Ctx *ctx, void *d_a, const void *d_b, const void *d_c, uint32_t log_n, void *d_out_repr);  // prover.rs:267-287
// msm.cu
// A batch of K multiexps over the SAME bases (K proofs over one CRS): exponent vector k starts k * scalar_stride exponents
// after the first, its density map k * density_stride bytes after the first; the K multiexps are separate bucket sets of one
// pipeline, so small multiexps fill the machine together.  Results: K consecutive Jacobian points and K consecutive status words.
struct MsmBatch {
    uint32_t K = 1;
    size_t scalar_stride = 0, density_stride = 0;
};
int msm_run(Ctx *ctx, const Bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp, const uint8_t *d_density,
            void *d_out_jac, void *d_status_out, int window_bits, MsmBatch batch = MsmBatch());
int msm_fixed_base(Ctx *ctx, int group, const void *d_base_affine, const void *d_scalars, size_t n, uint32_t scalar_bits, void *d_out_affine,
                   uint8_t *d_out_inf);
int msm_into_affine(Ctx *ctx, int group, const void *d_jac, size_t n, void *d_out_xy, uint8_t *d_out_inf);
// sum of n Jacobian points stored `stride` bytes apart (0 = packed)
int msm_sum_points(Ctx *ctx, int group, const void *d_jac_in, size_t n, void *d_jac_out, size_t stride = 0);
int msm_precompute(Ctx *ctx, Bases *bases, uint32_t c);
int msm_build_table(Ctx *ctx, int group, const void *d_base_affine, void *d_table, uint32_t nwin);  // 255 * nwin XYZZ entries
// codec.cu
int codec_decode_uncompressed(Ctx *ctx, int group, const void *d_bytes, size_t n, int checked, int allow_infinity, void *d_out, uint8_t *d_inf,
                              unsigned long long *d_err);
int codec_encode(Ctx *ctx, int group, const void *d_pts, const uint8_t *d_inf, size_t n, int compressed, void *d_out);
// groth16.cu
struct Crs {
    Ctx *ctx;
    Bases *h, *l, *a, *b_g1, *b_g2;
    void *vk;            // device: alpha_g1 | beta_g1 | delta_g1 (3 x 96 B) | beta_g2 | delta_g2 (2 x 192 B)
    void *table_delta_g1;  // 32 x 255 XYZZ<fq>
    void *table_delta_g2;  // 32 x 255 XYZZ<fq2>
    bool subverted;      // delta_g1 or delta_g2 is the identity (prover.rs:320-324)
    uint8_t vk_inf[5] = {0, 0, 0, 0, 0};  // infinity flags of alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2 (add_assign_mixed skips an identity, ec.rs:447-449)
    // Parameters::read keeps the whole VerifyingKey (groth16/mod.rs:100-126) for Parameters::write and for the verifier:
    bool owns_bases = false;           // the five query vectors were created by b200zk_parameters_read and are freed with the CRS
    bool has_full_vk = false;
    uint64_t vk_host[108] = {};        // alpha_g1 | beta_g1 | beta_g2 | gamma_g2 | delta_g1 | delta_g2 (affine x||y, Montgomery)
    uint8_t vk_host_inf[6] = {};
    std::vector<uint64_t> ic;          // 12 words per element
};
struct ProveArgs {
    const uint64_t *a, *b, *c; size_t n_constraints;
    const uint64_t *inputs; size_t n_inputs;
    const uint64_t *aux; size_t n_aux;
    const uint8_t *a_aux_density, *b_input_density, *b_aux_density;
    const uint64_t *r, *s;
};
int groth16_prove(Ctx *ctx, const Crs *crs, const ProveArgs &args, uint64_t *proof_a, uint64_t *proof_b, uint64_t *proof_c, uint8_t *inf_flags);
// K proofs of one circuit in lock-step; outputs are K consecutive proofs (12 / 24 / 12 words, 3 flags each)
int groth16_prove_batch(Ctx *ctx, const Crs *crs, const ProveArgs *args, uint32_t K, uint64_t *proof_a, uint64_t *proof_b, uint64_t *proof_c,
                        uint8_t *inf_flags);

}  // namespace b200zk

// ---- the opaque handle types of the C ABI (defined here so that every translation unit sees one definition)
struct b200zk_ctx : public b200zk::Ctx {};
struct b200zk_bases : public b200zk::Bases {};
struct b200zk_crs : public b200zk::Crs {};
struct b200zk_job { b200zk_ctx *ctx; int slot; int group; };

namespace b200zk {
// A partial result travels as one RECORD: the Jacobian point (144 / 288 B) at +0 and the status word at +REC_STATUS.
static constexpr size_t REC_STATUS = 288, REC_BYTES = 320;
// Enqueue one multiexp with HOST operands on `ctx` (copy stream -> compute stream) and leave its record in the job slot
// *slot_out (device: slot.dev + slot.o_res).  d_record_out != nullptr: the window-combine kernel writes the record there instead
// (a peer pointer: the partial of a shard goes straight into the gather buffer of the combining GPU).  The caller holds no lock.
int multiexp_enqueue(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp, const uint8_t *density,
                     void *d_record_out, int *slot_out);
int first_status(Ctx *ctx, cudaStream_t st, const void *d_records, size_t n, void *d_status_out);  // first non-zero status word, in record order
}  // namespace b200zk
