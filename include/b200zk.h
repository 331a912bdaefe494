/* b200zk.h -- C ABI of libb200zk.so: B200 (sm_100a) Groth16 prover numerics.
 *
 * Drop-in boundary for the reference's (UrosTesic/zcash-gpu-thesis = librustzcash fork) prover hot path.
 * The reference has no FFI for this path (its OpenCL experiments bypass the API); the boundary is the Rust
 * signatures below, which a thin `-sys` crate binds to these entry points:
 *
 *   bellman/src/multiexp.rs:285-295   pub fn multiexp<Q,D,G,S>(pool, bases, density_map, exponents)
 *   bellman/src/multiexp.rs:19-68     SourceBuilder / Source for (Arc<Vec<G>>, usize)
 *   bellman/src/domain.rs:35-189      EvaluationDomain::{from_coeffs, fft, ifft, coset_fft, icoset_fft,
 *                                     distribute_powers, z, divide_by_z_on_coset, mul_assign, sub_assign}
 *   bellman/src/groth16/prover.rs:192-364  create_random_proof / create_proof (H-polynomial block :256-287)
 *
 * Conventions
 *   - All integers are little-endian u64 limbs exactly as in Rust memory: FqRepr([u64;6]), FrRepr([u64;4]).
 *     Field elements / point coordinates are in Montgomery form (what `Fq(FqRepr)` / `Fr(FrRepr)` hold);
 *     MSM scalars are canonical `FrRepr` (what `into_repr()` returns, prover.rs:287).
 *   - G1 affine = x||y (12 u64), G2 affine = x.c0||x.c1||y.c0||y.c1 (24 u64); the `infinity: bool` of
 *     G1Affine/G2Affine (ec.rs:13-18) travels as a separate byte (or via a stride into the Rust struct).
 *   - Projective results are Jacobian (X, Y, Z) Montgomery triples (18 / 36 u64) like `G1`/`G2` (ec.rs:20-24).
 *     The representative is not unique; compare with the reference's PartialEq (ec.rs:45-85) or after
 *     into_affine.  Field / affine / NTT outputs are canonical and bit-identical to the reference.
 *   - No exceptions, no aborts: every call returns a status; b200zk_last_error() has the text.
 *   - One context = one GPU + one CUDA stream.  Calls on one context are stream-ordered and thread-safe (each entry
 *     holds the context's lock, so concurrent callers are serialised -- `Worker` is Clone in the reference and may be
 *     shared); for concurrency use one context per thread, the _async futures (the prover keeps 8 multiexps in
 *     flight, prover.rs:289-318) or the batch entry points.
 *   - There is NO CPU fallback: without a CUDA device every compute entry returns B200ZK_ERR_CUDA.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; 1..3 map onto bellman's SynthesisError (bellman/src/lib.rs:171-188) */
enum {
    B200ZK_OK = 0,
    B200ZK_ERR_UNEXPECTED_IDENTITY = 1, /* SynthesisError::UnexpectedIdentity   (multiexp.rs:48-50)  */
    B200ZK_ERR_UNEXPECTED_EOF = 2,      /* SynthesisError::IoError(UnexpectedEof) (multiexp.rs:44-46) */
    B200ZK_ERR_DEGREE_TOO_LARGE = 3,    /* SynthesisError::PolynomialDegreeTooLarge (domain.rs:59-61) */
    B200ZK_ERR_BAD_ARG = 4,
    B200ZK_ERR_CUDA = 5,
    B200ZK_ERR_NCCL = 6,
    B200ZK_ERR_DECODE = 7               /* GroupDecodingError (pairing/src/lib.rs:512-530) from the wire-format entry points */
};

enum { B200ZK_G1 = 1, B200ZK_G2 = 2 };
enum { B200ZK_FR = 0, B200ZK_FQ = 1, B200ZK_FQ2 = 2 /* c0 || c1, 12 u64 (fq2.rs) */,
       B200ZK_FQ2_PAIR = 3 /* the same elements computed by the lane-pair Fq2 of the G2 bucket accumulation (one component per lane):
                              ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, and MULSUB on pairs of Fq2 elements */ };
/* EvaluationDomain transforms, domain.rs:83-132 */
enum { B200ZK_FFT = 0, B200ZK_IFFT = 1, B200ZK_COSET_FFT = 2, B200ZK_ICOSET_FFT = 3 };
/* element-wise field ops (fr.rs / fq.rs); used by the parity tests and by EvaluationDomain::{mul,sub}_assign */
enum {
    B200ZK_OP_ADD = 0, B200ZK_OP_SUB = 1, B200ZK_OP_MUL = 2, B200ZK_OP_SQUARE = 3, B200ZK_OP_DOUBLE = 4,
    B200ZK_OP_NEGATE = 5, B200ZK_OP_INTO_REPR = 6, B200ZK_OP_FROM_REPR = 7, B200ZK_OP_INVERSE = 8,
    B200ZK_OP_INVERSE_BINARY = 9, /* same value as INVERSE, by the reference's binary extended Euclid (fq.rs:849-903) */
    B200ZK_OP_MULSUB = 10 /* Fq and FQ2_PAIR: a[i] = (p, q), b[i] = (r, s) pairs; out[i] = p q - r s (the fused Y3 of the point additions) */
};
/* point ops (ec.rs:296-526) for the parity tests */
enum { B200ZK_POINT_DOUBLE = 0, B200ZK_POINT_ADD = 1, B200ZK_POINT_ADD_MIXED = 2 };

typedef struct b200zk_ctx b200zk_ctx;
typedef struct b200zk_bases b200zk_bases;
typedef struct b200zk_crs b200zk_crs;

/* ---- context ------------------------------------------------------------------------------------------------- */
/* Replaces Worker::new() (bellman/src/multicore.rs:24-31): a context owns one CUDA stream on `device`. */
int b200zk_init(int device, b200zk_ctx **out);
void b200zk_destroy(b200zk_ctx *ctx);
const char *b200zk_last_error(b200zk_ctx *ctx);
int b200zk_sync(b200zk_ctx *ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int b200zk_set_stream(b200zk_ctx *ctx, void *cuda_stream);
int b200zk_device_count(void);
int b200zk_sm_count(b200zk_ctx *ctx);
const char *b200zk_version(void);

/* device memory + timing helpers (so a host with no CUDA bindings can keep operands resident in HBM) */
int b200zk_dev_alloc(b200zk_ctx *ctx, size_t bytes, void **dptr);
int b200zk_dev_free(b200zk_ctx *ctx, void *dptr);
int b200zk_h2d(b200zk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int b200zk_d2h(b200zk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int b200zk_d2d(b200zk_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes); /* stream-ordered device copy */
int b200zk_host_alloc_pinned(size_t bytes, void **hptr);
int b200zk_host_free_pinned(void *hptr);
int b200zk_timer_start(b200zk_ctx *ctx);            /* cudaEventRecord on the context's stream */
int b200zk_timer_stop(b200zk_ctx *ctx, float *ms);  /* record + synchronize + elapsed          */

/* ---- bases (the SourceBuilder `(Arc<Vec<G>>, usize)`, multiexp.rs:34-68; ParameterSource, groth16/mod.rs:395-482) */
/* Upload n affine bases once and keep them resident in HBM (the CRS is constant across proofs).
 * `points` + i*stride -> x||y Montgomery limbs (96 B for G1, 192 B for G2); stride = 96/192 for packed arrays or
 * sizeof(G1Affine)=104 / sizeof(G2Affine)=200 to read a Rust Vec<G1Affine> in place.
 * `infinity` + i*inf_stride -> the `infinity: bool` byte, or NULL when no base is the identity. */
int b200zk_bases_upload(b200zk_ctx *ctx, int group, const void *points, size_t n, size_t stride, const uint8_t *infinity,
                        size_t inf_stride, b200zk_bases **out);
/* Same from device memory (packed x||y, optional infinity bytes). The data is copied. */
int b200zk_bases_from_device(b200zk_ctx *ctx, int group, const void *d_points, size_t n, const uint8_t *d_infinity, b200zk_bases **out);
/* Optional, one-time (CRS load): build the W-fold table 2^(c w) * P_i (affine) in HBM so that every Pippenger window adds
 * into one shared bucket set -- the per-window joins of multiexp.rs:223-229 (c doublings each) and W - 1 of the W bucket
 * reductions disappear and wider windows pay off.  Costs W x the base memory (12 x 1.5 GiB for 2^24 G1 points at c = 22).
 * window_bits = 0 picks c from n.  Later multiexps on these bases use the table automatically.  Results are unchanged. */
int b200zk_bases_precompute(b200zk_ctx *ctx, b200zk_bases *bases, int window_bits);
/* The wire format of Parameters::read (groth16/mod.rs:287-382): n uncompressed big-endian points (96 B G1 / 192 B G2, Fq2 as
 * c1 then c0; pairing/src/bls12_381/README.md:59-75).  Decoded on the device straight into Montgomery limbs
 * (into_affine_unchecked, ec.rs:686-736); checked != 0 adds the curve-equation and subgroup tests of into_affine
 * (ec.rs:125-144).  Any bad point -> B200ZK_ERR_DECODE with its index and reason in b200zk_last_error(); a point at infinity
 * is such an error unless allow_infinity (Parameters::read rejects it, mod.rs:300-304). */
int b200zk_bases_upload_encoded(b200zk_ctx *ctx, int group, const uint8_t *bytes, size_t n, int checked, int allow_infinity, b200zk_bases **out);
/* Same decode / the matching encode for host arrays (VerifyingKey elements, Proof::write groth16/mod.rs:43-53):
 * compressed != 0 writes the 48 / 96-byte form with the sign flag (ec.rs:839-868, 2801-2830). */
int b200zk_decode_points(b200zk_ctx *ctx, int group, const uint8_t *bytes, size_t n, int checked, uint64_t *out_xy, uint8_t *out_inf);
int b200zk_encode_points(b200zk_ctx *ctx, int group, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out_bytes);
size_t b200zk_bases_len(const b200zk_bases *bases);
void b200zk_bases_free(b200zk_bases *bases);

/* ---- multiexp (bellman/src/multiexp.rs:285-335) ------------------------------------------------------------------ */
/* result = sum over i with density[i] != 0 of scalars[i] * bases[base_offset + rank(i)], rank(i) = number of set
 * density bytes before i (density == NULL is FullDensity: rank(i) = i).  Semantics of multiexp_inner
 * (multiexp.rs:140-233): a zero scalar skips its base; a base at infinity that is *consumed* (non-zero scalar)
 * -> UNEXPECTED_IDENTITY; running out of bases -> UNEXPECTED_EOF; the first offending exponent in iteration
 * order decides which of the two is reported.  scalars: n_exp x 4 u64 canonical FrRepr.
 * out_jacobian: 18 (G1) / 36 (G2) u64.  Synchronous: returns when the result is in `out_jacobian`. */
int b200zk_multiexp(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                    const uint8_t *density, uint64_t *out_jacobian);
/* Device-resident operands; result written to device memory, no host synchronisation (stream-ordered).
 * The status word (0/1/2 as above) is written to d_status (4 bytes, device) and also folded into the next
 * synchronous call's return value; pass NULL to ignore. */
int b200zk_multiexp_dev(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp,
                        const uint8_t *d_density, void *d_out_jacobian, void *d_status);
/* `batch` multiexps over the same bases in one pipeline (the multiexps of `batch` proofs over one CRS, or the futures
 * prover.rs:289-318 holds together): exponent vector k starts k * scalar_stride exponents after d_scalars, density map k
 * (if any) k * density_stride bytes after d_density; each has n_exp exponents and starts at base_offset.  Results: `batch`
 * consecutive Jacobian points in d_out_jacobians and `batch` status words in d_status (device memory, stream-ordered). */
int b200zk_multiexp_batch_dev(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp,
                              size_t scalar_stride, const uint8_t *d_density, size_t density_stride, uint32_t batch, void *d_out_jacobians,
                              void *d_status);
/* The reference's multiexp() returns a future and the prover keeps several in flight (prover.rs:289-318, 339-354).
 * _async enqueues the host->device copy of the exponents on the context's copy stream and the multiexp behind it on the
 * compute stream, then returns; copies of later jobs overlap the computation of earlier ones.  `scalars` / `density` must
 * stay valid (and should be pinned) until b200zk_job_wait, which blocks, writes the Jacobian result, returns the same status
 * codes as b200zk_multiexp and releases the job.  Jobs of one context complete in submission order. */
typedef struct b200zk_job b200zk_job;
int b200zk_multiexp_async(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                          const uint8_t *density, b200zk_job **job);
/* Multi-GPU form: this rank's (bases, scalars) are its shard; the job's result is the sum over all ranks of the communicator
 * (NCCL all-gather of the partials + adds, enqueued right behind the shard's multiexp). Every rank must submit the call. */
int b200zk_multiexp_sharded_async(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                                  const uint8_t *density, b200zk_job **job);
int b200zk_job_wait(b200zk_job *job, uint64_t *out_jacobian);
/* Pippenger window override for tuning (0 = automatic). */
int b200zk_set_msm_window(b200zk_ctx *ctx, int window_bits);

/* Sum of n Jacobian points (device, packed 144 / 288 B each) -> one Jacobian point (device). Used for the
 * cross-GPU combine of per-shard partial sums and by tests (multiexp.rs:942-1200 reduction tests). */
int b200zk_sum_points_dev(b200zk_ctx *ctx, int group, const void *d_points, size_t n, void *d_out);
/* Jacobian -> affine on the device (ec.rs:586-619); out: x||y and one infinity byte per point. Host buffers. */
int b200zk_into_affine(b200zk_ctx *ctx, int group, const uint64_t *jacobian, size_t n, uint64_t *out_xy, uint8_t *out_inf);

/* Fixed-base batch multiplication out[i] = scalars[i] * base (ec.rs:87-99 semantics, windowed on the device).
 * Generates CRS-like synthetic bases in HBM (generator.rs:266-288 uses wNAF tables for the same job).
 * d_scalars: n x 4 u64 canonical, only the low `scalar_bits` bits are used; d_out_affine packed x||y. */
int b200zk_fixed_base_mul_dev(b200zk_ctx *ctx, int group, const uint64_t *base_affine_host, const void *d_scalars, size_t n,
                              uint32_t scalar_bits, void *d_out_affine, uint8_t *d_out_inf);

/* ---- multi-GPU (one process per GPU; MSM sharded by base range, NCCL gather of the partial sums) ------------------- */
/* Join an NCCL communicator. unique_id: 128 bytes from b200zk_nccl_unique_id() on rank 0, distributed by the launcher. */
int b200zk_nccl_unique_id(uint8_t out_id[128]);
int b200zk_comm_init(b200zk_ctx *ctx, const uint8_t unique_id[128], int rank, int world);
/* Each rank passes the partial Jacobian sum of its shard (device); all ranks receive the total (device).
 * ncclAllGather of one 320-byte record (point + status word) per rank on the context's stream followed by a (world-1)-add kernel.
 * In b200zk_multiexp_sharded_async the shard's last kernel writes its record straight into the send slot, and the status words
 * are reduced with the points: an UnexpectedIdentity / UnexpectedEof of any shard is returned by every rank. */
int b200zk_allgather_sum_dev(b200zk_ctx *ctx, int group, const void *d_partial, void *d_total);

/* ---- one process, several GPUs ------------------------------------------------------------------------------------------
 * The reference's only product caller issues all the multiexps of a proof from ONE process (prover.rs:289-318, reached from the C
 * ABI at librustzcash/src/rustzcash.rs:1556), so a drop-in must use several GPUs from one process: a *group* holds one context
 * per entry of `devices` (an entry may repeat: two shards on one GPU).  Bases are sharded by contiguous range over the group's
 * devices at upload; a multiexp cuts its exponent vector where the base cursor crosses a shard boundary (density-aware, see
 * b200zk_multi_plan), runs the shards concurrently and sums the per-shard partials on the first device -- the last kernel of a
 * shard stores its partial straight into the first device's memory over NVLink when peer access is available.  Semantics and
 * status codes are those of b200zk_multiexp (the first offending exponent in iteration order decides, whatever shard it is in). */
typedef struct b200zk_group b200zk_group;
typedef struct b200zk_group_bases b200zk_group_bases;
typedef struct b200zk_group_job b200zk_group_job;
int b200zk_init_multi(const int *devices, int n_dev, b200zk_group **out);
void b200zk_group_destroy(b200zk_group *g);
const char *b200zk_group_last_error(b200zk_group *g);
int b200zk_group_size(const b200zk_group *g);
b200zk_ctx *b200zk_group_ctx(b200zk_group *g, int i);        /* the context of shard i (for NTTs / proofs on a chosen GPU) */
int b200zk_group_peer_access(const b200zk_group *g, int i);  /* 1 when shard i stores its partial directly into the first device */
int b200zk_multi_bases_upload(b200zk_group *g, int group, const void *points, size_t n, size_t stride, const uint8_t *infinity,
                              size_t inf_stride, b200zk_group_bases **out);
int b200zk_multi_bases_precompute(b200zk_group *g, b200zk_group_bases *bases, int window_bits);
size_t b200zk_multi_bases_len(const b200zk_group_bases *bases);
void b200zk_multi_bases_free(b200zk_group_bases *bases);
int b200zk_multi_multiexp(b200zk_group *g, const b200zk_group_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                          const uint8_t *density, uint64_t *out_jacobian);
/* the future-returning form (up to 4 in flight per group); scalars / density must stay valid until the wait */
int b200zk_multi_multiexp_async(b200zk_group *g, const b200zk_group_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                                const uint8_t *density, b200zk_group_job **job);
int b200zk_multi_job_wait(b200zk_group_job *job, uint64_t *out_jacobian);
/* The H block of create_proof (prover.rs:256-287) over the group: a, b and c are independent until the pointwise a * b - c
 * (the reference runs them as three scoped tasks, prover.rs:257-266), so each is uploaded to and transformed on its own GPU
 * (ifft + coset_fft, round-robin over the group); b and c then cross to the first device peer to peer and it finishes.  Same
 * arguments and result as b200zk_h_poly.  Pays off for large domains (Sprout's m = 2^21); at 2^17 one GPU is as fast. */
int b200zk_multi_h_poly(b200zk_group *g, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint32_t log_m, uint64_t *out);
/* Host-only helper (needs no device): where a sharded multiexp cuts its exponents.  bounds[0..n_dev] = base-range boundaries of
 * the shards; e_lo[0..n_dev] receives the exponent split points, local_offset[0..n_dev-1] the base cursor of each shard relative
 * to its own first base.  The last shard also takes every exponent beyond the end of the bases (it reports UnexpectedEof). */
int b200zk_multi_plan(const size_t *bounds, int n_dev, size_t base_offset, const uint8_t *density, size_t n_exp, size_t *e_lo,
                      size_t *local_offset);

/* ---- EvaluationDomain (bellman/src/domain.rs) -------------------------------------------------------------------- */
/* Host-only helper (needs no device): how a transform of 2^log_m elements is cut into passes over HBM -- stages[i] butterfly stages
 * and 2^columns_log[i] adjacent columns per tile in pass i, *radix4 = 1 for the large-transform kernels (radix-4 rounds), 0 for the
 * latency-oriented ones.  `large_from` = the log2 size from which the large kernels are used (B200ZK_NTT_LARGE_FROM, default 20),
 * `batch` = equal transforms launched together.  Arrays of 8 entries.  Returns the number of passes (0 for log_m outside 3..30:
 * smaller transforms run in one thread).  The tests use it to check that every size maps to kernels the library was built with. */
int b200zk_ntt_plan(uint32_t log_m, int large_from, int sm_count, uint32_t batch, uint32_t *stages, uint32_t *columns_log, int *radix4);
/* In-place transform of m = 2^log_m Montgomery Fr coefficients, natural order in and out (domain.rs:83-132).
 * log_m >= 32 (= Fr::S) -> DEGREE_TOO_LARGE like from_coeffs (domain.rs:59-61). */
int b200zk_ntt(b200zk_ctx *ctx, uint64_t *coeffs_host, uint32_t log_m, int kind);
int b200zk_ntt_dev(b200zk_ctx *ctx, void *d_coeffs, uint32_t log_m, int kind);
/* distribute_powers (domain.rs:105-118): coeffs[i] *= g^i, g given as 4 u64 Montgomery limbs (host). */
int b200zk_distribute_powers_dev(b200zk_ctx *ctx, void *d_coeffs, size_t n, const uint64_t g[4]);
/* element-wise ops on device vectors of Fr/Fq elements: out[i] = op(a[i], b[i]) (b ignored for unary ops) */
int b200zk_field_vec_dev(b200zk_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n);
int b200zk_field_vec(b200zk_ctx *ctx, int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* divide_by_z_on_coset (domain.rs:146-159): coeffs[i] *= 1 / (g^m - 1), g = multiplicative_generator() = 7 */
int b200zk_divide_by_z_on_coset_dev(b200zk_ctx *ctx, void *d_coeffs, uint32_t log_m);
/* z(tau) = tau^m - 1 (domain.rs:136-141); tau and the result are Montgomery limbs on the host */
int b200zk_domain_z(b200zk_ctx *ctx, const uint64_t tau[4], uint32_t log_m, uint64_t out[4]);
/* coeffs[i] *= s (ifft's m^-1 scaling; domain.rs:88-103) */
int b200zk_fr_scale_dev(b200zk_ctx *ctx, void *d_coeffs, size_t n, const uint64_t s[4]);
/* y[v] = sum over k in [row_ptr[v], row_ptr[v+1]) of val[k] * x[col[k]]  (Fr, Montgomery; u32 CSR indices; device memory,
 * stream-ordered).  The numeric core of generate_parameters' eval (generator.rs:301-414): row v = the (coeff, constraint)
 * terms of variable v in A, B or C, x = the Lagrange coefficients L_i(tau) (powers of tau after ifft, generator.rs:292). */
int b200zk_fr_spmv_dev(b200zk_ctx *ctx, const void *d_row_ptr, const void *d_col, const void *d_val, const void *d_x, size_t n_rows, void *d_y);
/* point ops on host arrays (parity tests): a = Jacobian points; b = Jacobian (ADD) or affine x||y (ADD_MIXED) */
int b200zk_point_op(b200zk_ctx *ctx, int group, int op, const uint64_t *a, const uint64_t *b, const uint8_t *b_inf, uint64_t *out, size_t n);

/* The H-polynomial block of create_proof (prover.rs:256-287), fused on the device:
 * a,b,c = evaluation vectors padded to m = 2^log_m (Montgomery); out = (m-1) x 4 u64 canonical coefficients
 * (the `into_repr` exponents of the H multiexp).  a, b, c are clobbered in the _dev form. */
int b200zk_h_poly(b200zk_ctx *ctx, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint32_t log_m, uint64_t *out);
int b200zk_h_poly_dev(b200zk_ctx *ctx, void *d_a, void *d_b, void *d_c, uint32_t log_m, void *d_out);

/* ---- groth16::create_proof (bellman/src/groth16/prover.rs:205-364), everything after circuit synthesis --------------- */
/* A proving key resident in HBM: the five query vectors of `Parameters` (groth16/mod.rs:215-238) as bases handles
 * (h, l, a, b_g1, b_g2; "never contains points at infinity") plus the VerifyingKey elements the prover touches
 * (vk.alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2: affine x||y Montgomery; vk_infinity[5] flags or NULL).
 * Window tables for delta_g1 / delta_g2 are built once here.  The handles must outlive the CRS object. */
int b200zk_crs_create(b200zk_ctx *ctx, const b200zk_bases *h, const b200zk_bases *l, const b200zk_bases *a, const b200zk_bases *b_g1,
                      const b200zk_bases *b_g2, const uint64_t alpha_g1[12], const uint64_t beta_g1[12], const uint64_t beta_g2[24],
                      const uint64_t delta_g1[12], const uint64_t delta_g2[24], const uint8_t *vk_infinity, b200zk_crs **out);
void b200zk_crs_free(b200zk_crs *crs);
/* Parameters::read (groth16/mod.rs:287-382; VerifyingKey::read :160-212): the whole proving key from its wire format -- vk
 * (alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, u32 count + ic), then h, l, a, b_g1, b_g2 each as a big-endian u32
 * count + uncompressed points.  The vk points are always fully checked (into_affine) and may be the identity; ic and the query
 * vectors may not ("point at infinity"); `checked` selects into_affine / into_affine_unchecked for the query vectors.  Points are
 * decoded on the device and stay resident; the CRS owns them (freed by b200zk_crs_free).  Bad or truncated input ->
 * B200ZK_ERR_DECODE with the reason in b200zk_last_error. */
int b200zk_parameters_read(b200zk_ctx *ctx, const uint8_t *bytes, size_t len, int checked, b200zk_crs **out);
/* Parameters::write (groth16/mod.rs:252-285) of a CRS made by b200zk_parameters_read: byte-identical to what was read. */
size_t b200zk_parameters_size(const b200zk_crs *crs);
int b200zk_parameters_write(b200zk_ctx *ctx, const b200zk_crs *crs, uint8_t *out, size_t cap);
/* The VerifyingKey of such a CRS: 108 words (alpha_g1 | beta_g1 | beta_g2 | gamma_g2 | delta_g1 | delta_g2, affine Montgomery),
 * six infinity flags and ic (12 words per element; pass ic = NULL to query the count first). */
int b200zk_crs_verifying_key(const b200zk_crs *crs, uint64_t vk[108], uint8_t vk_inf[6], uint64_t *ic, size_t *n_ic);
int b200zk_crs_query_sizes(const b200zk_crs *crs, size_t sizes[5]);            /* lengths of h, l, a, b_g1, b_g2 */
int b200zk_crs_precompute(b200zk_ctx *ctx, b200zk_crs *crs, int window_bits);  /* b200zk_bases_precompute on every query vector */
/* create_proof for an already synthesized ProvingAssignment (prover.rs:84-190): a, b, c = the evaluation vectors
 * (n_constraints x 4 u64 Montgomery, including the `x * 0 = 0` input rows of prover.rs:228-234); inputs / aux = the
 * assignments as canonical FrRepr (prover.rs:290-291); the three density maps as bytes; r, s canonical FrRepr.
 * Outputs: proof.a (G1 affine x||y), proof.b (G2 affine), proof.c (G1 affine), Montgomery limbs, + 3 infinity flags.
 * Errors as the reference: UNEXPECTED_IDENTITY (subversion check / identity base), UNEXPECTED_EOF, DEGREE_TOO_LARGE. */
int b200zk_groth16_prove(b200zk_ctx *ctx, const b200zk_crs *crs, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                         const uint64_t *inputs, size_t n_inputs, const uint64_t *aux, size_t n_aux, const uint8_t *a_aux_density,
                         const uint8_t *b_input_density, const uint8_t *b_aux_density, const uint64_t r[4], const uint64_t s[4],
                         uint64_t proof_a[12], uint64_t proof_b[24], uint64_t proof_c[12], uint8_t inf_flags[3]);

/* A batch of create_proof calls over one CRS and one circuit -- what a wallet or block producer issues as a run of
 * librustzcash_sapling_spend_proof calls (librustzcash/src/rustzcash.rs:1375; each ends in create_random_proof,
 * sapling-crypto/src/... prover.rs:192-203).  Every proof has n_constraints / n_inputs / n_aux of the same size; the
 * per-proof pointers have the meaning of b200zk_groth16_prove's arguments.  The proofs are proved `lockstep` at a time
 * (0 = 8): the five multiexps of those proofs run as five batched multiexps, which fills the GPU where one 10^5-point
 * multiexp cannot.  Outputs: n_proofs consecutive proofs (12 / 24 / 12 u64 and 3 flags each).  On an error the status
 * is that of the first failing group; proofs of earlier groups are complete. */
typedef struct b200zk_prove_input {
    const uint64_t *a, *b, *c;          /* n_constraints x 4 u64, Montgomery */
    const uint64_t *inputs, *aux;       /* n_inputs x 4, n_aux x 4 canonical FrRepr */
    const uint8_t *a_aux_density;       /* n_aux bytes */
    const uint8_t *b_input_density;     /* n_inputs bytes */
    const uint8_t *b_aux_density;       /* n_aux bytes */
    const uint64_t *r, *s;              /* 4 u64 each, canonical FrRepr */
} b200zk_prove_input;
int b200zk_groth16_prove_batch(b200zk_ctx *ctx, const b200zk_crs *crs, const b200zk_prove_input *proofs, size_t n_proofs, size_t n_constraints,
                               size_t n_inputs, size_t n_aux, int lockstep, uint64_t *proofs_a, uint64_t *proofs_b, uint64_t *proofs_c,
                               uint8_t *inf_flags);

/* The batch as the outer FFI of the reference sees it: librustzcash_sapling_spend_proof (librustzcash/src/rustzcash.rs:1375-1626)
 * ends in create_random_proof + Proof::write into a 192-byte buffer (rustzcash.rs:1556-1601).  N assignments in, N x 192 proof
 * bytes out (a | b | c compressed, groth16/mod.rs:43-53), encoded on the device. */
int b200zk_groth16_prove_batch_bytes(b200zk_ctx *ctx, const b200zk_crs *crs, const b200zk_prove_input *proofs, size_t n_proofs, size_t n_constraints,
                                     size_t n_inputs, size_t n_aux, int lockstep, uint8_t *out_proofs);

/* ---- groth16::verify_proof (bellman/src/groth16/verifier.rs:18-66) as a batch verifier on the device ----------------------
 * The prover path needs no pairing; this closes the prove -> verify loop of the outer FFI (rustzcash.rs:1556-1601 verifies every
 * proof it has just made).  b200zk_prepare_verifying_key = prepare_verifying_key (verifier.rs:18-33): e(alpha_g1, beta_g2), -gamma_g2,
 * -delta_g2 and ic resident on the device.  b200zk_verify_proofs checks n proofs (affine Montgomery limbs as b200zk_groth16_prove
 * returns them, 3 infinity flags each or NULL) against n x n_inputs public inputs (canonical FrRepr, WITHOUT the leading ONE):
 * ok[i] = 1 iff e(A, B) = e(alpha, beta) e(sum_j input_j ic_j, gamma) e(C, delta).  n_inputs + 1 != ic.len() ->
 * B200ZK_ERR_BAD_ARG (MalformedVerifyingKey, verifier.rs:41-43).  One thread block per proof.
 * b200zk_pairing = Engine::pairing (pairing/src/lib.rs:86-96) for n independent pairs; out: n x 72 u64, the Fq12 value as
 * c0.c0.c0, c0.c0.c1, c0.c1.c0 ... (6 u64 Montgomery limbs each) -- the canonical value the reference computes. */
typedef struct b200zk_pvk b200zk_pvk;
int b200zk_pairing(b200zk_ctx *ctx, const uint64_t *g1_xy, const uint8_t *g1_inf, const uint64_t *g2_xy, const uint8_t *g2_inf, size_t n, uint64_t *out_fq12);
int b200zk_prepare_verifying_key(b200zk_ctx *ctx, const uint64_t alpha_g1[12], const uint64_t beta_g2[24], const uint64_t gamma_g2[24], const uint64_t delta_g2[24],
                                 const uint64_t *ic, size_t n_ic, b200zk_pvk **out);
void b200zk_pvk_free(b200zk_pvk *pvk);
int b200zk_verify_proofs(b200zk_ctx *ctx, const b200zk_pvk *pvk, const uint64_t *proofs_a, const uint64_t *proofs_b, const uint64_t *proofs_c, const uint8_t *inf_flags,
                         const uint64_t *public_inputs, size_t n_inputs, size_t n_proofs, uint8_t *ok);

/* Per-kernel timing for the roofline report: when enabled, b200zk_multiexp(_dev) brackets its dominant kernel
 * (bucket accumulation) with CUDA events on the context's stream; read() synchronises and returns the summed
 * milliseconds and the number of launches since the last read. */
/* Kernels launched on this context by multiexp / ntt / h_poly / groth16_prove since creation (or the last reset). */
unsigned long long b200zk_launch_count(b200zk_ctx *ctx, int reset);
int b200zk_profile_enable(b200zk_ctx *ctx, int on);
int b200zk_profile_read(b200zk_ctx *ctx, double *accumulate_ms, int *launches);

/* ---- calibration: integer-pipe roofline (SURVEY.md section 8d asks the build to measure it) -------------------------- */
/* kind: 0 = IMAD (32-bit mad.lo), 1 = IMAD.WIDE (mad.wide.u32), 2 = IMAD.HI, 3 = IADD3 carry chain,
 *       4 = Fq Montgomery multiply, 5 = Fr Montgomery multiply.  Returns operations per second over the whole GPU. */
int b200zk_microbench(b200zk_ctx *ctx, int kind, int iters, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
