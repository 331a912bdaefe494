// Element-wise field / point kernels: the device twins of the reference's per-primitive OpenCL test kernels
// (bellman/src/bls12-381.cl:799-887 test_fq_*, :1612-1700 test_fr_*, :1045-1170 test_projective_*), of
// EvaluationDomain::{mul_assign, sub_assign} and the scaling loops (bellman/src/domain.rs:88-103, 146-189),
// plus the integer-pipe calibration micro-benchmarks.
#include <algorithm>

#include "ec.cuh"
#include "internal.h"

namespace b200zk {

template <class F>
__global__ void k_field_vec(int op, const F *__restrict__ a, const F *__restrict__ b, F *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i];
    F r;
    switch (op) {
    case B200ZK_OP_ADD: r = x + b[i]; break;
    case B200ZK_OP_SUB: r = x - b[i]; break;
    case B200ZK_OP_MUL: r = x * b[i]; break;
    case B200ZK_OP_SQUARE: r = x.sqr(); break;
    case B200ZK_OP_DOUBLE: r = x.dbl(); break;
    case B200ZK_OP_NEGATE: r = x.neg(); break;
    case B200ZK_OP_INTO_REPR: r = x.from_mont(); break;
    case B200ZK_OP_FROM_REPR: r = x.to_mont(); break;
    case B200ZK_OP_INVERSE_BINARY: r = x.inverse_binary(); break;
    default: r = x.inverse(); break;
    }
    out[i] = r;
}

// Fq2 (fq2.rs:84-205): the same op codes on c0||c1 elements.  The two repr conversions act per component.
__global__ void k_fq2_vec(int op, const fq2_t *__restrict__ a, const fq2_t *__restrict__ b, fq2_t *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq2_t x = a[i];
    fq2_t r;
    switch (op) {
    case B200ZK_OP_ADD: r = x + b[i]; break;
    case B200ZK_OP_SUB: r = x - b[i]; break;
    case B200ZK_OP_MUL: r = x * b[i]; break;
    case B200ZK_OP_SQUARE: r = x.sqr(); break;
    case B200ZK_OP_DOUBLE: r = x.dbl(); break;
    case B200ZK_OP_NEGATE: r = x.neg(); break;
    case B200ZK_OP_INTO_REPR: r = {x.c0.from_mont(), x.c1.from_mont()}; break;
    case B200ZK_OP_FROM_REPR: r = {x.c0.to_mont(), x.c1.to_mont()}; break;
    case B200ZK_OP_INVERSE_BINARY: r = x.inverse_binary(); break;
    default: r = x.inverse(); break;
    }
    out[i] = r;
}
// a[i] = (p, q), b[i] = (r, s) as pairs of Fq elements; out[i] = p q - r s by the fused one-reduction body that every point
// addition uses for its Y3 (fp.cuh mulsub_call)
__global__ void k_fq_mulsub(const fq_t *__restrict__ a, const fq_t *__restrict__ b, fq_t *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fq_t::mulsub_call(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// The lane-pair Fq2 (fq2.cuh fq2h_t) on its own: two lanes per element, each holding one component.  MULSUB: a[i] = (p, q),
// b[i] = (r, s) pairs of Fq2 elements, out[i] = p q - r s (the four-product Y3 of a G2 addition).
__global__ void k_fq2_pair_vec(int op, const fq_t *__restrict__ a, const fq_t *__restrict__ b, fq_t *__restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const unsigned role = threadIdx.x & 1u;
    if (i >= n) return;  // (both lanes of a pair leave together)
    if (op == B200ZK_OP_MULSUB) {
        const fq2h_t p{a[4 * i + role]}, q{a[4 * i + 2 + role]}, r{b[4 * i + role]}, s{b[4 * i + 2 + role]};
        out[2 * i + role] = mul_sub(p, q, r, s).c;
        return;
    }
    const fq2h_t x{a[2 * i + role]};
    fq2h_t y{fq_t::zero()};
    if (b) y.c = b[2 * i + role];
    fq2h_t r;
    switch (op) {
    case B200ZK_OP_ADD: r = x + y; break;
    case B200ZK_OP_SUB: r = x - y; break;
    case B200ZK_OP_MUL: r = x * y; break;
    case B200ZK_OP_SQUARE: r = x.sqr(); break;
    case B200ZK_OP_DOUBLE: r = x.dbl(); break;
    default: r = x.neg(); break;
    }
    out[2 * i + role] = r.c;
}

int launch_field_vec(Ctx *ctx, int field, int op, const void *a, const void *b, void *out, size_t n) {
    if (n == 0) return B200ZK_OK;
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (field == B200ZK_FQ2_PAIR) {
        const bool ok = op == B200ZK_OP_ADD || op == B200ZK_OP_SUB || op == B200ZK_OP_MUL || op == B200ZK_OP_SQUARE || op == B200ZK_OP_DOUBLE ||
                        op == B200ZK_OP_NEGATE || op == B200ZK_OP_MULSUB;
        if (!ok) return set_error(ctx, B200ZK_ERR_BAD_ARG, "op not defined for the lane-pair Fq2");
        k_fq2_pair_vec<<<(unsigned)((2 * n + 127) / 128), 128, 0, ctx->stream>>>(op, (const fq_t *)a, (const fq_t *)b, (fq_t *)out, n);
        B200ZK_CUDA(ctx, cudaGetLastError());
        return B200ZK_OK;
    }
    if (op == B200ZK_OP_MULSUB) {
        if (field != B200ZK_FQ) return set_error(ctx, B200ZK_ERR_BAD_ARG, "MULSUB is an Fq op");
        k_fq_mulsub<<<blocks, 128, 0, ctx->stream>>>((const fq_t *)a, (const fq_t *)b, (fq_t *)out, n);
        B200ZK_CUDA(ctx, cudaGetLastError());
        return B200ZK_OK;
    }
    if (op < 0 || op > B200ZK_OP_INVERSE_BINARY) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad field op");
    if (field == B200ZK_FQ2)
        k_fq2_vec<<<blocks, 128, 0, ctx->stream>>>(op, (const fq2_t *)a, (const fq2_t *)b, (fq2_t *)out, n);
    else if (field == B200ZK_FR)
        k_field_vec<fr_t><<<blocks, 128, 0, ctx->stream>>>(op, (const fr_t *)a, (const fr_t *)b, (fr_t *)out, n);
    else if (field == B200ZK_FQ)
        k_field_vec<fq_t><<<blocks, 128, 0, ctx->stream>>>(op, (const fq_t *)a, (const fq_t *)b, (fq_t *)out, n);
    else
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad field");
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

__global__ void k_fr_scale(fr_t *__restrict__ a, const fr_t *__restrict__ s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = a[i] * s[0];
}
int launch_fr_scale(Ctx *ctx, void *a, const void *scalar_dev, size_t n) {
    if (n == 0) return B200ZK_OK;
    k_fr_scale<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((fr_t *)a, (const fr_t *)scalar_dev, n);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// ---- y = M x over Fr, M in CSR form: the evaluation of the QAP polynomials at tau in generate_parameters
// (generator.rs:361-380 eval_at_tau: sum over the (coeff, constraint index) terms of one variable of coeff * L_index(tau)).
// One warp per row: rows are short except the one of the constant ONE, which touches most constraints.
__global__ void __launch_bounds__(128) k_fr_spmv(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const fr_t *__restrict__ val,
                                                const fr_t *__restrict__ x, size_t n_rows, fr_t *__restrict__ y) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        fr_t acc = fr_t::zero();
        for (uint32_t k = row_ptr[row] + lane; k < row_ptr[row + 1]; k += 32) acc = acc + val[k] * x[col[k]];
        for (uint32_t off = 16; off > 0; off >>= 1) {
            fr_t other;
#pragma unroll
            for (int j = 0; j < 8; j++) other.v[j] = __shfl_down_sync(0xffffffffu, acc.v[j], off);
            acc = acc + other;
        }
        if (lane == 0) y[row] = acc;
    }
}
int launch_fr_spmv(Ctx *ctx, const void *row_ptr, const void *col, const void *val, const void *x, size_t n_rows, void *y) {
    if (n_rows == 0) return B200ZK_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((n_rows + 3) / 4, (size_t)ctx->sm_count * 16);
    k_fr_spmv<<<blocks, 128, 0, ctx->stream>>>((const uint32_t *)row_ptr, (const uint32_t *)col, (const fr_t *)val, (const fr_t *)x, n_rows, (fr_t *)y);
    ctx->launches++;
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// ---- point ops, one thread per element (ec.rs:296-526)
template <class F>
__global__ void k_point_op(int op, const Jacobian<F> *__restrict__ a, const void *__restrict__ b, const uint8_t *__restrict__ b_inf,
                           Jacobian<F> *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jacobian<F> p = a[i];
    if (op == B200ZK_POINT_DOUBLE) {
        jacobian_double(p);
    } else if (op == B200ZK_POINT_ADD) {
        jacobian_add(p, ((const Jacobian<F> *)b)[i]);
    } else {
        jacobian_add_mixed(p, ((const Affine<F> *)b)[i], b_inf ? b_inf[i] != 0 : false);
    }
    out[i] = p;
}
int launch_point_op(Ctx *ctx, int group, int op, const void *a, const void *b, const uint8_t *b_inf, void *out, size_t n) {
    if (n == 0) return B200ZK_OK;
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == B200ZK_G1)
        k_point_op<fq_t><<<blocks, 64, 0, ctx->stream>>>(op, (const g1_jac_t *)a, b, b_inf, (g1_jac_t *)out, n);
    else if (group == B200ZK_G2)
        k_point_op<fq2_t><<<blocks, 64, 0, ctx->stream>>>(op, (const g2_jac_t *)a, b, b_inf, (g2_jac_t *)out, n);
    else
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

// ---- integer-pipe calibration -----------------------------------------------------------------------------------
// Independent instruction streams (8 accumulators per thread) so the pipe, not latency, is measured.
template <int KIND>
__global__ void k_microbench(int iters, uint32_t seed, uint32_t *out) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed + threadIdx.x * 8 + i; y[i] = seed * 3 + i; }
    uint32_t m = seed | 1;
    if (KIND == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i]));
        }
    } else if (KIND == 1) {
        unsigned long long w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = x[i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) {  // the multiplicand depends on the previous result: nothing is loop invariant
                    uint32_t lo = (uint32_t)w[i];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(m));
                }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    } else if (KIND == 2) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i]));
        }
    } else if (KIND == 3) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(y[0]));
#pragma unroll
                for (int i = 1; i < 7; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("addc.u32 %0, %0, %1;" : "+r"(x[7]) : "r"(y[7]));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
__global__ void k_microbench_mul(int iters, uint32_t seed, uint32_t *out) {
    F a = F::one(), b = F::r2();
    a.v[0] ^= threadIdx.x;
    b.v[1] ^= seed;
    F c = a + b, d = a - b;
    for (int it = 0; it < iters; it++) {  // two independent chains per thread
        a = a * c;
        b = b * d;
    }
    F r = a + b;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) acc ^= r.v[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ops per launch are returned through *ops_per_launch so the caller can time with its own events
int launch_microbench(Ctx *ctx, int kind, int iters, int blocks, int threads, void *out) {
    uint32_t *o = (uint32_t *)out;
    switch (kind) {
    case 0: k_microbench<0><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    case 1: k_microbench<1><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    case 2: k_microbench<2><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    case 3: k_microbench<3><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    case 4: k_microbench_mul<fq_t><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    case 5: k_microbench_mul<fr_t><<<blocks, threads, 0, ctx->stream>>>(iters, 12345u, o); break;
    default: return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad microbench kind");
    }
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
