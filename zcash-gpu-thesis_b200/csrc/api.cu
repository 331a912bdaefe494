// C ABI of libb200zk.so (include/b200zk.h): context / memory plumbing and the thin entry points that the
// reference-side bindings call in place of bellman::multiexp::multiexp (multiexp.rs:285), the
// EvaluationDomain methods (domain.rs:83-189) and the H block of create_proof (prover.rs:256-287).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>

#include "internal.h"

namespace b200zk {

int set_error(Ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

int ensure_scratch(Ctx *ctx, void **buf, size_t *cur, size_t bytes) {
    if (bytes <= *cur) return B200ZK_OK;
    if (*buf) {
        // earlier work on the stream may still use the old buffer
        B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        B200ZK_CUDA(ctx, cudaFree(*buf));
        *buf = nullptr;
        *cur = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(buf, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        B200ZK_CUDA(ctx, cudaMalloc(buf, want));
    }
    *cur = want;
    return B200ZK_OK;
}

// ---- NCCL through dlopen: the library has no link-time dependency on it; single-GPU users never load it
struct NcclId { char internal[128]; };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /* ncclUniqueId by value */ NcclId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool load_nccl(std::string &err) {
    std::lock_guard<std::mutex> l(g_nccl_mu);
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) { err = "NCCL symbols missing"; dlclose(h); return false; }
    g_nccl.handle = h;
    return true;
}

}  // namespace b200zk

using namespace b200zk;

namespace b200zk {
int ctx_lanes(Ctx *ctx, int n) {
    while ((int)ctx->lanes.size() < n) {
        b200zk_ctx *lane = nullptr;
        int rc = b200zk_init(ctx->device, &lane);
        if (rc) return set_error(ctx, rc, "could not create a prover lane");
        if (ctx->lanes.size() >= 1) {
            // lanes 1-3 carry the A, B-G1 and B-G2 multiexps of a proof, each followed by a serial piece of the assembly (a 255-bit
            // scalar multiplication / the G2 chain): the longest chains of the call, so their blocks are scheduled first
            int lo = 0, hi = 0;
            cudaStream_t s = nullptr;
            if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess && cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi) == cudaSuccess) {
                cudaStreamDestroy(lane->stream);
                lane->stream = s;
            }
            cudaGetLastError();
        }
        ctx->lanes.push_back(lane);
    }
    return B200ZK_OK;
}
}  // namespace b200zk

#define CHECK_CTX(ctx) do { if (!(ctx)) return B200ZK_ERR_BAD_ARG; } while (0)
#define USE_DEVICE(ctx)                                     \
    std::lock_guard<std::recursive_mutex> ctx_lock__((ctx)->mu); \
    B200ZK_CUDA(ctx, cudaSetDevice((ctx)->device))

extern "C" {

const char *b200zk_version(void) { return "b200zk 0.1 (sm_100a)"; }

int b200zk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b200zk_init(int device, b200zk_ctx **out) {
    if (!out) return B200ZK_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return B200ZK_ERR_CUDA; }  // no CPU fallback
    if (device < 0 || device >= n) return B200ZK_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return B200ZK_ERR_CUDA;
    b200zk_ctx *ctx = new b200zk_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B200ZK_ERR_CUDA; }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char *e = getenv("B200ZK_NTT_LARGE_FROM")) ctx->ntt_large_from = atoi(e);
    *out = ctx;
    return B200ZK_OK;
}

void b200zk_destroy(b200zk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (Ctx *lane : ctx->lanes) b200zk_destroy(static_cast<b200zk_ctx *>(lane));
    ctx->lanes.clear();
    cudaSetDevice(ctx->device);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    ntt_free_all_tables(ctx);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    cudaFree(ctx->gather_buf);
    cudaFree(ctx->small_slot);
    cudaFree(ctx->scratch);
    cudaFree(ctx->scratch2);
    cudaFree(ctx->scratch3);
    for (auto &sl : ctx->slots) { cudaFree(sl.dev); if (sl.host_res) cudaFreeHost(sl.host_res); if (sl.copied) cudaEventDestroy(sl.copied); if (sl.done) cudaEventDestroy(sl.done); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *b200zk_last_error(b200zk_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int b200zk_sync(b200zk_ctx *ctx) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_set_stream(b200zk_ctx *ctx, void *cuda_stream) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return B200ZK_OK;
}

int b200zk_sm_count(b200zk_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int b200zk_dev_alloc(b200zk_ctx *ctx, size_t bytes, void **dptr) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return B200ZK_OK;
}
int b200zk_dev_free(b200zk_ctx *ctx, void *dptr) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    B200ZK_CUDA(ctx, cudaFree(dptr));
    return B200ZK_OK;
}
int b200zk_h2d(b200zk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B200ZK_OK;
}
int b200zk_d2d(b200zk_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200ZK_OK;
}
int b200zk_d2h(b200zk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}
int b200zk_host_alloc_pinned(size_t bytes, void **hptr) { return cudaMallocHost(hptr, bytes ? bytes : 1) == cudaSuccess ? B200ZK_OK : B200ZK_ERR_CUDA; }
int b200zk_host_free_pinned(void *hptr) { return cudaFreeHost(hptr) == cudaSuccess ? B200ZK_OK : B200ZK_ERR_CUDA; }

int b200zk_timer_start(b200zk_ctx *ctx) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return B200ZK_OK;
}
int b200zk_timer_stop(b200zk_ctx *ctx, float *ms) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    B200ZK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    B200ZK_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return B200ZK_OK;
}

// ---------------------------------------------------------------------------------------------------------------- bases
static size_t point_bytes(int group) { return group == B200ZK_G1 ? 96 : 192; }

int b200zk_bases_from_device(b200zk_ctx *ctx, int group, const void *d_points, size_t n, const uint8_t *d_infinity, b200zk_bases **out) {
    CHECK_CTX(ctx);
    if (!out || (group != B200ZK_G1 && group != B200ZK_G2)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group / out");
    USE_DEVICE(ctx);
    b200zk_bases *b = new b200zk_bases();
    b->ctx = ctx; b->group = group; b->n = n; b->points = nullptr; b->infinity = nullptr;
    size_t bytes = n * point_bytes(group);
    if (cudaMalloc(&b->points, bytes ? bytes : 1) != cudaSuccess) { delete b; return set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(bases) failed"); }
    if (bytes) cudaMemcpyAsync(b->points, d_points, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    if (d_infinity && n) {
        if (cudaMalloc((void **)&b->infinity, n) != cudaSuccess) { cudaFree(b->points); delete b; return set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(inf) failed"); }
        cudaMemcpyAsync(b->infinity, d_infinity, n, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = b;
    return B200ZK_OK;
}

int b200zk_bases_upload(b200zk_ctx *ctx, int group, const void *points, size_t n, size_t stride, const uint8_t *infinity, size_t inf_stride,
                        b200zk_bases **out) {
    CHECK_CTX(ctx);
    if (!out || (group != B200ZK_G1 && group != B200ZK_G2)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group / out");
    const size_t pb = point_bytes(group);
    if (stride < pb) return set_error(ctx, B200ZK_ERR_BAD_ARG, "stride smaller than a point");
    USE_DEVICE(ctx);
    b200zk_bases *b = new b200zk_bases();
    b->ctx = ctx; b->group = group; b->n = n; b->points = nullptr; b->infinity = nullptr;
    if (cudaMalloc(&b->points, n ? n * pb : 1) != cudaSuccess) { delete b; return set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(bases) failed"); }
    if (n) {
        // strided host layout (e.g. a Rust Vec<G1Affine>, 104 B records) -> packed device array
        cudaError_t e = cudaMemcpy2DAsync(b->points, pb, points, stride, pb, n, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(b->points); delete b; return set_error(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e)); }
        bool any_inf = false;
        if (infinity) {
            if (inf_stride == 0) inf_stride = 1;
            for (size_t i = 0; i < n && !any_inf; i++) any_inf = infinity[i * inf_stride] != 0;
        }
        if (any_inf) {
            if (cudaMalloc((void **)&b->infinity, n) != cudaSuccess) { cudaFree(b->points); delete b; return set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(inf) failed"); }
            cudaMemcpy2DAsync(b->infinity, 1, infinity, inf_stride, 1, n, cudaMemcpyHostToDevice, ctx->stream);
        }
    }
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = b;
    return B200ZK_OK;
}

int b200zk_bases_precompute(b200zk_ctx *ctx, b200zk_bases *bases, int window_bits) {
    CHECK_CTX(ctx);
    if (!bases) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null bases");
    USE_DEVICE(ctx);
    uint32_t c = (uint32_t)window_bits;
    if (window_bits <= 0) {
        // Automatic.  Measured on B200 (G1 and G2, 2^16 ... 2^26, profiles/r02_msm_window_sweep.txt): what counts is the number
        // of windows W = ceil(255 / c) (mixed adds per point) and a TOP window that is not nearly empty -- with 255 = (W - 1) c + t
        // the 2^(t-1) buckets of the top window receive n / 2^(t-1) points each, and t = 2 or 3 (c = 18, 21, 23) costs 20-30 %.
        // c = 16 / 20 / 22 / 24 leave t = 15 / 15 / 13 / 15.
        uint32_t lg = 0;
        while (((size_t)1 << (lg + 1)) <= bases->n) lg++;
        c = lg <= 18 ? lg : lg <= 21 ? 20 : lg <= 25 ? 22 : 24;
        if (const char *e = getenv("B200ZK_PRE_DELTA")) c = lg - (uint32_t)atoi(e);
        if (c < 8) c = 8;
        // 255 = 15 x 17: with c = 15 or 17 the scalar fills its windows exactly and the signed-digit carry of the last one lands in
        // an extra window whose only bucket (digit 1) then receives half of all the points; one bit more avoids that
        if (255 % c == 0) c++;
        if (const char *e = getenv("B200ZK_PRE_C")) c = (uint32_t)atoi(e);
    }
    return msm_precompute(ctx, bases, c);
}

static int decode_error(b200zk_ctx *ctx, unsigned long long e) {
    static const char *why[] = {"ok", "unexpected compression mode", "unexpected information in the flag bits", "coordinate is not in the field",
                                "point is not on the curve", "point is not in the prime-order subgroup", "point at infinity"};
    unsigned code = (unsigned)(e & 0xff);
    return set_error(ctx, B200ZK_ERR_DECODE, "point " + std::to_string(e >> 8) + ": " + (code < 7 ? why[code] : "decoding error"));
}

// bytes (host) -> packed affine + infinity flags in freshly allocated device buffers
static int decode_to_device(b200zk_ctx *ctx, int group, const uint8_t *bytes, size_t n, int checked, int allow_infinity, void **d_pts, uint8_t **d_inf) {
    const size_t pb = point_bytes(group);
    *d_pts = nullptr;
    *d_inf = nullptr;
    void *d_bytes = nullptr;
    unsigned long long *d_err = nullptr;
    B200ZK_CUDA(ctx, cudaMalloc(d_pts, n ? n * pb : 1));
    B200ZK_CUDA(ctx, cudaMalloc((void **)d_inf, n ? n : 1));
    B200ZK_CUDA(ctx, cudaMalloc(&d_bytes, n ? n * pb : 1));
    B200ZK_CUDA(ctx, cudaMalloc((void **)&d_err, 8));
    B200ZK_CUDA(ctx, cudaMemsetAsync(d_err, 0xff, 8, ctx->stream));
    if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(d_bytes, bytes, n * pb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = codec_decode_uncompressed(ctx, group, d_bytes, n, checked, allow_infinity, *d_pts, *d_inf, d_err);
    unsigned long long e = ~0ull;
    if (!rc && cudaMemcpyAsync(&e, d_err, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = B200ZK_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, B200ZK_ERR_CUDA, "decode failed");
    cudaFree(d_bytes);
    cudaFree(d_err);
    if (!rc && e != ~0ull) rc = decode_error(ctx, e);
    if (rc) { cudaFree(*d_pts); cudaFree(*d_inf); *d_pts = nullptr; *d_inf = nullptr; }
    return rc;
}

int b200zk_bases_upload_encoded(b200zk_ctx *ctx, int group, const uint8_t *bytes, size_t n, int checked, int allow_infinity, b200zk_bases **out) {
    CHECK_CTX(ctx);
    if (!out || (n && !bytes) || (group != B200ZK_G1 && group != B200ZK_G2)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad argument");
    USE_DEVICE(ctx);
    void *d_pts;
    uint8_t *d_inf;
    int rc = decode_to_device(ctx, group, bytes, n, checked, allow_infinity, &d_pts, &d_inf);
    if (rc) return rc;
    b200zk_bases *b = new b200zk_bases();
    b->ctx = ctx; b->group = group; b->n = n; b->points = d_pts; b->infinity = nullptr;
    if (allow_infinity) b->infinity = d_inf; else cudaFree(d_inf);  // without allow_infinity no base is the identity
    *out = b;
    return B200ZK_OK;
}

int b200zk_decode_points(b200zk_ctx *ctx, int group, const uint8_t *bytes, size_t n, int checked, uint64_t *out_xy, uint8_t *out_inf) {
    CHECK_CTX(ctx);
    if ((n && (!bytes || !out_xy)) || (group != B200ZK_G1 && group != B200ZK_G2)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad argument");
    USE_DEVICE(ctx);
    void *d_pts;
    uint8_t *d_inf;
    int rc = decode_to_device(ctx, group, bytes, n, checked, 1, &d_pts, &d_inf);
    if (rc) return rc;
    cudaError_t e1 = n ? cudaMemcpy(out_xy, d_pts, n * point_bytes(group), cudaMemcpyDeviceToHost) : cudaSuccess;
    cudaError_t e2 = (n && out_inf) ? cudaMemcpy(out_inf, d_inf, n, cudaMemcpyDeviceToHost) : cudaSuccess;
    cudaFree(d_pts);
    cudaFree(d_inf);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error(ctx, B200ZK_ERR_CUDA, "copy back failed");
    return B200ZK_OK;
}

int b200zk_encode_points(b200zk_ctx *ctx, int group, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out_bytes) {
    CHECK_CTX(ctx);
    if ((n && (!xy || !out_bytes)) || (group != B200ZK_G1 && group != B200ZK_G2)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad argument");
    USE_DEVICE(ctx);
    if (n == 0) return B200ZK_OK;
    const size_t pb = point_bytes(group), ob = compressed ? pb / 2 : pb;
    size_t o_inf = (n * pb + 255) / 256 * 256, o_out = o_inf + (n + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, o_out + n * ob + 256);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, xy, n * pb, cudaMemcpyHostToDevice, ctx->stream));
    if (inf) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_inf, inf, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = codec_encode(ctx, group, s, inf ? (const uint8_t *)(s + o_inf) : nullptr, n, compressed, s + o_out);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_bytes, s + o_out, n * ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

size_t b200zk_bases_len(const b200zk_bases *bases) { return bases ? bases->n : 0; }

void b200zk_bases_free(b200zk_bases *bases) {
    if (!bases) return;
    cudaSetDevice(bases->ctx->device);
    cudaStreamSynchronize(bases->ctx->stream);
    cudaFree(bases->points);
    cudaFree(bases->infinity);
    cudaFree(bases->pre);
    delete bases;
}

// ---------------------------------------------------------------------------------------------------------------- multiexp
int b200zk_set_msm_window(b200zk_ctx *ctx, int window_bits) {
    CHECK_CTX(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    if (window_bits != 0 && (window_bits < 2 || window_bits > 24)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "window bits must be 0 or in [2, 24]");
    ctx->window_override = window_bits;
    return B200ZK_OK;
}

int b200zk_multiexp_dev(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp,
                        const uint8_t *d_density, void *d_out_jacobian, void *d_status) {
    CHECK_CTX(ctx);
    if (!bases || !d_out_jacobian || (n_exp && !d_scalars)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (bases->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bases live on another device");
    USE_DEVICE(ctx);
    return msm_run(ctx, bases, base_offset, d_scalars, n_exp, d_density, d_out_jacobian, d_status, ctx->window_override);
}

int b200zk_multiexp_batch_dev(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp,
                              size_t scalar_stride, const uint8_t *d_density, size_t density_stride, uint32_t batch, void *d_out_jacobians,
                              void *d_status) {
    CHECK_CTX(ctx);
    if (!bases || !d_out_jacobians || (n_exp && !d_scalars) || batch == 0) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument or empty batch");
    if (batch > 1 && (scalar_stride < n_exp || (d_density && density_stride < n_exp)))
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "strides must be at least n_exp");
    if (bases->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bases live on another device");
    USE_DEVICE(ctx);
    MsmBatch b;
    b.K = batch;
    b.scalar_stride = scalar_stride;
    b.density_stride = density_stride;
    return msm_run(ctx, bases, base_offset, d_scalars, n_exp, d_density, d_out_jacobians, d_status, ctx->window_override, b);
}

int b200zk_multiexp(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                    const uint8_t *density, uint64_t *out_jacobian) {
    CHECK_CTX(ctx);
    if (!bases || !out_jacobian || (n_exp && !scalars)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (bases->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bases live on another device");
    USE_DEVICE(ctx);
    const size_t jac_bytes = bases->group == B200ZK_G1 ? 144 : 288;
    // staging: scalars | density | result | status
    size_t sc_bytes = n_exp * 32, den_bytes = density ? n_exp : 0;
    size_t o_den = (sc_bytes + 255) / 256 * 256, o_res = o_den + (den_bytes + 255) / 256 * 256, o_st = o_res + 512, total = o_st + 256;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, total);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    if (sc_bytes) B200ZK_CUDA(ctx, cudaMemcpyAsync(s, scalars, sc_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (den_bytes) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_den, density, den_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = msm_run(ctx, bases, base_offset, s, n_exp, density ? (const uint8_t *)(s + o_den) : nullptr, s + o_res, s + o_st, ctx->window_override);
    if (rc) return rc;
    uint32_t status = 0;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_jacobian, s + o_res, jac_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(&status, s + o_st, 4, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (status == B200ZK_ERR_UNEXPECTED_IDENTITY) return set_error(ctx, status, "UnexpectedIdentity: a base at infinity was consumed");
    if (status == B200ZK_ERR_UNEXPECTED_EOF) return set_error(ctx, status, "IoError(UnexpectedEof): expected more bases from source");
    if (status == B200ZK_ERR_BAD_ARG) return set_error(ctx, status, "an exponent is not a canonical FrRepr (it has 256 significant bits)");
    return B200ZK_OK;
}

static int status_to_error(b200zk_ctx *ctx, uint32_t status) {
    if (status == B200ZK_ERR_BAD_ARG) return set_error(ctx, status, "an exponent is not a canonical FrRepr (it has 256 significant bits)");
    if (status == B200ZK_ERR_UNEXPECTED_IDENTITY) return set_error(ctx, status, "UnexpectedIdentity: a base at infinity was consumed");
    if (status == B200ZK_ERR_UNEXPECTED_EOF) return set_error(ctx, status, "IoError(UnexpectedEof): expected more bases from source");
    return B200ZK_OK;
}

}  // extern "C"

namespace b200zk {
static __global__ void k_first_status(const char *__restrict__ recs, size_t n, uint32_t *__restrict__ out) {
    uint32_t st = 0;
    for (size_t i = 0; i < n && st == 0; i++) st = *reinterpret_cast<const uint32_t *>(recs + i * REC_BYTES + REC_STATUS);
    *out = st;
}
int first_status(Ctx *ctx, cudaStream_t st, const void *d_records, size_t n, void *d_status_out) {
    k_first_status<<<1, 1, 0, st>>>((const char *)d_records, n, (uint32_t *)d_status_out);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

int multiexp_enqueue(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp, const uint8_t *density,
                     void *d_record_out, int *slot_out) {
    CHECK_CTX(ctx);
    if (!bases || !slot_out || (n_exp && !scalars)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (bases->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bases live on another device");
    USE_DEVICE(ctx);
    if (!ctx->copy_stream) B200ZK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int si = -1;
    for (int i = 0; i < 4; i++) if (!ctx->slots[i].busy) { si = i; break; }
    if (si < 0) return set_error(ctx, B200ZK_ERR_BAD_ARG, "too many multiexp jobs in flight on this context (max 4): wait for one first");
    Ctx::JobSlot &sl = ctx->slots[si];
    const size_t sc_bytes = n_exp * 32, den_bytes = density ? n_exp : 0;
    const size_t o_den = (sc_bytes + 255) / 256 * 256, o_res = o_den + (den_bytes + 255) / 256 * 256, total = o_res + 512;
    if (sl.bytes < total) {
        if (sl.dev) { B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); B200ZK_CUDA(ctx, cudaFree(sl.dev)); sl.dev = nullptr; sl.bytes = 0; }
        B200ZK_CUDA(ctx, cudaMalloc(&sl.dev, total));
        sl.bytes = total;
    }
    if (!sl.host_res) {
        B200ZK_CUDA(ctx, cudaMallocHost(&sl.host_res, 512));
        B200ZK_CUDA(ctx, cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
        B200ZK_CUDA(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    char *d = (char *)sl.dev;
    if (sc_bytes) B200ZK_CUDA(ctx, cudaMemcpyAsync(d, scalars, sc_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (den_bytes) B200ZK_CUDA(ctx, cudaMemcpyAsync(d + o_den, density, den_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    B200ZK_CUDA(ctx, cudaEventRecord(sl.copied, ctx->copy_stream));
    B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, sl.copied, 0));
    char *rec = d_record_out ? (char *)d_record_out : d + o_res;
    int rc = msm_run(ctx, bases, base_offset, d, n_exp, density ? (const uint8_t *)(d + o_den) : nullptr, rec, rec + REC_STATUS, ctx->window_override);
    if (rc) return rc;
    sl.o_res = o_res;
    sl.busy = true;
    *slot_out = si;
    return B200ZK_OK;
}
}  // namespace b200zk

extern "C" {

static int allgather_records(b200zk_ctx *ctx, int group, void *d_record_out);

static int multiexp_async_impl(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                               const uint8_t *density, b200zk_job **job, bool gather) {
    CHECK_CTX(ctx);
    if (!job) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    USE_DEVICE(ctx);
    const bool sharded = gather && ctx->world > 1;
    if (sharded && !ctx->nccl_comm) return set_error(ctx, B200ZK_ERR_NCCL, "communicator not initialised");
    int si = -1;
    // sharded: the window-combine kernel writes this rank's record straight into the send slot of the gather buffer
    int rc = multiexp_enqueue(ctx, bases, base_offset, scalars, n_exp, density, sharded ? ctx->gather_buf : nullptr, &si);
    if (rc) return rc;
    Ctx::JobSlot &sl = ctx->slots[si];
    char *rec = (char *)sl.dev + sl.o_res;
    if (sharded && (rc = allgather_records(ctx, bases->group, rec))) { sl.busy = false; return rc; }
    B200ZK_CUDA(ctx, cudaMemcpyAsync(sl.host_res, rec, REC_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaEventRecord(sl.done, ctx->stream));
    *job = new b200zk_job{ctx, si, bases->group};
    return B200ZK_OK;
}

int b200zk_multiexp_async(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                          const uint8_t *density, b200zk_job **job) {
    return multiexp_async_impl(ctx, bases, base_offset, scalars, n_exp, density, job, false);
}
int b200zk_multiexp_sharded_async(b200zk_ctx *ctx, const b200zk_bases *bases, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                                  const uint8_t *density, b200zk_job **job) {
    return multiexp_async_impl(ctx, bases, base_offset, scalars, n_exp, density, job, true);
}

int b200zk_job_wait(b200zk_job *job, uint64_t *out_jacobian) {
    if (!job) return B200ZK_ERR_BAD_ARG;
    b200zk_ctx *ctx = job->ctx;
    Ctx::JobSlot &sl = ctx->slots[job->slot];
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaEventSynchronize(sl.done);  // blocking wait outside the context lock: other threads may submit meanwhile
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    uint32_t status = *(const uint32_t *)((const char *)sl.host_res + REC_STATUS);
    if (e == cudaSuccess && out_jacobian) memcpy(out_jacobian, sl.host_res, job->group == B200ZK_G1 ? 144 : 288);
    sl.busy = false;
    delete job;
    if (e != cudaSuccess) return set_error(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return status_to_error(ctx, status);
}

int b200zk_sum_points_dev(b200zk_ctx *ctx, int group, const void *d_points, size_t n, void *d_out) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    return msm_sum_points(ctx, group, d_points, n, d_out);
}

int b200zk_into_affine(b200zk_ctx *ctx, int group, const uint64_t *jacobian, size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    CHECK_CTX(ctx);
    if (group != B200ZK_G1 && group != B200ZK_G2) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
    USE_DEVICE(ctx);
    const size_t jb = group == B200ZK_G1 ? 144 : 288, ab = point_bytes(group);
    size_t o_out = (n * jb + 255) / 256 * 256, o_inf = o_out + (n * ab + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, o_inf + n + 256);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, jacobian, n * jb, cudaMemcpyHostToDevice, ctx->stream));
    rc = msm_into_affine(ctx, group, s, n, s + o_out, (uint8_t *)(s + o_inf));
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_xy, s + o_out, n * ab, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_inf) B200ZK_CUDA(ctx, cudaMemcpyAsync(out_inf, s + o_inf, n, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_fixed_base_mul_dev(b200zk_ctx *ctx, int group, const uint64_t *base_affine_host, const void *d_scalars, size_t n, uint32_t scalar_bits,
                              void *d_out_affine, uint8_t *d_out_inf) {
    CHECK_CTX(ctx);
    if (group != B200ZK_G1 && group != B200ZK_G2) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
    USE_DEVICE(ctx);
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, 256);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, base_affine_host, point_bytes(group), cudaMemcpyHostToDevice, ctx->stream));
    return msm_fixed_base(ctx, group, ctx->scratch, d_scalars, n, scalar_bits, d_out_affine, d_out_inf);
}

// ---------------------------------------------------------------------------------------------------------------- multi-GPU
int b200zk_nccl_unique_id(uint8_t out_id[128]) {
    std::string err;
    if (!load_nccl(err)) return B200ZK_ERR_NCCL;
    NcclId id;
    if (g_nccl.GetUniqueId(&id) != 0) return B200ZK_ERR_NCCL;
    memcpy(out_id, id.internal, 128);
    return B200ZK_OK;
}

int b200zk_comm_init(b200zk_ctx *ctx, const uint8_t unique_id[128], int rank, int world) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    std::string err;
    if (!load_nccl(err)) return set_error(ctx, B200ZK_ERR_NCCL, err);
    if (world < 1 || rank < 0 || rank >= world) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad rank/world");
    NcclId id;
    memcpy(id.internal, unique_id, 128);
    int r = g_nccl.CommInitRank(&ctx->nccl_comm, world, id, rank);
    if (r != 0) return set_error(ctx, B200ZK_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    ctx->rank = rank;
    ctx->world = world;
    B200ZK_CUDA(ctx, cudaMalloc(&ctx->gather_buf, ((size_t)world + 1) * REC_BYTES));
    B200ZK_CUDA(ctx, cudaMemset(ctx->gather_buf, 0, ((size_t)world + 1) * REC_BYTES));
    return B200ZK_OK;
}

// All ranks: send record (gather_buf[0], already written) -> gathered records -> sum of the points and the first non-zero status
// in rank order (an UnexpectedIdentity / EOF of any shard fails the multiexp on every rank) -> d_record_out.
static int allgather_records(b200zk_ctx *ctx, int group, void *d_record_out) {
    char *send = (char *)ctx->gather_buf, *recv = send + REC_BYTES;
    int r = g_nccl.AllGather(send, recv, REC_BYTES, /* ncclUint8 */ 1, ctx->nccl_comm, ctx->stream);
    if (r != 0) return set_error(ctx, B200ZK_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    int rc = msm_sum_points(ctx, group, recv, (size_t)ctx->world, d_record_out, REC_BYTES);
    if (rc) return rc;
    return first_status(ctx, ctx->stream, recv, (size_t)ctx->world, (char *)d_record_out + REC_STATUS);
}

int b200zk_allgather_sum_dev(b200zk_ctx *ctx, int group, const void *d_partial, void *d_total) {
    CHECK_CTX(ctx);
    if (group != B200ZK_G1 && group != B200ZK_G2) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
    USE_DEVICE(ctx);
    const size_t jb = group == B200ZK_G1 ? 144 : 288;
    if (ctx->world == 1 || !ctx->nccl_comm) {
        if (ctx->world != 1) return set_error(ctx, B200ZK_ERR_NCCL, "communicator not initialised");
        if (d_total != d_partial) B200ZK_CUDA(ctx, cudaMemcpyAsync(d_total, d_partial, jb, cudaMemcpyDeviceToDevice, ctx->stream));
        return B200ZK_OK;
    }
    char *send = (char *)ctx->gather_buf, *recv = send + REC_BYTES;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(send, d_partial, jb, cudaMemcpyDeviceToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemsetAsync(send + REC_STATUS, 0, 4, ctx->stream));
    int r = g_nccl.AllGather(send, recv, REC_BYTES, /* ncclUint8 */ 1, ctx->nccl_comm, ctx->stream);
    if (r != 0) return set_error(ctx, B200ZK_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return msm_sum_points(ctx, group, recv, (size_t)ctx->world, d_total, REC_BYTES);
}

// ---------------------------------------------------------------------------------------------------------------- domain
int b200zk_ntt_dev(b200zk_ctx *ctx, void *d_coeffs, uint32_t log_m, int kind) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    return ntt_run(ctx, d_coeffs, log_m, kind);
}

int b200zk_ntt(b200zk_ctx *ctx, uint64_t *coeffs_host, uint32_t log_m, int kind) {
    CHECK_CTX(ctx);
    if (log_m >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");
    USE_DEVICE(ctx);
    const size_t bytes = ((size_t)1 << log_m) * 32;
    int rc = ensure_scratch(ctx, &ctx->scratch2, &ctx->scratch2_bytes, bytes);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch2, coeffs_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = ntt_run(ctx, ctx->scratch2, log_m, kind);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(coeffs_host, ctx->scratch2, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

static int stage_small(b200zk_ctx *ctx, const uint64_t v[4], void **dptr) {
    // a 32-byte constant staged at the end of the gather buffer area (own small allocation, stream ordered)
    if (!ctx->small_slot) B200ZK_CUDA(ctx, cudaMalloc(&ctx->small_slot, 256));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->small_slot, v, 32, cudaMemcpyHostToDevice, ctx->stream));
    *dptr = ctx->small_slot;
    return B200ZK_OK;
}

int b200zk_distribute_powers_dev(b200zk_ctx *ctx, void *d_coeffs, size_t n, const uint64_t g[4]) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    void *dg;
    int rc = stage_small(ctx, g, &dg);
    if (rc) return rc;
    rc = ntt_distribute_powers(ctx, d_coeffs, n, dg);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `g` slot is reused by the next call
    return B200ZK_OK;
}

int b200zk_divide_by_z_on_coset_dev(b200zk_ctx *ctx, void *d_coeffs, uint32_t log_m) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    return ntt_divide_by_z_on_coset(ctx, d_coeffs, log_m);
}

int b200zk_domain_z(b200zk_ctx *ctx, const uint64_t tau[4], uint32_t log_m, uint64_t out[4]) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    void *d;
    int rc = stage_small(ctx, tau, &d);
    if (rc) return rc;
    rc = ntt_domain_z(ctx, d, log_m, (char *)d + 64);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, (char *)d + 64, 32, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_fr_scale_dev(b200zk_ctx *ctx, void *d_coeffs, size_t n, const uint64_t s[4]) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    void *ds;
    int rc = stage_small(ctx, s, &ds);
    if (rc) return rc;
    rc = launch_fr_scale(ctx, d_coeffs, ds, n);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_fr_spmv_dev(b200zk_ctx *ctx, const void *d_row_ptr, const void *d_col, const void *d_val, const void *d_x, size_t n_rows, void *d_y) {
    CHECK_CTX(ctx);
    if (n_rows && (!d_row_ptr || !d_y)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    USE_DEVICE(ctx);
    return launch_fr_spmv(ctx, d_row_ptr, d_col, d_val, d_x, n_rows, d_y);
}

int b200zk_field_vec_dev(b200zk_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    return launch_field_vec(ctx, field, op, d_a, d_b, d_out, n);
}

int b200zk_field_vec(b200zk_ctx *ctx, int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    CHECK_CTX(ctx);
    if (field != B200ZK_FR && field != B200ZK_FQ && field != B200ZK_FQ2 && field != B200ZK_FQ2_PAIR) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad field");
    USE_DEVICE(ctx);
    const size_t ob = field == B200ZK_FR ? 32 : field == B200ZK_FQ ? 48 : 96;  // bytes per output element
    const size_t eb = op == B200ZK_OP_MULSUB ? 2 * ob : ob;                      // MULSUB reads pairs
    const size_t bytes = (n * eb + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, 3 * bytes + 256);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(s, a, n * eb, cudaMemcpyHostToDevice, ctx->stream));
    if (n && b) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + bytes, b, n * eb, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_field_vec(ctx, field, op, s, b ? s + bytes : s, s + 2 * bytes, n);
    if (rc) return rc;
    if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(out, s + 2 * bytes, n * ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_point_op(b200zk_ctx *ctx, int group, int op, const uint64_t *a, const uint64_t *b, const uint8_t *b_inf, uint64_t *out, size_t n) {
    CHECK_CTX(ctx);
    if (group != B200ZK_G1 && group != B200ZK_G2) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
    if (op < B200ZK_POINT_DOUBLE || op > B200ZK_POINT_ADD_MIXED) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad point op");
    USE_DEVICE(ctx);
    const size_t jb = group == B200ZK_G1 ? 144 : 288, ab = point_bytes(group);
    const size_t bb = op == B200ZK_POINT_ADD ? jb : ab;
    size_t o_b = (n * jb + 255) / 256 * 256, o_inf = o_b + (n * jb + 255) / 256 * 256, o_out = o_inf + (n + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, o_out + n * jb + 256);
    if (rc) return rc;
    char *s = (char *)ctx->scratch;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, a, n * jb, cudaMemcpyHostToDevice, ctx->stream));
    if (op != B200ZK_POINT_DOUBLE) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_b, b, n * bb, cudaMemcpyHostToDevice, ctx->stream));
    if (op == B200ZK_POINT_ADD_MIXED && b_inf) B200ZK_CUDA(ctx, cudaMemcpyAsync(s + o_inf, b_inf, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_point_op(ctx, group, op, s, s + o_b, (op == B200ZK_POINT_ADD_MIXED && b_inf) ? (const uint8_t *)(s + o_inf) : nullptr, s + o_out, n);
    if (rc) return rc;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, s + o_out, n * jb, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_h_poly_dev(b200zk_ctx *ctx, void *d_a, void *d_b, void *d_c, uint32_t log_m, void *d_out) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (log_m >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");
    return ntt_h_poly(ctx, d_a, d_b, d_c, log_m, d_out);
}

int b200zk_ntt_plan(uint32_t log_m, int large_from, int sm_count, uint32_t batch, uint32_t *stages, uint32_t *columns_log, int *radix4) {
    if (!stages || !columns_log || !radix4 || log_m < 3 || log_m > 30 || batch == 0 || sm_count <= 0) return 0;
    return (int)ntt_plan(log_m, large_from, sm_count, batch, stages, columns_log, radix4);
}

int b200zk_h_poly(b200zk_ctx *ctx, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint32_t log_m, uint64_t *out) {
    CHECK_CTX(ctx);
    if (log_m >= 32) return set_error(ctx, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");
    USE_DEVICE(ctx);
    const size_t bytes = ((size_t)1 << log_m) * 32;
    int rc = ensure_scratch(ctx, &ctx->scratch2, &ctx->scratch2_bytes, 4 * bytes);
    if (rc) return rc;
    char *s = (char *)ctx->scratch2;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s + bytes, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s + 2 * bytes, c, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = ntt_h_poly_batch(ctx, s, log_m, s + 3 * bytes, 1);  // a | b | c contiguous: the three transforms of a step in one set of launches
    if (rc) return rc;
    if (bytes > 32) B200ZK_CUDA(ctx, cudaMemcpyAsync(out, s + 3 * bytes, bytes - 32, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

// ---------------------------------------------------------------------------------------------------------------- groth16
int b200zk_crs_create(b200zk_ctx *ctx, const b200zk_bases *h, const b200zk_bases *l, const b200zk_bases *a, const b200zk_bases *b_g1,
                      const b200zk_bases *b_g2, const uint64_t alpha_g1[12], const uint64_t beta_g1[12], const uint64_t beta_g2[24],
                      const uint64_t delta_g1[12], const uint64_t delta_g2[24], const uint8_t *vk_infinity, b200zk_crs **out) {
    CHECK_CTX(ctx);
    if (!out || !h || !l || !a || !b_g1 || !b_g2) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (h->group != B200ZK_G1 || l->group != B200ZK_G1 || a->group != B200ZK_G1 || b_g1->group != B200ZK_G1 || b_g2->group != B200ZK_G2)
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "query vector in the wrong group");
    USE_DEVICE(ctx);
    b200zk_crs *c = new b200zk_crs();
    c->ctx = ctx;
    c->h = const_cast<b200zk_bases *>(h); c->l = const_cast<b200zk_bases *>(l); c->a = const_cast<b200zk_bases *>(a);
    c->b_g1 = const_cast<b200zk_bases *>(b_g1); c->b_g2 = const_cast<b200zk_bases *>(b_g2);
    c->vk = c->table_delta_g1 = c->table_delta_g2 = nullptr;
    c->subverted = vk_infinity && (vk_infinity[3] || vk_infinity[4]);
    if (vk_infinity) memcpy(c->vk_inf, vk_infinity, 5);
    const size_t t1 = (size_t)32 * 255 * 192, t2 = (size_t)32 * 255 * 384;
    if (cudaMalloc(&c->vk, 3 * 96 + 2 * 192) != cudaSuccess || cudaMalloc(&c->table_delta_g1, t1) != cudaSuccess ||
        cudaMalloc(&c->table_delta_g2, t2) != cudaSuccess) {
        b200zk_crs_free(c);
        return set_error(ctx, B200ZK_ERR_CUDA, "cudaMalloc(crs) failed");
    }
    char *v = (char *)c->vk;
    cudaMemcpyAsync(v, alpha_g1, 96, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(v + 96, beta_g1, 96, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(v + 192, delta_g1, 96, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(v + 288, beta_g2, 192, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(v + 480, delta_g2, 192, cudaMemcpyHostToDevice, ctx->stream);
    int rc = B200ZK_OK;
    if (!c->subverted) {
        rc = msm_build_table(ctx, B200ZK_G1, v + 192, c->table_delta_g1, 32);
        if (!rc) rc = msm_build_table(ctx, B200ZK_G2, v + 480, c->table_delta_g2, 32);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, B200ZK_ERR_CUDA, "crs upload failed");
    if (rc) { b200zk_crs_free(c); return rc; }
    *out = c;
    return B200ZK_OK;
}

void b200zk_crs_free(b200zk_crs *crs) {
    if (!crs) return;
    cudaSetDevice(crs->ctx->device);
    cudaStreamSynchronize(crs->ctx->stream);
    cudaFree(crs->vk);
    cudaFree(crs->table_delta_g1);
    cudaFree(crs->table_delta_g2);
    if (crs->owns_bases)
        for (Bases *b : {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2}) b200zk_bases_free(static_cast<b200zk_bases *>(b));
    delete crs;
}

int b200zk_groth16_prove(b200zk_ctx *ctx, const b200zk_crs *crs, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_constraints,
                         const uint64_t *inputs, size_t n_inputs, const uint64_t *aux, size_t n_aux, const uint8_t *a_aux_density,
                         const uint8_t *b_input_density, const uint8_t *b_aux_density, const uint64_t r[4], const uint64_t s[4],
                         uint64_t proof_a[12], uint64_t proof_b[24], uint64_t proof_c[12], uint8_t inf_flags[3]) {
    CHECK_CTX(ctx);
    if (!crs || !a || !b || !c || !inputs || (n_aux && !aux) || !a_aux_density && n_aux || !b_input_density || (n_aux && !b_aux_density) || !r || !s ||
        !proof_a || !proof_b || !proof_c)
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (crs->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "CRS lives on another device");
    USE_DEVICE(ctx);
    ProveArgs g{a, b, c, n_constraints, inputs, n_inputs, aux, n_aux, a_aux_density, b_input_density, b_aux_density, r, s};
    return groth16_prove(ctx, crs, g, proof_a, proof_b, proof_c, inf_flags);
}

int b200zk_groth16_prove_batch(b200zk_ctx *ctx, const b200zk_crs *crs, const b200zk_prove_input *proofs, size_t n_proofs, size_t n_constraints,
                               size_t n_inputs, size_t n_aux, int lockstep, uint64_t *proofs_a, uint64_t *proofs_b, uint64_t *proofs_c,
                               uint8_t *inf_flags) {
    CHECK_CTX(ctx);
    if (!crs || (n_proofs && (!proofs || !proofs_a || !proofs_b || !proofs_c)) || lockstep < 0 || lockstep > 256)
        return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument or bad lockstep");
    if (crs->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "CRS lives on another device");
    USE_DEVICE(ctx);
    size_t group = lockstep ? (size_t)lockstep : 8;
    // a batched multiexp indexes (proof, exponent, window) with 32 bits: large circuits are proved fewer at a time
    const size_t largest = std::max<size_t>(std::max(n_constraints * 2, n_inputs + n_aux), 1);
    group = std::max<size_t>(1, std::min(group, ((size_t)1 << 31) / (largest * 32)));
    std::vector<ProveArgs> args;
    for (size_t first = 0; first < n_proofs; first += group) {
        const size_t K = std::min(group, n_proofs - first);
        args.clear();
        for (size_t k = 0; k < K; k++) {
            const b200zk_prove_input &p = proofs[first + k];
            if (!p.a || !p.b || !p.c || !p.inputs || (n_aux && (!p.aux || !p.a_aux_density || !p.b_aux_density)) || !p.b_input_density || !p.r || !p.s)
                return set_error(ctx, B200ZK_ERR_BAD_ARG, "null pointer in proof " + std::to_string(first + k));
            args.push_back(ProveArgs{p.a, p.b, p.c, n_constraints, p.inputs, n_inputs, p.aux, n_aux, p.a_aux_density, p.b_input_density, p.b_aux_density, p.r, p.s});
        }
        int rc = groth16_prove_batch(ctx, crs, args.data(), (uint32_t)K, proofs_a + 12 * first, proofs_b + 24 * first, proofs_c + 12 * first,
                                     inf_flags ? inf_flags + 3 * first : nullptr);
        if (rc) return rc;
    }
    return B200ZK_OK;
}

// The batch as the outer FFI sees it (librustzcash_sapling_spend_proof, rustzcash.rs:1375-1626, ends in create_random_proof and
// Proof::write into a 192-byte buffer, rustzcash.rs:1556-1601): N assignments in, N x 192 proof bytes out.
int b200zk_groth16_prove_batch_bytes(b200zk_ctx *ctx, const b200zk_crs *crs, const b200zk_prove_input *proofs, size_t n_proofs, size_t n_constraints,
                                     size_t n_inputs, size_t n_aux, int lockstep, uint8_t *out_proofs /* n_proofs x 192 */) {
    CHECK_CTX(ctx);
    if (n_proofs && !out_proofs) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null output");
    if (n_proofs == 0) return B200ZK_OK;
    std::vector<uint64_t> a(12 * n_proofs), b(24 * n_proofs), c(12 * n_proofs);
    std::vector<uint8_t> inf(3 * n_proofs), fa(n_proofs), fb(n_proofs), fc(n_proofs);
    int rc = b200zk_groth16_prove_batch(ctx, crs, proofs, n_proofs, n_constraints, n_inputs, n_aux, lockstep, a.data(), b.data(), c.data(), inf.data());
    if (rc) return rc;
    for (size_t i = 0; i < n_proofs; i++) { fa[i] = inf[3 * i]; fb[i] = inf[3 * i + 1]; fc[i] = inf[3 * i + 2]; }
    std::vector<uint8_t> ea(48 * n_proofs), eb(96 * n_proofs), ec(48 * n_proofs);
    if ((rc = b200zk_encode_points(ctx, B200ZK_G1, a.data(), fa.data(), n_proofs, 1, ea.data()))) return rc;
    if ((rc = b200zk_encode_points(ctx, B200ZK_G2, b.data(), fb.data(), n_proofs, 1, eb.data()))) return rc;
    if ((rc = b200zk_encode_points(ctx, B200ZK_G1, c.data(), fc.data(), n_proofs, 1, ec.data()))) return rc;
    for (size_t i = 0; i < n_proofs; i++) {  // Proof::write: a (48) | b (96) | c (48), groth16/mod.rs:43-53
        memcpy(out_proofs + 192 * i, ea.data() + 48 * i, 48);
        memcpy(out_proofs + 192 * i + 48, eb.data() + 96 * i, 96);
        memcpy(out_proofs + 192 * i + 144, ec.data() + 48 * i, 48);
    }
    return B200ZK_OK;
}

unsigned long long b200zk_launch_count(b200zk_ctx *ctx, int reset) {
    if (!ctx) return 0;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    unsigned long long v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int b200zk_profile_enable(b200zk_ctx *ctx, int on) {
    CHECK_CTX(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    ctx->prof_on = on != 0;
    return B200ZK_OK;
}
int b200zk_profile_read(b200zk_ctx *ctx, double *accumulate_ms, int *launches) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double total = 0;
    int n = 0;
    for (auto &pr : ctx->prof_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { total += ms; n++; }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->prof_events.clear();
    if (accumulate_ms) *accumulate_ms = total;
    if (launches) *launches = n;
    return B200ZK_OK;
}

// ---------------------------------------------------------------------------------------------------------------- calibration
int b200zk_microbench(b200zk_ctx *ctx, int kind, int iters, double *ops_per_s) {
    CHECK_CTX(ctx);
    USE_DEVICE(ctx);
    if (iters < 1 || !ops_per_s) return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad iters");
    const int threads = 256, blocks = ctx->sm_count * 4;
    int rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, (size_t)threads * blocks * 4);
    if (rc) return rc;
    if ((rc = launch_microbench(ctx, kind, iters, blocks, threads, ctx->scratch))) return rc;  // warm-up
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if ((rc = launch_microbench(ctx, kind, iters, blocks, threads, ctx->scratch))) return rc;
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    B200ZK_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    B200ZK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    double per_thread = kind <= 3 ? 64.0 * iters : 2.0 * iters;
    *ops_per_s = per_thread * threads * blocks / (ms * 1e-3);
    return B200ZK_OK;
}

}  // extern "C"
