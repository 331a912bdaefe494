// One host process, several GPUs: the form in which the reference's only product caller uses the path.  A single zcashd /
// librustzcash process reaches create_proof through the C ABI (librustzcash/src/rustzcash.rs:1556) and issues all of a proof's
// multiexps itself (bellman/src/groth16/prover.rs:289-318); it cannot be split into one process per GPU.  A *group* is a set of
// contexts (one per device) behind one handle:
//   b200zk_multi_bases_upload   shards a base vector by contiguous range over the group's devices (multiexp.rs:34-68: the
//                               SourceBuilder's Arc<Vec<G>> becomes one resident slice per GPU)
//   b200zk_multi_multiexp*      cuts the exponent vector where the base cursor crosses a shard boundary (density-aware), runs the
//                               shards side by side on their own streams, and sums the per-shard partial points on the first
//                               device.  With peer access the last kernel of a shard (k_msm_window_combine) stores its 144 / 288
//                               byte partial + status word straight into the combining GPU's memory over NVLink; without it the
//                               record is moved by cudaMemcpyPeerAsync.  No host round trip between the shards and the sum.
// The process-per-GPU form (b200zk_comm_init + b200zk_multiexp_sharded_async, NCCL) stays for torchrun-style launchers.
#include <algorithm>
#include <cstring>

#include "internal.h"

using namespace b200zk;

static constexpr int GROUP_SLOTS = 4;  // multiexps in flight per group (every context has 4 job slots, see Ctx::slots)

struct b200zk_group {
    std::vector<b200zk_ctx *> ctxs;  // one per entry of `devices` (a device may appear twice: two shards on one GPU)
    b200zk_ctx *combine = nullptr;   // a context of its own on the first device: the sum does not queue behind shard 0's next multiexp
    std::vector<int> peer_ok;        // device of ctxs[d] can store into the first device's memory
    char *records = nullptr;         // device 0: GROUP_SLOTS x (n + 1) records; [slot][n] is the summed result
    void *host_res[GROUP_SLOTS] = {};
    cudaEvent_t done[GROUP_SLOTS] = {};
    std::vector<cudaEvent_t> shard_done;  // GROUP_SLOTS x n
    bool busy[GROUP_SLOTS] = {};
    std::mutex mu;
    std::string last_error;
};
struct b200zk_group_bases {
    b200zk_group *g;
    int group;
    size_t n;
    std::vector<b200zk_bases *> shard;
    std::vector<size_t> lo;  // n_dev + 1 boundaries
};
struct b200zk_group_job {
    b200zk_group *g;
    int slot, group;
    std::vector<int> ctx_slot;
};

static int gerr(b200zk_group *g, int code, const std::string &msg) {
    if (g) g->last_error = msg;
    return code;
}
#define G_CUDA(g, call)                                                                                  \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) return gerr(g, B200ZK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

// Pure host logic (no device needed): exponent split points and per-shard base cursors of a sharded multiexp.
// Shard d owns the bases [bounds[d], bounds[d+1]).  Exponent i with a set density byte consumes base base_offset + rank(i)
// (multiexp.rs:174-196).  e_lo[d] = the first exponent whose base lies at or beyond bounds[d]; shard d gets exponents
// [e_lo[d], e_lo[d+1]) and starts its own cursor at local_offset[d].  The LAST shard also receives every exponent beyond the end
// of the base vector, so that it reports UnexpectedEof exactly where the unsharded source would (multiexp.rs:44-46).
int b200zk_multi_plan(const size_t *bounds, int n_dev, size_t base_offset, const uint8_t *density, size_t n_exp, size_t *e_lo /* n_dev + 1 */,
                      size_t *local_offset /* n_dev */) {
    if (!bounds || n_dev < 1 || !e_lo || !local_offset) return B200ZK_ERR_BAD_ARG;
    size_t consumed = 0, i = 0;  // bases consumed by the exponents before i
    for (int d = 0; d < n_dev; d++) {
        if (d > 0) {
            const size_t want = bounds[d] > base_offset ? bounds[d] - base_offset : 0;  // bases that must be consumed before shard d starts
            if (!density) {
                i = std::min(std::max(i, want), n_exp);
                consumed = i;
            } else {
                while (i < n_exp && consumed < want) consumed += density[i++] != 0;
            }
        }
        e_lo[d] = i;
        const size_t cursor = base_offset + consumed;  // global index of the next base
        local_offset[d] = cursor > bounds[d] ? cursor - bounds[d] : 0;
    }
    e_lo[n_dev] = n_exp;
    return B200ZK_OK;
}

int b200zk_init_multi(const int *devices, int n_dev, b200zk_group **out) {
    if (!out || !devices || n_dev < 1 || n_dev > 64) return B200ZK_ERR_BAD_ARG;
    *out = nullptr;
    b200zk_group *g = new b200zk_group();
    int rc = B200ZK_OK;
    for (int d = 0; d < n_dev && !rc; d++) {
        b200zk_ctx *c = nullptr;
        rc = b200zk_init(devices[d], &c);
        if (!rc) g->ctxs.push_back(c);
    }
    if (!rc) rc = b200zk_init(devices[0], &g->combine);
    if (rc) { b200zk_group_destroy(g); return rc; }
    const int dev0 = devices[0];
    g->peer_ok.assign(n_dev, 0);
    for (int d = 0; d < n_dev; d++) {
        if (devices[d] == dev0) { g->peer_ok[d] = 1; continue; }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[d], dev0) == cudaSuccess && can) {
            cudaSetDevice(devices[d]);
            cudaError_t e = cudaDeviceEnablePeerAccess(dev0, 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->peer_ok[d] = 1;
            cudaGetLastError();
        }
    }
    cudaSetDevice(dev0);
    const size_t per_slot = ((size_t)n_dev + 1) * REC_BYTES;
    bool ok = cudaMalloc((void **)&g->records, GROUP_SLOTS * per_slot) == cudaSuccess && cudaMemset(g->records, 0, GROUP_SLOTS * per_slot) == cudaSuccess;
    g->shard_done.assign((size_t)GROUP_SLOTS * n_dev, nullptr);
    for (int s = 0; s < GROUP_SLOTS && ok; s++) {
        ok = cudaMallocHost(&g->host_res[s], 512) == cudaSuccess && cudaEventCreateWithFlags(&g->done[s], cudaEventDisableTiming) == cudaSuccess;
        for (int d = 0; d < n_dev && ok; d++) {
            cudaSetDevice(devices[d]);
            ok = cudaEventCreateWithFlags(&g->shard_done[(size_t)s * n_dev + d], cudaEventDisableTiming) == cudaSuccess;
        }
        cudaSetDevice(dev0);
    }
    if (!ok) { b200zk_group_destroy(g); return B200ZK_ERR_CUDA; }
    *out = g;
    return B200ZK_OK;
}

void b200zk_group_destroy(b200zk_group *g) {
    if (!g) return;
    for (b200zk_ctx *c : g->ctxs) b200zk_sync(c);
    if (g->combine) b200zk_sync(g->combine);
    for (cudaEvent_t e : g->shard_done) if (e) cudaEventDestroy(e);
    for (int s = 0; s < GROUP_SLOTS; s++) {
        if (g->done[s]) cudaEventDestroy(g->done[s]);
        if (g->host_res[s]) cudaFreeHost(g->host_res[s]);
    }
    if (g->records) { cudaSetDevice(g->ctxs.empty() ? 0 : g->ctxs[0]->device); cudaFree(g->records); }
    if (g->combine) b200zk_destroy(g->combine);
    for (b200zk_ctx *c : g->ctxs) b200zk_destroy(c);
    delete g;
}

const char *b200zk_group_last_error(b200zk_group *g) { return g ? g->last_error.c_str() : "null group"; }
int b200zk_group_size(const b200zk_group *g) { return g ? (int)g->ctxs.size() : 0; }
b200zk_ctx *b200zk_group_ctx(b200zk_group *g, int i) { return (g && i >= 0 && i < (int)g->ctxs.size()) ? g->ctxs[i] : nullptr; }
int b200zk_group_peer_access(const b200zk_group *g, int i) { return (g && i >= 0 && i < (int)g->peer_ok.size()) ? g->peer_ok[i] : 0; }

int b200zk_multi_bases_upload(b200zk_group *g, int group, const void *points, size_t n, size_t stride, const uint8_t *infinity, size_t inf_stride,
                              b200zk_group_bases **out) {
    if (!g) return B200ZK_ERR_BAD_ARG;
    if (!out || (group != B200ZK_G1 && group != B200ZK_G2) || (n && !points)) return gerr(g, B200ZK_ERR_BAD_ARG, "bad group / null argument");
    const int nd = (int)g->ctxs.size();
    b200zk_group_bases *gb = new b200zk_group_bases();
    gb->g = g; gb->group = group; gb->n = n;
    gb->lo.resize(nd + 1);
    const size_t base = n / nd, extra = n % nd;  // contiguous, balanced: the first n % n_dev shards hold one base more
    for (int d = 0; d <= nd; d++) gb->lo[d] = (size_t)d * base + std::min<size_t>(d, extra);
    if (inf_stride == 0) inf_stride = 1;
    for (int d = 0; d < nd; d++) {
        b200zk_bases *b = nullptr;
        const size_t lo = gb->lo[d], cnt = gb->lo[d + 1] - lo;
        int rc = b200zk_bases_upload(g->ctxs[d], group, (const char *)points + lo * stride, cnt, stride, infinity ? infinity + lo * inf_stride : nullptr, inf_stride, &b);
        if (rc) {
            gerr(g, rc, b200zk_last_error(g->ctxs[d]));
            b200zk_multi_bases_free(gb);
            return rc;
        }
        gb->shard.push_back(b);
    }
    *out = gb;
    return B200ZK_OK;
}

int b200zk_multi_bases_precompute(b200zk_group *g, b200zk_group_bases *gb, int window_bits) {
    if (!g || !gb) return B200ZK_ERR_BAD_ARG;
    for (size_t d = 0; d < gb->shard.size(); d++) {
        int rc = b200zk_bases_precompute(g->ctxs[d], gb->shard[d], window_bits);
        if (rc) return gerr(g, rc, b200zk_last_error(g->ctxs[d]));
    }
    return B200ZK_OK;
}

size_t b200zk_multi_bases_len(const b200zk_group_bases *gb) { return gb ? gb->n : 0; }

void b200zk_multi_bases_free(b200zk_group_bases *gb) {
    if (!gb) return;
    for (b200zk_bases *b : gb->shard) b200zk_bases_free(b);
    delete gb;
}

int b200zk_multi_multiexp_async(b200zk_group *g, const b200zk_group_bases *gb, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                                const uint8_t *density, b200zk_group_job **job) {
    if (!g) return B200ZK_ERR_BAD_ARG;
    if (!gb || !job || gb->g != g || (n_exp && !scalars)) return gerr(g, B200ZK_ERR_BAD_ARG, "null argument / bases of another group");
    std::lock_guard<std::mutex> lock(g->mu);
    const int nd = (int)g->ctxs.size();
    int slot = -1;
    for (int s = 0; s < GROUP_SLOTS; s++) if (!g->busy[s]) { slot = s; break; }
    if (slot < 0) return gerr(g, B200ZK_ERR_BAD_ARG, "too many multiexp jobs in flight on this group (max 4): wait for one first");
    std::vector<size_t> e_lo(nd + 1), off(nd);
    b200zk_multi_plan(gb->lo.data(), nd, base_offset, density, n_exp, e_lo.data(), off.data());
    char *recs = g->records + (size_t)slot * (nd + 1) * REC_BYTES;
    b200zk_group_job *j = new b200zk_group_job{g, slot, gb->group, std::vector<int>(nd, -1)};
    int rc = B200ZK_OK;
    for (int d = 0; d < nd && !rc; d++) {
        b200zk_ctx *c = g->ctxs[d];
        const size_t cnt = e_lo[d + 1] - e_lo[d];
        char *rec = recs + (size_t)d * REC_BYTES;
        rc = multiexp_enqueue(c, gb->shard[d], off[d], scalars ? scalars + 4 * e_lo[d] : nullptr, cnt, density ? density + e_lo[d] : nullptr,
                              g->peer_ok[d] ? rec : nullptr, &j->ctx_slot[d]);
        if (rc) { gerr(g, rc, b200zk_last_error(c)); break; }
        std::lock_guard<std::recursive_mutex> cl(c->mu);
        cudaSetDevice(c->device);
        if (!g->peer_ok[d]) {  // no direct stores into the first device: move the record with a peer copy on the shard's stream
            Ctx::JobSlot &sl = c->slots[j->ctx_slot[d]];
            cudaError_t e = cudaMemcpyPeerAsync(rec, g->ctxs[0]->device, (char *)sl.dev + sl.o_res, c->device, REC_BYTES, c->stream);
            if (e != cudaSuccess) { rc = gerr(g, B200ZK_ERR_CUDA, cudaGetErrorString(e)); break; }
        }
        cudaError_t e = cudaEventRecord(g->shard_done[(size_t)slot * nd + d], c->stream);
        if (e != cudaSuccess) rc = gerr(g, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc) {
        for (int d = 0; d < nd; d++)
            if (j->ctx_slot[d] >= 0) { b200zk_sync(g->ctxs[d]); std::lock_guard<std::recursive_mutex> cl(g->ctxs[d]->mu); g->ctxs[d]->slots[j->ctx_slot[d]].busy = false; }
        delete j;
        return rc;
    }
    // the sum on the first device, behind every shard's event
    b200zk_ctx *cc = g->combine;
    std::lock_guard<std::recursive_mutex> cl(cc->mu);
    G_CUDA(g, cudaSetDevice(cc->device));
    for (int d = 0; d < nd; d++) G_CUDA(g, cudaStreamWaitEvent(cc->stream, g->shard_done[(size_t)slot * nd + d], 0));
    char *total = recs + (size_t)nd * REC_BYTES;
    if ((rc = msm_sum_points(cc, gb->group, recs, (size_t)nd, total, REC_BYTES))) return gerr(g, rc, cc->last_error);
    if ((rc = first_status(cc, cc->stream, recs, (size_t)nd, total + REC_STATUS))) return gerr(g, rc, cc->last_error);
    G_CUDA(g, cudaMemcpyAsync(g->host_res[slot], total, REC_BYTES, cudaMemcpyDeviceToHost, cc->stream));
    G_CUDA(g, cudaEventRecord(g->done[slot], cc->stream));
    g->busy[slot] = true;
    *job = j;
    return B200ZK_OK;
}

int b200zk_multi_job_wait(b200zk_group_job *job, uint64_t *out_jacobian) {
    if (!job) return B200ZK_ERR_BAD_ARG;
    b200zk_group *g = job->g;
    cudaSetDevice(g->combine->device);
    cudaError_t e = cudaEventSynchronize(g->done[job->slot]);
    std::lock_guard<std::mutex> lock(g->mu);
    const uint32_t status = *(const uint32_t *)((const char *)g->host_res[job->slot] + REC_STATUS);
    if (e == cudaSuccess && out_jacobian) memcpy(out_jacobian, g->host_res[job->slot], job->group == B200ZK_G1 ? 144 : 288);
    for (size_t d = 0; d < g->ctxs.size(); d++) {
        std::lock_guard<std::recursive_mutex> cl(g->ctxs[d]->mu);
        g->ctxs[d]->slots[job->ctx_slot[d]].busy = false;
    }
    g->busy[job->slot] = false;
    delete job;
    if (e != cudaSuccess) return gerr(g, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    if (status == B200ZK_ERR_UNEXPECTED_IDENTITY) return gerr(g, status, "UnexpectedIdentity: a base at infinity was consumed");
    if (status == B200ZK_ERR_UNEXPECTED_EOF) return gerr(g, status, "IoError(UnexpectedEof): expected more bases from source");
    if (status == B200ZK_ERR_BAD_ARG) return gerr(g, status, "an exponent is not a canonical FrRepr (it has 256 significant bits)");
    return B200ZK_OK;
}

int b200zk_multi_multiexp(b200zk_group *g, const b200zk_group_bases *gb, size_t base_offset, const uint64_t *scalars, size_t n_exp,
                          const uint8_t *density, uint64_t *out_jacobian) {
    b200zk_group_job *job = nullptr;
    int rc = b200zk_multi_multiexp_async(g, gb, base_offset, scalars, n_exp, density, &job);
    if (rc) return rc;
    return b200zk_multi_job_wait(job, out_jacobian);
}

// The H block of create_proof (prover.rs:256-287) over the group's GPUs: the a, b and c vectors are independent until the
// pointwise a * b - c (prover.rs:257-266 runs them as three scoped tasks), so each goes to its own GPU (round-robin over the
// group), is transformed there (ifft + coset_fft), and b, c are then copied peer to peer (NVLink) into the first device, which
// finishes (combine, divide by z, icoset_fft, into_repr).  Worth it for large domains (2 x m x 32 B cross the link once).
int b200zk_multi_h_poly(b200zk_group *g, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint32_t log_m, uint64_t *out) {
    if (!g) return B200ZK_ERR_BAD_ARG;
    if (!a || !b || !c || !out) return gerr(g, B200ZK_ERR_BAD_ARG, "null argument");
    if (log_m >= 32) return gerr(g, B200ZK_ERR_DEGREE_TOO_LARGE, "log_m >= Fr::S (32)");
    std::lock_guard<std::mutex> lock(g->mu);
    const int nd = (int)g->ctxs.size();
    const size_t bytes = ((size_t)1 << log_m) * 32;
    const uint64_t *src[3] = {a, b, c};
    b200zk_ctx *owner[3] = {g->ctxs[0], g->ctxs[1 % nd], g->ctxs[2 % nd]};
    b200zk_ctx *c0 = g->ctxs[0];
    // device 0 workspace: a | b | c | out (its own vectors are transformed in place there); other owners: one vector each
    {
        std::lock_guard<std::recursive_mutex> l0(c0->mu);
        G_CUDA(g, cudaSetDevice(c0->device));
        int rc = ensure_scratch(c0, &c0->scratch3, &c0->scratch3_bytes, 4 * bytes);
        if (rc) return gerr(g, rc, c0->last_error);
    }
    char *w0 = (char *)c0->scratch3;
    void *where[3];
    for (int i = 0; i < 3; i++) {
        b200zk_ctx *cx = owner[i];
        std::lock_guard<std::recursive_mutex> lx(cx->mu);
        G_CUDA(g, cudaSetDevice(cx->device));
        if (cx == c0) {
            where[i] = w0 + (size_t)i * bytes;
        } else {
            // two of the three vectors may share a context when the group has two GPUs: slot i of that context's workspace
            int rc = ensure_scratch(cx, &cx->scratch3, &cx->scratch3_bytes, 3 * bytes);
            if (rc) return gerr(g, rc, cx->last_error);
            where[i] = (char *)cx->scratch3 + (size_t)i * bytes;
        }
        G_CUDA(g, cudaMemcpyAsync(where[i], src[i], bytes, cudaMemcpyHostToDevice, cx->stream));
        int rc = ntt_h_poly_front(cx, where[i], log_m);
        if (rc) return gerr(g, rc, cx->last_error);
        if (cx != c0) {  // hand the transformed vector to the first device over the peer link, on the owner's stream
            G_CUDA(g, cudaMemcpyPeerAsync(w0 + (size_t)i * bytes, c0->device, where[i], cx->device, bytes, cx->stream));
            G_CUDA(g, cudaEventRecord(cx->ev_join, cx->stream));
        }
    }
    std::lock_guard<std::recursive_mutex> l0(c0->mu);
    G_CUDA(g, cudaSetDevice(c0->device));
    for (int i = 1; i < 3; i++)
        if (owner[i] != c0) G_CUDA(g, cudaStreamWaitEvent(c0->stream, owner[i]->ev_join, 0));
    int rc = ntt_h_poly_tail(c0, w0, w0 + bytes, w0 + 2 * bytes, log_m, w0 + 3 * bytes);
    if (rc) return gerr(g, rc, c0->last_error);
    if (bytes > 32) G_CUDA(g, cudaMemcpyAsync(out, w0 + 3 * bytes, bytes - 32, cudaMemcpyDeviceToHost, c0->stream));
    G_CUDA(g, cudaStreamSynchronize(c0->stream));
    for (int i = 1; i < 3; i++)
        if (owner[i] != c0) { cudaSetDevice(owner[i]->device); G_CUDA(g, cudaStreamSynchronize(owner[i]->stream)); }
    return B200ZK_OK;
}

}  // extern "C"
