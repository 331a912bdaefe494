// The wire format of the proving key: groth16::Parameters::{read, write} and VerifyingKey::{read, write}
// (bellman/src/groth16/mod.rs:140-212, 252-382).  The framing (field order, big-endian u32 counts, which points may be the
// identity, which are always checked) is host logic; every point is decoded / encoded on the device (codec.cu: big-endian
// uncompressed coordinates <-> Montgomery limbs, curve-equation and subgroup tests when `checked`).
//
//   VerifyingKey:  alpha_g1 | beta_g1 | beta_g2 | gamma_g2 | delta_g1 | delta_g2 | u32 ic_len | ic[...]      (always into_affine(): checked)
//   Parameters:    vk | u32 | h[...] | u32 | l[...] | u32 | a[...] | u32 | b_g1[...] | u32 | b_g2[...]        (checked or not; no identity)
#include <cstring>

#include "internal.h"

using namespace b200zk;

namespace {
struct Cursor {
    const uint8_t *p;
    size_t len, pos = 0;
    bool take(size_t n, const uint8_t **out) {
        if (n > len - pos) return false;
        *out = p + pos;
        pos += n;
        return true;
    }
    bool u32be(uint32_t *v) {
        const uint8_t *q;
        if (!take(4, &q)) return false;
        *v = (uint32_t)q[0] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[2] << 8 | q[3];
        return true;
    }
};
int eof(b200zk_ctx *ctx) { return set_error(ctx, B200ZK_ERR_DECODE, "io::ErrorKind::UnexpectedEof: the parameter bytes end early"); }
void put_u32be(uint8_t *q, uint32_t v) { q[0] = v >> 24; q[1] = v >> 16; q[2] = v >> 8; q[3] = v; }
}  // namespace

extern "C" {

int b200zk_parameters_read(b200zk_ctx *ctx, const uint8_t *bytes, size_t len, int checked, b200zk_crs **out) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (!bytes || !out) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    Cursor c{bytes, len};
    // ---- VerifyingKey::read (mod.rs:160-212): six points, always checked, the identity is accepted for them
    const uint8_t *q[6];
    static const size_t sz[6] = {96, 96, 192, 192, 96, 192};
    for (int i = 0; i < 6; i++)
        if (!c.take(sz[i], &q[i])) return eof(ctx);
    uint8_t g1b[3 * 96], g2b[3 * 192];
    memcpy(g1b, q[0], 96); memcpy(g1b + 96, q[1], 96); memcpy(g1b + 192, q[4], 96);          // alpha_g1, beta_g1, delta_g1
    memcpy(g2b, q[2], 192); memcpy(g2b + 192, q[3], 192); memcpy(g2b + 384, q[5], 192);      // beta_g2, gamma_g2, delta_g2
    uint64_t g1[36], g2[72];
    uint8_t i1[3], i2[3];
    int rc = b200zk_decode_points(ctx, B200ZK_G1, g1b, 3, 1, g1, i1);
    if (rc) return rc;
    if ((rc = b200zk_decode_points(ctx, B200ZK_G2, g2b, 3, 1, g2, i2))) return rc;
    uint32_t ic_len;
    if (!c.u32be(&ic_len)) return eof(ctx);
    const uint8_t *icb;
    if (!c.take((size_t)ic_len * 96, &icb)) return eof(ctx);
    std::vector<uint64_t> ic((size_t)ic_len * 12);
    std::vector<uint8_t> ic_inf(ic_len);
    if (ic_len) {
        if ((rc = b200zk_decode_points(ctx, B200ZK_G1, icb, ic_len, 1, ic.data(), ic_inf.data()))) return rc;
        for (uint32_t i = 0; i < ic_len; i++)
            if (ic_inf[i]) return set_error(ctx, B200ZK_ERR_DECODE, "ic[" + std::to_string(i) + "]: point at infinity");
    }
    // ---- the five query vectors (mod.rs:340-372): decoded straight into resident bases; the identity is an error (:300-304)
    b200zk_bases *vec[5] = {};
    static const int grp[5] = {B200ZK_G1, B200ZK_G1, B200ZK_G1, B200ZK_G1, B200ZK_G2};
    auto fail = [&](int code) { for (b200zk_bases *b : vec) b200zk_bases_free(b); return code; };
    for (int v = 0; v < 5; v++) {
        uint32_t n;
        if (!c.u32be(&n)) return fail(eof(ctx));
        const uint8_t *pts;
        if (!c.take((size_t)n * (grp[v] == B200ZK_G1 ? 96 : 192), &pts)) return fail(eof(ctx));
        static const uint8_t none = 0;
        if ((rc = b200zk_bases_upload_encoded(ctx, grp[v], n ? pts : &none, n, checked, 0, &vec[v]))) return fail(rc);
    }
    const uint8_t vk_inf[5] = {i1[0], i1[1], i2[0], i1[2], i2[2]};  // alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2
    b200zk_crs *crs = nullptr;
    rc = b200zk_crs_create(ctx, vec[0], vec[1], vec[2], vec[3], vec[4], g1, g1 + 12, g2, g1 + 24, g2 + 48, vk_inf, &crs);
    if (rc) return fail(rc);
    crs->owns_bases = true;
    crs->has_full_vk = true;
    memcpy(crs->vk_host, g1, 96); memcpy(crs->vk_host + 12, g1 + 12, 96);                   // alpha_g1, beta_g1
    memcpy(crs->vk_host + 24, g2, 192); memcpy(crs->vk_host + 48, g2 + 24, 192);            // beta_g2, gamma_g2
    memcpy(crs->vk_host + 72, g1 + 24, 96); memcpy(crs->vk_host + 84, g2 + 48, 192);        // delta_g1, delta_g2
    const uint8_t all_inf[6] = {i1[0], i1[1], i2[0], i2[1], i1[2], i2[2]};
    memcpy(crs->vk_host_inf, all_inf, 6);
    crs->ic = std::move(ic);
    *out = crs;
    return B200ZK_OK;
}

size_t b200zk_parameters_size(const b200zk_crs *crs) {
    if (!crs || !crs->has_full_vk) return 0;
    return 3 * 96 + 3 * 192 + 4 + crs->ic.size() / 12 * 96 + 5 * 4 + (crs->h->n + crs->l->n + crs->a->n + crs->b_g1->n) * 96 + crs->b_g2->n * 192;
}

int b200zk_parameters_write(b200zk_ctx *ctx, const b200zk_crs *crs, uint8_t *out, size_t cap) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (!crs || !out) return set_error(ctx, B200ZK_ERR_BAD_ARG, "null argument");
    if (!crs->has_full_vk) return set_error(ctx, B200ZK_ERR_BAD_ARG, "this CRS was not made by b200zk_parameters_read: gamma_g2 and ic are unknown");
    if (cap < b200zk_parameters_size(crs)) return set_error(ctx, B200ZK_ERR_BAD_ARG, "output buffer too small (see b200zk_parameters_size)");
    if (crs->ctx->device != ctx->device) return set_error(ctx, B200ZK_ERR_BAD_ARG, "CRS lives on another device");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    uint8_t *w = out;
    // VerifyingKey::write (mod.rs:140-158)
    static const int vgrp[6] = {B200ZK_G1, B200ZK_G1, B200ZK_G2, B200ZK_G2, B200ZK_G1, B200ZK_G2};
    static const int voff[6] = {0, 12, 24, 48, 72, 84};
    int rc;
    for (int i = 0; i < 6; i++) {
        if ((rc = b200zk_encode_points(ctx, vgrp[i], crs->vk_host + voff[i], crs->vk_host_inf + i, 1, 0, w))) return rc;
        w += vgrp[i] == B200ZK_G1 ? 96 : 192;
    }
    const size_t n_ic = crs->ic.size() / 12;
    put_u32be(w, (uint32_t)n_ic); w += 4;
    if (n_ic && (rc = b200zk_encode_points(ctx, B200ZK_G1, crs->ic.data(), nullptr, n_ic, 0, w))) return rc;
    w += n_ic * 96;
    // the query vectors, encoded from their resident copies (mod.rs:259-282)
    const Bases *vec[5] = {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2};
    for (const Bases *b : vec) {
        put_u32be(w, (uint32_t)b->n); w += 4;
        const size_t pb = b->group == B200ZK_G1 ? 96 : 192;
        if (b->n) {
            if (cudaSetDevice(ctx->device) != cudaSuccess) return set_error(ctx, B200ZK_ERR_CUDA, "cudaSetDevice failed");
            rc = ensure_scratch(ctx, &ctx->scratch, &ctx->scratch_bytes, b->n * pb + 256);
            if (rc) return rc;
            if ((rc = codec_encode(ctx, b->group, b->points, b->infinity, b->n, 0, ctx->scratch))) return rc;
            B200ZK_CUDA(ctx, cudaMemcpyAsync(w, ctx->scratch, b->n * pb, cudaMemcpyDeviceToHost, ctx->stream));
            B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        w += b->n * pb;
    }
    return B200ZK_OK;
}

// The VerifyingKey of a CRS made by b200zk_parameters_read: 108 words (alpha_g1 | beta_g1 | beta_g2 | gamma_g2 | delta_g1 |
// delta_g2, affine Montgomery), six infinity flags, and ic (12 words each; *n_ic receives the count, ic may be NULL to query it)
int b200zk_crs_verifying_key(const b200zk_crs *crs, uint64_t vk[108], uint8_t vk_inf[6], uint64_t *ic, size_t *n_ic) {
    if (!crs || !crs->has_full_vk) return B200ZK_ERR_BAD_ARG;
    if (vk) memcpy(vk, crs->vk_host, sizeof(crs->vk_host));
    if (vk_inf) memcpy(vk_inf, crs->vk_host_inf, 6);
    if (n_ic) *n_ic = crs->ic.size() / 12;
    if (ic && !crs->ic.empty()) memcpy(ic, crs->ic.data(), crs->ic.size() * 8);
    return B200ZK_OK;
}
// lengths of the h, l, a, b_g1, b_g2 query vectors
int b200zk_crs_query_sizes(const b200zk_crs *crs, size_t sizes[5]) {
    if (!crs || !sizes) return B200ZK_ERR_BAD_ARG;
    const Bases *vec[5] = {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2};
    for (int i = 0; i < 5; i++) sizes[i] = vec[i]->n;
    return B200ZK_OK;
}
// one-time table build for every query vector of the CRS (b200zk_bases_precompute on each)
int b200zk_crs_precompute(b200zk_ctx *ctx, b200zk_crs *crs, int window_bits) {
    if (!ctx || !crs) return B200ZK_ERR_BAD_ARG;
    for (Bases *b : {crs->h, crs->l, crs->a, crs->b_g1, crs->b_g2}) {
        int rc = b200zk_bases_precompute(ctx, static_cast<b200zk_bases *>(b), window_bits);
        if (rc) return rc;
    }
    return B200ZK_OK;
}

}  // extern "C"
