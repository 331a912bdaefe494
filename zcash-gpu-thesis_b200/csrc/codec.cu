// Point (de)serialisation on the device: the wire format of groth16::Parameters::read / Proof::write
// (bellman/src/groth16/mod.rs:43-53, 287-382) = pairing's big-endian encodings
// (pairing/src/bls12_381/ec.rs:686-752 G1 uncompressed decode, :796-868 G1 compressed encode, :2624-2830 the G2 twins;
// format: pairing/src/bls12_381/README.md:59-75).  Fq = 48 bytes big-endian canonical; Fq2 = c1 then c0; top three bits of
// byte 0: 0x80 compressed, 0x40 infinity, 0x20 "y is the lexicographically largest root".
// Decoding converts straight into the Montgomery limbs the kernels use (from_repr = multiply by R^2, fq.rs:740-751),
// checks canonicity, and when `checked` also y^2 = x^3 + b and r * P = identity (ec.rs:125-144) like into_affine().
#include "ec.cuh"
#include "internal.h"

namespace b200zk {

enum { DEC_OK = 0, DEC_COMPRESSED = 1, DEC_UNEXPECTED_INFO = 2, DEC_NOT_IN_FIELD = 3, DEC_NOT_ON_CURVE = 4, DEC_NOT_IN_SUBGROUP = 5, DEC_INFINITY = 6 };

__device__ __forceinline__ bool fq_read_be(const uint8_t *p, uint32_t first_byte_mask, fq_t &out) {
    // 48 big-endian bytes -> 12 little-endian u32 limbs; returns false when the value is >= q
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const uint8_t *q = p + 44 - 4 * j;
        uint32_t b0 = q[0];
        if (j == 11) b0 &= first_byte_mask;
        out.v[j] = (b0 << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    bool lt = false, decided = false;
#pragma unroll
    for (int j = 11; j >= 0; j--) {
        if (!decided && out.v[j] != FqParams::mod(j)) { lt = out.v[j] < FqParams::mod(j); decided = true; }
    }
    return lt;
}
__device__ __forceinline__ void fq_write_be(uint8_t *p, const fq_t &canonical) {
#pragma unroll
    for (int j = 0; j < 12; j++) {
        uint8_t *q = p + 44 - 4 * j;
        uint32_t v = canonical.v[j];
        q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v;
    }
}
// a > b on canonical values (fq.rs:703-708)
__device__ __forceinline__ bool fq_gt(const fq_t &a, const fq_t &b) {
#pragma unroll
    for (int j = 11; j >= 0; j--) {
        if (a.v[j] != b.v[j]) return a.v[j] > b.v[j];
    }
    return false;
}

template <class F> struct Coord;
template <> struct Coord<fq_t> {
    static constexpr int BYTES = 48;
    __device__ static bool read(const uint8_t *p, uint32_t mask, fq_t &out) { fq_t c; bool ok = fq_read_be(p, mask, c); out = c.to_mont(); return ok; }
    __device__ static void write(uint8_t *p, const fq_t &m) { fq_write_be(p, m.from_mont()); }
    __device__ static bool lex_largest(const fq_t &y) { return fq_gt(y.from_mont(), y.neg().from_mont()); }
    __device__ static fq_t curve_b() { fq_t four = fq_t::zero(); four.v[0] = 4; return four.to_mont(); }  // fq.rs:69-76
};
template <> struct Coord<fq2_t> {
    static constexpr int BYTES = 96;
    __device__ static bool read(const uint8_t *p, uint32_t mask, fq2_t &out) {  // c1 then c0
        fq_t c1, c0;
        bool ok1 = fq_read_be(p, mask, c1), ok0 = fq_read_be(p + 48, 0xff, c0);
        out.c0 = c0.to_mont();
        out.c1 = c1.to_mont();
        return ok0 && ok1;
    }
    __device__ static void write(uint8_t *p, const fq2_t &m) { fq_write_be(p, m.c1.from_mont()); fq_write_be(p + 48, m.c0.from_mont()); }
    __device__ static bool lex_largest(const fq2_t &y) {  // fq2.rs:21-30: compare c1 first, then c0
        fq_t a1 = y.c1.from_mont(), b1 = y.c1.neg().from_mont();
        if (a1 != b1) return fq_gt(a1, b1);
        return fq_gt(y.c0.from_mont(), y.c0.neg().from_mont());
    }
    __device__ static fq2_t curve_b() { fq_t four = fq_t::zero(); four.v[0] = 4; four = four.to_mont(); return {four, four}; }  // 4(u + 1), ec.rs:2848-2853
};

// Fr::char() (fr.rs:5-11) for the subgroup check
__device__ __constant__ uint32_t FR_CHAR[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};

template <class F>
__global__ void __launch_bounds__(64) k_decode_uncompressed(const uint8_t *__restrict__ bytes, size_t n, int checked, int allow_infinity,
                                                           Affine<F> *__restrict__ out, uint8_t *__restrict__ out_inf, unsigned long long *__restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int CB = Coord<F>::BYTES;
    const uint8_t *p = bytes + i * 2 * CB;
    int code = DEC_OK;
    Affine<F> a;
    a.x = F::zero();
    a.y = F::one();
    bool inf = false;
    const uint8_t flags = p[0];
    if (flags & 0x80) {
        code = DEC_COMPRESSED;
    } else if (flags & 0x40) {
        bool zero = (flags & 0x3f) == 0;
        for (int k = 1; k < 2 * CB; k++) zero = zero && p[k] == 0;
        if (!zero) code = DEC_UNEXPECTED_INFO;
        else if (!allow_infinity) code = DEC_INFINITY;
        inf = true;
    } else if (flags & 0x20) {
        code = DEC_UNEXPECTED_INFO;
    } else {
        bool okx = Coord<F>::read(p, 0x1f, a.x), oky = Coord<F>::read(p + CB, 0xff, a.y);
        if (!okx || !oky) code = DEC_NOT_IN_FIELD;
        else if (checked) {
            F lhs = a.y.sqr(), rhs = a.x.sqr() * a.x + Coord<F>::curve_b();
            if (lhs != rhs) code = DEC_NOT_ON_CURVE;
            else {
                Jacobian<F> acc = Jacobian<F>::zero();  // r * P, MSB first (ec.rs:87-99, 141-144)
                bool found = false;
                for (int b = 255; b >= 0; b--) {
                    bool bit = (FR_CHAR[b >> 5] >> (b & 31)) & 1;
                    if (found) jacobian_double(acc); else found = bit;
                    if (bit) jacobian_add_mixed(acc, a, false);
                }
                if (!acc.is_zero()) code = DEC_NOT_IN_SUBGROUP;
            }
        }
    }
    out[i] = a;
    out_inf[i] = inf ? 1 : 0;
    if (code != DEC_OK) atomicMin(err, ((unsigned long long)i << 8) | (unsigned long long)code);
}

// mode 0: uncompressed (x || y), mode 1: compressed (x with the sign flag)
template <class F>
__global__ void __launch_bounds__(64) k_encode_points(const Affine<F> *__restrict__ pts, const uint8_t *__restrict__ inf, size_t n, int compressed,
                                                     uint8_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int CB = Coord<F>::BYTES;
    const int sz = compressed ? CB : 2 * CB;
    uint8_t *p = out + i * sz;
    if (inf && inf[i]) {
        for (int k = 0; k < sz; k++) p[k] = 0;
        p[0] = compressed ? 0xC0 : 0x40;
        return;
    }
    Affine<F> a = pts[i];
    Coord<F>::write(p, a.x);
    if (compressed) {
        p[0] |= 0x80;
        if (Coord<F>::lex_largest(a.y)) p[0] |= 0x20;
    } else {
        Coord<F>::write(p + CB, a.y);
    }
}

int codec_decode_uncompressed(Ctx *ctx, int group, const void *d_bytes, size_t n, int checked, int allow_infinity, void *d_out, uint8_t *d_inf,
                              unsigned long long *d_err) {
    if (n == 0) return B200ZK_OK;
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == B200ZK_G1)
        k_decode_uncompressed<fq_t><<<blocks, 64, 0, ctx->stream>>>((const uint8_t *)d_bytes, n, checked, allow_infinity, (g1_affine_t *)d_out, d_inf, d_err);
    else
        k_decode_uncompressed<fq2_t><<<blocks, 64, 0, ctx->stream>>>((const uint8_t *)d_bytes, n, checked, allow_infinity, (g2_affine_t *)d_out, d_inf, d_err);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}
int codec_encode(Ctx *ctx, int group, const void *d_pts, const uint8_t *d_inf, size_t n, int compressed, void *d_out) {
    if (n == 0) return B200ZK_OK;
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == B200ZK_G1) k_encode_points<fq_t><<<blocks, 64, 0, ctx->stream>>>((const g1_affine_t *)d_pts, d_inf, n, compressed, (uint8_t *)d_out);
    else k_encode_points<fq2_t><<<blocks, 64, 0, ctx->stream>>>((const g2_affine_t *)d_pts, d_inf, n, compressed, (uint8_t *)d_out);
    B200ZK_CUDA(ctx, cudaGetLastError());
    return B200ZK_OK;
}

}  // namespace b200zk
