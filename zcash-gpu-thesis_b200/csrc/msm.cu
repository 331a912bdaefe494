// Group dispatch for the MSM pipeline (the kernels live in msm_impl.cuh, instantiated by msm_g1.cu / msm_g2.cu).
#include "internal.h"

namespace b200zk {

#define DECL(SUFFIX)                                                                                                                 \
    int msm_run_##SUFFIX(Ctx *, const Bases *, size_t, const void *, size_t, const uint8_t *, void *, void *, int, MsmBatch);        \
    int msm_fixed_base_##SUFFIX(Ctx *, const void *, const void *, size_t, uint32_t, void *, uint8_t *);                             \
    int msm_into_affine_##SUFFIX(Ctx *, const void *, size_t, void *, uint8_t *);                                                    \
    int msm_sum_points_##SUFFIX(Ctx *, const void *, size_t, void *, size_t);                                                              \
    int msm_build_table_##SUFFIX(Ctx *, const void *, void *, uint32_t);                                                             \
    int msm_precompute_##SUFFIX(Ctx *, Bases *, uint32_t);
DECL(g1)
DECL(g2)

int msm_run(Ctx *ctx, const Bases *bases, size_t base_offset, const void *d_scalars, size_t n_exp, const uint8_t *d_density, void *d_out_jac,
            void *d_status_out, int window_bits, MsmBatch batch) {
    if (bases->group == B200ZK_G1) return msm_run_g1(ctx, bases, base_offset, d_scalars, n_exp, d_density, d_out_jac, d_status_out, window_bits, batch);
    return msm_run_g2(ctx, bases, base_offset, d_scalars, n_exp, d_density, d_out_jac, d_status_out, window_bits, batch);
}
int msm_fixed_base(Ctx *ctx, int group, const void *d_base_affine, const void *d_scalars, size_t n, uint32_t scalar_bits, void *d_out_affine,
                   uint8_t *d_out_inf) {
    if (group == B200ZK_G1) return msm_fixed_base_g1(ctx, d_base_affine, d_scalars, n, scalar_bits, d_out_affine, d_out_inf);
    if (group == B200ZK_G2) return msm_fixed_base_g2(ctx, d_base_affine, d_scalars, n, scalar_bits, d_out_affine, d_out_inf);
    return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
}
int msm_into_affine(Ctx *ctx, int group, const void *d_jac, size_t n, void *d_out_xy, uint8_t *d_out_inf) {
    if (group == B200ZK_G1) return msm_into_affine_g1(ctx, d_jac, n, d_out_xy, d_out_inf);
    if (group == B200ZK_G2) return msm_into_affine_g2(ctx, d_jac, n, d_out_xy, d_out_inf);
    return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
}
int msm_sum_points(Ctx *ctx, int group, const void *d_jac_in, size_t n, void *d_jac_out, size_t stride) {
    if (group == B200ZK_G1) return msm_sum_points_g1(ctx, d_jac_in, n, d_jac_out, stride);
    if (group == B200ZK_G2) return msm_sum_points_g2(ctx, d_jac_in, n, d_jac_out, stride);
    return set_error(ctx, B200ZK_ERR_BAD_ARG, "bad group");
}

int msm_precompute(Ctx *ctx, Bases *bases, uint32_t c) {
    return bases->group == B200ZK_G1 ? msm_precompute_g1(ctx, bases, c) : msm_precompute_g2(ctx, bases, c);
}
int msm_build_table(Ctx *ctx, int group, const void *d_base_affine, void *d_table, uint32_t nwin) {
    if (group == B200ZK_G1) return msm_build_table_g1(ctx, d_base_affine, d_table, nwin);
    return msm_build_table_g2(ctx, d_base_affine, d_table, nwin);
}

}  // namespace b200zk
