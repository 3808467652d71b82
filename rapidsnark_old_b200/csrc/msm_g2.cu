// G2 instantiation of the Pippenger pipeline (F = Fq2; bases 128 B affine, buckets 256 B XYZZ).
#include "msm.cuh"
namespace b200 {
int msm_g2_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G2Xyzz *out_host) {
    return msm_run_impl<Fq2>(ctx, d_bases, d_scalars, scalar_size, n, out_host);
}
}  // namespace b200
