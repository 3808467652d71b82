// G1 instantiation of the Pippenger pipeline (F = Fq; bases 64 B affine, buckets 128 B XYZZ).
#include "msm.cuh"
namespace b200 {
int msm_g1_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G1Xyzz *out_host) {
    return msm_run_impl<Fq>(ctx, d_bases, d_scalars, scalar_size, n, out_host);
}
}  // namespace b200
