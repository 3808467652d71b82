// G2 instantiation of the Pippenger pipeline (msm.cuh).
#include "msm.cuh"
namespace b200 {

int msm_g2_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G2Xyzz *out_host,
               const MsmTableRaw *table) {
    MsmTable<Fq2> t;
    if (table && table->tbl) { t.tbl = (const Affine<Fq2> *)table->tbl; t.n = table->n; t.c = table->c; t.nwin = table->nwin; }
    return msm_run_impl<Fq2>(ctx, d_bases, d_scalars, scalar_size, n, out_host, &t);
}
int msm_g2_precompute(Ctx *ctx, const void *d_pts, u32 n, int c, void *d_tbl) {
    return msm_precompute_table<Fq2>(ctx, (const Affine<Fq2> *)d_pts, n, c, (Affine<Fq2> *)d_tbl);
}
}  // namespace b200
