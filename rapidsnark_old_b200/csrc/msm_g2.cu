// G2 instantiation of the Pippenger pipeline (msm.cuh).
#include "msm.cuh"
namespace b200 {

static MsmTable<Fq2> as_table_g2(const MsmTableRaw *table) {
    MsmTable<Fq2> t;
    if (table && table->tbl) { t.tbl = (const Affine<Fq2> *)table->tbl; t.n = table->n; t.c = table->c; t.nwin = table->nwin; }
    return t;
}
int msm_g2_enqueue(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, int slot,
                   const MsmTableRaw *table, bool reuse_sort, bool tail, int ws, cudaStream_t sort_stream) {
    MsmTable<Fq2> t = as_table_g2(table);
    return msm_enqueue_impl<Fq2>(ctx, d_bases, d_scalars, scalar_size, n, slot, &t, reuse_sort, tail, ws, sort_stream);
}
int msm_g2_collect(Ctx *ctx, int slot, G2Xyzz *out_host) { return msm_collect_impl<Fq2>(ctx, slot, out_host); }
int msm_g2_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G2Xyzz *out_host,
               const MsmTableRaw *table) {
    *out_host = G2Xyzz::zero();
    B200_TRY(msm_g2_enqueue(ctx, d_bases, d_scalars, scalar_size, n, 0, table, false, true, 0, nullptr));
    return msm_g2_collect(ctx, 0, out_host);
}
int msm_g2_precompute(Ctx *ctx, const void *d_pts, u32 n, int c, void *d_tbl) {
    return msm_precompute_table<Fq2>(ctx, (const Affine<Fq2> *)d_pts, n, c, (Affine<Fq2> *)d_tbl);
}
}  // namespace b200
