// G1 instantiation of the Pippenger pipeline (msm.cuh).
#include "msm.cuh"
namespace b200 {
int msm_table_windows(int c) { return (256 + c) / c; }
static MsmTable<Fq> as_table_g1(const MsmTableRaw *table) {
    MsmTable<Fq> t;
    if (table && table->tbl) { t.tbl = (const Affine<Fq> *)table->tbl; t.n = table->n; t.c = table->c; t.nwin = table->nwin; }
    return t;
}
int msm_g1_enqueue(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, int slot,
                   const MsmTableRaw *table, bool reuse_sort, bool tail, int ws, cudaStream_t sort_stream, bool defer) {
    MsmTable<Fq> t = as_table_g1(table);
    MsmFuse<Fq> *fuse = nullptr;
    if (defer && t.tbl && n > 0) {
        if (!ctx->fuse_g1) ctx->fuse_g1 = new MsmFuse<Fq>();
        fuse = (MsmFuse<Fq> *)ctx->fuse_g1;
    }
    return msm_enqueue_impl<Fq>(ctx, d_bases, d_scalars, scalar_size, n, slot, &t, reuse_sort, tail, ws, sort_stream, fuse);
}
int msm_g1_flush(Ctx *ctx) {
    if (!ctx->fuse_g1) return B200_OK;
    return msm_fuse_flush<Fq>(ctx, (MsmFuse<Fq> *)ctx->fuse_g1);
}
void msm_g1_fuse_free(Ctx *ctx) {
    delete (MsmFuse<Fq> *)ctx->fuse_g1;
    ctx->fuse_g1 = nullptr;
}
int msm_g1_collect(Ctx *ctx, int slot, G1Xyzz *out_host) { return msm_collect_impl<Fq>(ctx, slot, out_host); }
int msm_g1_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G1Xyzz *out_host,
               const MsmTableRaw *table) {
    *out_host = G1Xyzz::zero();
    B200_TRY(msm_g1_enqueue(ctx, d_bases, d_scalars, scalar_size, n, 0, table, false, true, 0, nullptr, false));
    return msm_g1_collect(ctx, 0, out_host);
}
int msm_g1_precompute(Ctx *ctx, const void *d_pts, u32 n, int c, void *d_tbl) {
    return msm_precompute_table<Fq>(ctx, (const Affine<Fq> *)d_pts, n, c, (Affine<Fq> *)d_tbl);
}
}  // namespace b200
