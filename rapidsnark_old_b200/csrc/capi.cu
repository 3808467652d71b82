// C-ABI entry points (include/b200snark.h): context management and the host-pointer / device-pointer
// front ends of the MSM and NTT primitives.  No CPU fallback exists: without a CUDA device b200_init
// fails with B200_ERR_NO_GPU.
#include "ctx.cuh"

using namespace b200;

static std::string g_init_err;


static const char *k_phase_names[PH_COUNT] = {"h2d", "msm_sort", "msm_accumulate_g1", "msm_accumulate_g2", "msm_merge", "msm_reduce",
                                              "msm_final", "build_abc", "ntt_h"};

extern "C" {

int b200_init(int device, b200_ctx **out) {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);   // only effective if this is the process's first CUDA call

    if (!out) return B200_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_init_err = std::string("b200_init: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
        return B200_ERR_NO_GPU;
    }
    if (device < 0 || device >= ndev) {
        g_init_err = "b200_init: device index out of range";
        return B200_ERR_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_init_err = cudaGetErrorString(e); return B200_ERR_CUDA; }
    b200_ctx *h = new b200_ctx();
    h->c.device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->c.sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_h, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_xchg, cudaEventDisableTiming);
    for (int i = 0; i < Ctx::SORT_WS && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&h->c.ev_sort[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_ws_acc[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) {   // side streams (H pipeline; bucket folding / reduction): highest priority so their CTAs are
                              // placed as soon as an accumulation CTA retires instead of waiting for its grid to drain
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        h->c.hi_prio = hi_p;
        e = cudaStreamCreateWithPriority(&h->c.hstream, cudaStreamNonBlocking, hi_p);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&h->c.ev_h_join[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_h_fork, cudaEventDisableTiming);
        // the per-slot side streams (and the b/c transform streams) are created on first use (ctx_side_stream): the
        // device multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (8 by default), and two of our
        // streams on one queue serialise - measured: the H MSM's bucket folding waited 1.6 ms behind the G2
        // reduction of another stream with 12 streams alive, none with 7
    }
    for (int i = 0; i < Ctx::MSM_SLOTS && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&h->c.ev_acc[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->c.ev_merge[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { g_init_err = cudaGetErrorString(e); delete h; return B200_ERR_CUDA; }
    *out = h;
    return B200_OK;
}

void b200_free(b200_ctx *h) {
    if (!h) return;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) if (c->side[i]) cudaStreamSynchronize(c->side[i]);
    DevBuf *all[] = {&c->w_win, &c->w_in_bases, &c->w_in_scalars, &c->w_ntt};
    for (DevBuf *b : all) if (b->p) cudaFree(b->p);
    for (int i = 0; i < Ctx::SORT_WS; i++) {
        DevBuf *ws[] = {&c->w_hist[i], &c->w_cursor[i], &c->w_entries[i], &c->w_hot[i], &c->w_scan_totals[i], &c->w_plan[i], &c->w_tasks[i]};
        for (DevBuf *b : ws) if (b->p) cudaFree(b->p);
        if (c->ev_sort[i]) cudaEventDestroy(c->ev_sort[i]);
        if (c->ev_ws_acc[i]) cudaEventDestroy(c->ev_ws_acc[i]);
    }
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) {
        if (c->w_buckets[i].p) cudaFree(c->w_buckets[i].p);
        if (c->w_partial[i].p) cudaFree(c->w_partial[i].p);
        if (c->w_segs[i].p) cudaFree(c->w_segs[i].p);
    }
    if (c->pinned) cudaFreeHost(c->pinned);
    for (cudaEvent_t ev : c->evpool) cudaEventDestroy(ev);
    ntt_free_tables(c);
    msm_g1_fuse_free(c);
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) { if (c->ev_acc[i]) cudaEventDestroy(c->ev_acc[i]); if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]); if (c->ev_merge[i]) cudaEventDestroy(c->ev_merge[i]); }
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->hstream) cudaStreamDestroy(c->hstream);
    for (int i = 0; i < 2; i++) { if (c->hstream_bc[i]) cudaStreamDestroy(c->hstream_bc[i]); if (c->ev_h_join[i]) cudaEventDestroy(c->ev_h_join[i]); }
    if (c->ev_h_fork) cudaEventDestroy(c->ev_h_fork);
    if (c->ev_h) cudaEventDestroy(c->ev_h);
    if (c->ev_xchg) cudaEventDestroy(c->ev_xchg);
    cudaStreamDestroy(c->stream);
    delete h;
}

const char *b200_last_error(b200_ctx *h) { return h ? h->c.err.c_str() : g_init_err.c_str(); }
uint64_t b200_launch_count(b200_ctx *h) { return h ? h->c.launches : 0; }
void *b200_stream(b200_ctx *h) { return h ? (void *)h->c.stream : nullptr; }
void b200_set_msm_window(b200_ctx *h, int c_bits) { if (h) h->c.force_c = c_bits; }
int b200_set_option(b200_ctx *h, const char *name, int value) {
    if (!h || !name) return B200_ERR_ARG;
    if (!strcmp(name, "msm_window")) h->c.force_c = value;
    else if (!strcmp(name, "acc_smem")) h->c.opt_acc_smem = value;
    else if (!strcmp(name, "h_streams")) h->c.opt_h_streams = value;
    else if (!strcmp(name, "warm_max")) h->c.opt_warm_max = value;
    else if (!strcmp(name, "tree_threads")) h->c.opt_tree_threads = (value == 32 || value == 64 || value == 128) ? value : 0;
    else if (!strcmp(name, "reduce_l")) h->c.opt_reduce_l = value;
    else if (!strcmp(name, "reduce_l_g2")) h->c.opt_reduce_l_g2 = value;
    else if (!strcmp(name, "reduce_l_tail")) h->c.opt_reduce_l_tail = value;
    else if (!strcmp(name, "plane_items")) h->c.opt_plane_items = value;
    else if (!strcmp(name, "g2_minb")) h->c.opt_g2_minb = value;
    else if (!strcmp(name, "lockstep_g1")) h->c.opt_lockstep_g1 = value;
    else if (!strcmp(name, "lockstep_g2")) h->c.opt_lockstep_g2 = value;
    else if (!strcmp(name, "precomp")) h->c.opt_precomp = value;
    else if (!strcmp(name, "precomp_c")) h->c.opt_precomp_c = value;
    else if (!strcmp(name, "target_tasks_log2")) h->c.opt_target_tasks_log2 = value;
    else if (!strcmp(name, "max_batch_log2")) h->c.opt_max_batch_log2 = value;
    else if (!strcmp(name, "timeline")) h->c.opt_timeline = value;
    else if (!strcmp(name, "fuse_g1")) h->c.opt_fuse_g1 = value;
    else if (!strcmp(name, "ntt_tma")) h->c.opt_ntt_tma = value;
    else if (!strcmp(name, "h_early")) h->c.opt_h_early = value;
    else { h->c.err = std::string("unknown option ") + name; return B200_ERR_ARG; }
    return B200_OK;
}

int b200_last_phase_ms(b200_ctx *h, float *out, int cap) {
    if (!h || !out) return 0;
    phase_collect_now(&h->c);
    int k = cap < PH_COUNT ? cap : PH_COUNT;
    for (int i = 0; i < k; i++) out[i] = h->c.phase_ms[i];
    return k;
}
int b200_last_timeline(b200_ctx *h, float *out, int cap) {
    if (!h || !out) return 0;
    phase_collect_now(&h->c);
    int k = (int)h->c.timeline.size();
    if (k > cap) k = cap - cap % 3;
    for (int i = 0; i < k; i++) out[i] = h->c.timeline[i];
    return k / 3;
}
const char *b200_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? k_phase_names[i] : ""; }

// ---------------------------------------------------------------------------------------------- MSM
static int msm_dev(b200_ctx *h, bool g2, const void *d_bases, const void *d_scalars, uint32_t scalar_size,
                   uint64_t n, void *out) {
    if (!h || !out) return B200_ERR_ARG;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    phase_reset(c);
    int rc;
    if (g2) { G2Xyzz r; rc = msm_g2_run(c, d_bases, d_scalars, scalar_size, n, &r); if (rc == B200_OK) memcpy(out, &r, sizeof r); }
    else    { G1Xyzz r; rc = msm_g1_run(c, d_bases, d_scalars, scalar_size, n, &r); if (rc == B200_OK) memcpy(out, &r, sizeof r); }
    cudaStreamSynchronize(c->stream);
    phase_collect(c);
    return rc;
}

static int msm_host(b200_ctx *h, bool g2, const void *bases, const void *scalars, uint32_t scalar_size,
                    uint64_t n, void *out) {
    if (!h || !out) return B200_ERR_ARG;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    if (n == 0) { memset(out, 0, g2 ? 256 : 128); return B200_OK; }
    if (!bases || !scalars) { c->err = "msm: null input"; return B200_ERR_ARG; }
    if (scalar_size == 0 || scalar_size > 32) { c->err = "msm: scalar_size must be 1..32 bytes"; return B200_ERR_ARG; }
    size_t bsz = (size_t)n * (g2 ? 128 : 64), ssz = (size_t)n * scalar_size;
    B200_TRY(ctx_reserve(c, c->w_in_bases, bsz));
    B200_TRY(ctx_reserve(c, c->w_in_scalars, ssz + 64));
    phase_reset(c);
    phase_begin(c, PH_H2D);
    B200_CUDA_CHECK(c, cudaMemcpyAsync(c->w_in_bases.p, bases, bsz, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA_CHECK(c, cudaMemcpyAsync(c->w_in_scalars.p, scalars, ssz, cudaMemcpyHostToDevice, c->stream));
    phase_end(c);
    int rc;
    if (g2) { G2Xyzz r; rc = msm_g2_run(c, c->w_in_bases.p, c->w_in_scalars.p, scalar_size, n, &r); if (rc == B200_OK) memcpy(out, &r, sizeof r); }
    else    { G1Xyzz r; rc = msm_g1_run(c, c->w_in_bases.p, c->w_in_scalars.p, scalar_size, n, &r); if (rc == B200_OK) memcpy(out, &r, sizeof r); }
    cudaStreamSynchronize(c->stream);
    phase_collect(c);
    return rc;
}

int b200_msm_g1(b200_ctx *h, const void *bases, const void *scalars, uint32_t scalar_size, uint64_t n, void *out) {
    return msm_host(h, false, bases, scalars, scalar_size, n, out);
}
int b200_msm_g2(b200_ctx *h, const void *bases, const void *scalars, uint32_t scalar_size, uint64_t n, void *out) {
    return msm_host(h, true, bases, scalars, scalar_size, n, out);
}
int b200_msm_g1_dev(b200_ctx *h, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, void *out) {
    return msm_dev(h, false, d_bases, d_scalars, scalar_size, n, out);
}
int b200_msm_g2_dev(b200_ctx *h, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, void *out) {
    return msm_dev(h, true, d_bases, d_scalars, scalar_size, n, out);
}

}  // extern "C"
