// Library context: one CUDA device, one stream, grow-only device workspaces, error text, phase timers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/b200snark.h"
#include "curve.cuh"

namespace b200 {

static const u32 MSM_MAX_BATCH = 1u << 24;   // points per sort batch (entries < 2^32, entry index < 2^31)

enum Phase {
    PH_H2D = 0,
    PH_MSM_SORT,       // digits + histogram + scan + scatter
    PH_MSM_ACCUM,      // bucket accumulation, G1 (dominant kernel: k_msm_accumulate<Fq>)
    PH_MSM_ACCUM_G2,   // bucket accumulation, G2
    PH_MSM_MERGE,      // boundary / hot bucket merge
    PH_MSM_REDUCE,     // weighted bucket reduction + window sums
    PH_MSM_FINAL,      // D2H of window sums + host Horner
    PH_BUILD_AB,       // coefs x witness -> a, b, c
    PH_NTT,            // the six transforms + twists + h combine
    PH_COUNT
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t hstream = nullptr;    // H pipeline (a,b,c build + NTTs) overlapping the witness MSMs
    cudaEvent_t ev_h = nullptr;
    cudaEvent_t ev_xchg = nullptr;     // in-process exchange: "my peer copies of the owners' polynomials are done"
    bool attr_ntt = false, attr_acc_staged = false, attr_ntt_tma = false;
    int opt_ntt_tma = 0;               // 1: NTT passes move their tiles with the TMA engine (k_ntt_pass_tma)   // cudaFuncSetAttribute is per DEVICE: remembered per context
    unsigned long long peer_enabled = 0;   // devices this ctx's device has been given peer access to
    int hi_prio = 0;                   // stream priority of the side / H streams
    cudaStream_t hstream_bc[2] = {nullptr, nullptr};   // b and c transform chains beside a's (opt_h_streams = 3)
    cudaEvent_t ev_h_fork = nullptr, ev_h_join[2] = {nullptr, nullptr};
    int opt_h_streams = 0;             // 0 auto (3 for sharded zkeys, where the H pipeline is the critical path), 1, 3
    static const int MSM_SLOTS = 8;    // in-flight MSM results; each slot owns a side stream and its bucket buffers
    cudaStream_t side[MSM_SLOTS] = {nullptr};   // folding + reduction of slot k's MSM run here (high priority) while
                                                // the next MSMs sort / accumulate on `stream`
    cudaEvent_t ev_acc[MSM_SLOTS] = {nullptr}, ev_done[MSM_SLOTS] = {nullptr}, ev_merge[MSM_SLOTS] = {nullptr};
    bool slot_busy[MSM_SLOTS] = {false};        // ev_done[slot] recorded and not yet waited for
    unsigned sort_readers[2] = {0, 0};          // per sort workspace: slots whose side-stream merge still reads it
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    int force_c = 0;
    int opt_acc_smem = -1;   // -1 auto, 0 registers, 1 shared memory (experiments)
    int opt_reduce_l_tail = 0;   // 0 = 32: segment length of an MSM's bucket reduction when nothing follows it
    int opt_reduce_l = 0, opt_reduce_l_g2 = 0;   // 0 = automatic segment length of the bucket reduction
    int opt_plane_items = 0;    // 0 = 1024 segments per CTA; > 0: segments each thread of k_msm_plane_sum adds serially before the CTA tree
    int opt_tree_threads = 0;   // 0 = automatic CTA size of the tree-sum kernels (power of two, 32..128)
    int opt_warm_max = 0;    // 0 = default (msm.cuh MSM_WARM_MAX)
    int opt_lockstep_g1 = -1, opt_lockstep_g2 = 0;  // accumulation warps of a CTA in lockstep (msm.cuh k_msm_accumulate): 0 off, 1 on, G1: -1 by size (msm_lockstep_g1), G2: 2 = 256-thread CTAs
    int opt_g2_minb = 0;     // 0 auto; 2 / 3: resident CTAs per SM the G2 accumulation is compiled for (255 / 168 registers)
    int opt_precomp = -1;    // -1 auto (on), 0 off; window bits of resident tables in opt_precomp_c
    int opt_precomp_c = 0;
    float phase_ms[PH_COUNT] = {0};
    struct Seg { int ph, e0, e1; };
    std::vector<cudaEvent_t> evpool;   // phase timing events (reused call after call)
    std::vector<Seg> segs;
    bool phases_pending = false;       // events of the last call recorded but not read back yet (phase_collect_now)
    int opt_timeline = 0;              // 1: phase_collect also keeps (phase, start, end) of every segment, ms from the first
    std::vector<float> timeline;       // triples, see b200_last_timeline
    int ev_used = 0;
    std::vector<DevBuf *> bufs;  // everything to free

    // MSM workspaces (shared by G1/G2 calls; grow-only)
    static const int SORT_WS = 2;       // independent sort workspaces (witness digits / h digits)
    DevBuf w_hist[SORT_WS], w_cursor[SORT_WS], w_entries[SORT_WS], w_hot[SORT_WS], w_scan_totals[SORT_WS], w_plan[SORT_WS], w_tasks[SORT_WS];
    DevBuf w_buckets[MSM_SLOTS], w_partial[MSM_SLOTS], w_segs[MSM_SLOTS], w_win;
    cudaEvent_t ev_sort[SORT_WS] = {nullptr}, ev_ws_acc[SORT_WS] = {nullptr};
    bool ws_acc_pending[SORT_WS] = {false, false};
    int opt_target_tasks_log2 = 0;     // 0 = default (msm.cuh)
    int opt_max_batch_log2 = 0;        // 0 = default 24; smaller values exercise the multi-batch path in tests
    DevBuf w_in_bases, w_in_scalars;   // staging for host-pointer calls
    DevBuf w_ntt;                      // staging for host-pointer NTT calls
    void *pinned = nullptr;            // small pinned host buffer for results
    size_t pinned_cap = 0;

    void *fuse_g1 = nullptr;           // MsmFuse<Fq> of msm_g1.cu: G1 MSMs queued for one fused accumulation launch
    int opt_h_early = -1;              // -1 auto (shards only), 0 / 1: start the H pipeline beside the witness sort instead of after it
    int opt_fuse_g1 = -1;              // -1 auto, 0 one accumulation launch per MSM, 1 fused (see prove.cu)

    struct SlotInfo { int nwin_b = 0, nwin = 0, c = 0, nplanes = 0, L = 0; bool used = false; } slot_info[MSM_SLOTS];

    // NTT twiddle tables (built lazily per log-size)
    struct Twiddles;
    Twiddles *tw = nullptr;
};

#define B200_CUDA_CHECK(ctx, call)                                                            \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            char b__[512];                                                                    \
            snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,           \
                     cudaGetErrorString(e__));                                                \
            (ctx)->err = b__;                                                                 \
            return B200_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define B200_TRY(expr)                   \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != B200_OK) return rc__; \
    } while (0)

inline int ctx_reserve(Ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return B200_OK;
    if (b.p) {
        B200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < Ctx::MSM_SLOTS; i++) if (ctx->side[i]) B200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->side[i]));
        B200_CUDA_CHECK(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;  // a little slack so a slightly larger next call does not realloc
    B200_CUDA_CHECK(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return B200_OK;
}

inline int ctx_pinned(Ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return B200_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    B200_CUDA_CHECK(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
    return B200_OK;
}

// side streams are created on first use (see b200_init)
inline cudaStream_t ctx_side_stream(Ctx *ctx, int slot) {
    if (!ctx->side[slot]) cudaStreamCreateWithPriority(&ctx->side[slot], cudaStreamNonBlocking, ctx->hi_prio);
    return ctx->side[slot];
}

// phase timers: CUDA events on the ctx stream around each phase; collected after the final sync
inline int phase_event(Ctx *ctx, cudaStream_t st) {
    if (ctx->ev_used == (int)ctx->evpool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->evpool.push_back(e);
    }
    cudaEventRecord(ctx->evpool[ctx->ev_used], st);
    return ctx->ev_used++;
}
inline void phase_reset(Ctx *ctx) {
    ctx->phases_pending = false;
    ctx->ev_used = 0;
    ctx->segs.clear();
    for (int i = 0; i < PH_COUNT; i++) ctx->phase_ms[i] = 0.f;
}
inline void phase_begin(Ctx *ctx, Phase ph, cudaStream_t st = nullptr) {
    Ctx::Seg s;
    s.ph = ph;
    s.e0 = phase_event(ctx, st ? st : ctx->stream);
    s.e1 = -1;
    ctx->segs.push_back(s);
}
inline void phase_end(Ctx *ctx, cudaStream_t st = nullptr) { ctx->segs.back().e1 = phase_event(ctx, st ? st : ctx->stream); }
// Reading the phase events back costs a few microseconds per segment (cudaEventElapsedTime): it is instrumentation,
// not part of the call - done lazily, when somebody asks for the numbers (b200_last_phase_ms / b200_last_timeline) or
// the next call starts, unless B200_TIMELINE wants the trace on stderr right away.
inline void phase_collect_now(Ctx *ctx);
inline void phase_collect(Ctx *ctx) {  // all streams must be synchronized
    static const bool timeline = getenv("B200_TIMELINE") != nullptr;
    ctx->phases_pending = true;
    if (timeline) phase_collect_now(ctx);
}
inline void phase_collect_now(Ctx *ctx) {
    static const bool timeline = getenv("B200_TIMELINE") != nullptr;
    if (!ctx->phases_pending) return;
    ctx->phases_pending = false;
    if (ctx->opt_timeline) ctx->timeline.clear();
    for (auto &s : ctx->segs) {
        if (s.e1 < 0) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evpool[s.e0], ctx->evpool[s.e1]) == cudaSuccess) ctx->phase_ms[s.ph] += ms;
        if ((timeline || ctx->opt_timeline) && !ctx->segs.empty()) {
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, ctx->evpool[ctx->segs[0].e0], ctx->evpool[s.e0]);
            cudaEventElapsedTime(&t1, ctx->evpool[ctx->segs[0].e0], ctx->evpool[s.e1]);
            if (timeline) fprintf(stderr, "[timeline] phase %d  %8.3f -> %8.3f ms\n", s.ph, t0, t1);
            if (ctx->opt_timeline) { ctx->timeline.push_back((float)s.ph); ctx->timeline.push_back(t0); ctx->timeline.push_back(t1); }
        }
    }
    ctx->segs.clear();
    ctx->ev_used = 0;
}

// launch bookkeeping: every kernel launch of ours goes through this macro
#define B200_LAUNCH(ctx, kernel, grid, block, smem, ...) B200_LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, smem, __VA_ARGS__)
#define B200_LAUNCH_ON(ctx, st_, kernel, grid, block, smem, ...)                    \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (st_)>>>(__VA_ARGS__);                    \
        (ctx)->launches++;                                                          \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            char b__[512];                                                          \
            snprintf(b__, sizeof b__, "%s:%d: launch %s -> %s", __FILE__, __LINE__, \
                     #kernel, cudaGetErrorString(e__));                             \
            (ctx)->err = b__;                                                       \
            return B200_ERR_CUDA;                                                   \
        }                                                                           \
    } while (0)

void ntt_free_tables(Ctx *ctx);   // ntt.cu
// src/hostmath.cpp (g++): planes[set][0..nplanes-1] = bit-plane sums of S, planes[set][nplanes] = sum of W
void host_msm_finish_g1(const void *planes, int nsets, int nplanes, int L, int nwin, int c, void *out);
void host_msm_finish_g2(const void *planes, int nsets, int nplanes, int L, int nwin, int c, void *out);
// src/hostmath.cpp: blinding of groth16.cpp:209-253 in pieces (see there)
void groth16_blind_prepare(const void *delta1, const void *delta2, const uint8_t *r32, const uint8_t *s32, void *prep640);
void groth16_blind_ab(const void *pi_a128, const void *pib1_128, const void *alpha1, const void *beta1, const void *prep640,
                      const uint8_t *r32, const uint8_t *s32, void *outA64, void *outT128);
void groth16_blind_b(const void *pi_b256, const void *beta2, const void *prep640, void *outB128);
void groth16_blind_c(const void *pi_c128, const void *pih128, const void *T128, const void *prep640, void *outC64);

// msm entry points implemented in msm_g1.cu / msm_g2.cu
struct MsmTableRaw {        // resident per-window table 2^(c*j) * P_i (see msm.cuh); tbl == nullptr: none
    const void *tbl = nullptr;
    u32 n = 0;
    int c = 0, nwin = 0;
};
int msm_table_windows(int c);   // rows of a table for 32-byte scalars
int msm_g1_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G1Xyzz *out_host,
               const MsmTableRaw *table = nullptr);
int msm_g2_run(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, G2Xyzz *out_host,
               const MsmTableRaw *table = nullptr);
// asynchronous pair: enqueue all kernels of one MSM (result lands in pinned slot `slot`), collect = wait + host Horner
// `tail`: no further MSM follows (nothing to overlap the bucket reduction with)
// `ws`: sort workspace (0/1); `sort_stream`: run the digit sort there (e.g. behind the H pipeline) instead of ctx->stream
// `defer`: everything up to the accumulation kernel is enqueued, the kernel itself is launched together with those of
// the other deferred G1 MSMs by msm_g1_flush (one grid for all of them; falls back to an immediate launch when the
// MSM is not a single-batch table MSM)
int msm_g1_enqueue(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, int slot,
                   const MsmTableRaw *table = nullptr, bool reuse_sort = false, bool tail = true, int ws = 0,
                   cudaStream_t sort_stream = nullptr, bool defer = false);
int msm_g1_flush(Ctx *ctx);
void msm_g1_fuse_free(Ctx *ctx);
int msm_g2_enqueue(Ctx *ctx, const void *d_bases, const void *d_scalars, uint32_t scalar_size, uint64_t n, int slot,
                   const MsmTableRaw *table = nullptr, bool reuse_sort = false, bool tail = true, int ws = 0,
                   cudaStream_t sort_stream = nullptr);
int msm_g1_collect(Ctx *ctx, int slot, G1Xyzz *out_host);
int msm_g2_collect(Ctx *ctx, int slot, G2Xyzz *out_host);
int msm_g1_precompute(Ctx *ctx, const void *d_pts, u32 n, int c, void *d_tbl);
int msm_g2_precompute(Ctx *ctx, const void *d_pts, u32 n, int c, void *d_tbl);

}  // namespace b200

// the opaque handle of include/b200snark.h
struct b200_ctx {
    b200::Ctx c;
};
