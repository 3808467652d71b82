// Pippenger multi-scalar multiplication for BN254 G1 / G2 on sm_100a, written once over the field F.
//
// Replaces ParallelMultiexp<Curve>::multiexp (depends/ffiasm/c/multiexp.cpp:98-144).  The reference
// gives every OpenMP thread a private bucket set per window and walks the points in order
// (processChunk :36-47, packThreads :49-60, reduce :62-96).  On the GPU the same sum is organised as
//
//   1. k_msm_digits<false>  signed c-bit digits of every scalar (getChunk, multiexp.cpp:22-34, plus
//                           a carry so that buckets only cover |digit| in 1..2^(c-1)); histogram of
//                           (window, |digit|) with one global atomic per non-zero digit
//   2. k_scan_*             exclusive scan of the histogram -> bucket offsets
//   3. k_msm_digits<true>   scatter of (point index | sign) into bucket order: a counting sort, so
//                           the order inside a bucket is arbitrary - the group sum does not care
//   4. k_msm_accumulate     THE hot kernel: the sorted entry list is cut into equal chunks of T
//                           entries, one per thread, regardless of bucket boundaries; each thread
//                           gathers its affine points (64 B / 128 B, read-only path, next point
//                           prefetched while the current mixed add runs) and flushes the running
//                           XYZZ sum whenever the bucket id changes.  Equal chunks = no divergence
//                           in trip count, and skewed witnesses (huge |digit| = 1 buckets) are split
//                           over many threads for free.  Buckets cut by a chunk boundary go to
//                           per-thread head/tail partial slots.
//   5. k_msm_merge / k_msm_merge_hot  fold those partials (a bucket spanning > 32 chunks is folded
//                           by a whole CTA: strided partial sums + shared-memory tree)
//   6. k_msm_reduce_segments / k_msm_window_sum   sum_k k*B_k per window: running sums over
//                           segments of L buckets + small in-thread multiplier, then a CTA tree
//   7. host: Horner over the <= 65 window sums (a serial chain of ~270 group operations is ~8x
//      faster on one CPU core than on one GPU thread)
//
// Zero scalars/digits are skipped like the reference (multiexp.cpp:43); zero bases (0,0) are skipped
// inside the mixed add (multiexp.cpp:40).  Scalars are read as plain 8*scalar_size-bit integers and
// never reduced (multiexp.cpp:118).
#pragma once
#include "ctx.cuh"
#include "memops.cuh"

namespace b200 {

struct MsmGeom {
    int c;          // window bits
    int nwin;       // windows (covers nbits + 1 for the signed-digit carry)
    u32 nbk;        // buckets per window = 2^(c-1)
    u32 NB;         // nwin * nbk
    u32 T;          // entries per accumulate thread
    u32 L;          // buckets per reduce segment
    u32 nseg;       // segments per window
};

static const u32 MSM_MAX_BATCH = 1u << 24;   // points per sort batch (entry index fits 31 bits, entries < 2^32)
static const int MSM_HOT_SPAN = 32;          // chunks; wider buckets are merged by a CTA
static const int MSM_MAX_WIN = 65;

inline MsmGeom msm_geometry(uint64_t n, uint32_t scalar_size, int force_c) {
    MsmGeom g;
    int lg = 0;
    while ((2ull << lg) <= n) lg++;   // floor(log2 n)
    int c = lg - 4;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    if (force_c >= 4 && force_c <= 20) c = force_c;
    int nbits = (int)scalar_size * 8;
    g.c = c;
    g.nwin = (nbits + c) / c;         // ceil((nbits + 1) / c)
    g.nbk = 1u << (c - 1);
    g.NB = (u32)g.nwin * g.nbk;
    g.T = 32;
    g.L = g.nbk >= 16 ? 16 : g.nbk;
    g.nseg = g.nbk / g.L;
    return g;
}

// ------------------------------------------------------------------------------------------------
// 1 + 3: signed digits, histogram / scatter
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const uint8_t *__restrict__ scalars, u32 scalar_size, u32 n,
                                                      int c, int nwin, u32 nbk, u32 *__restrict__ counters,
                                                      u32 *__restrict__ entries) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s[9];
    if (scalar_size == 32) {
        const uint4 *p = reinterpret_cast<const uint4 *>(scalars) + (size_t)i * 2;
        uint4 lo = __ldg(p), hi = __ldg(p + 1);
        s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
        s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = 0;
        const uint8_t *p = scalars + (size_t)i * scalar_size;
        for (u32 k = 0; k < scalar_size; k++) s[k >> 2] |= (u32)p[k] << (8 * (k & 3));
    }
    s[8] = 0;
    u32 any = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) any |= s[k];
    if (any == 0) return;                               // zero scalar: no digits (multiexp.cpp:43)
    const u32 mask = (1u << c) - 1, half = 1u << (c - 1);
    u32 carry = 0;
    for (int w = 0; w < nwin; w++) {
        int o = w * c;
        int word = o >> 5, sh = o & 31;
        u32 raw = 0;
        if (word < 8) {
            u64 two = ((u64)s[word + 1] << 32) | s[word];
            raw = (u32)(two >> sh) & mask;
        }
        raw += carry;
        u32 neg = raw > half;
        u32 mag = neg ? (1u << c) - raw : raw;          // raw == 2^c -> digit 0, carry 1
        carry = neg;
        if (mag != 0) {
            u32 b = (u32)w * nbk + mag - 1;
            if (!SCATTER) {
                atomicAdd(&counters[b], 1u);
            } else {
                u32 pos = atomicAdd(&counters[b], 1u);
                entries[pos] = i | (neg << 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2: exclusive scan of u32 counters (tile = 4096 elements per CTA)
// ------------------------------------------------------------------------------------------------
static const int SCAN_TILE = 4096;

static __global__ void __launch_bounds__(1024) k_scan_tile(u32 *__restrict__ data, u32 *__restrict__ totals) {
    __shared__ u32 warp_sums[32];
    uint4 *p = reinterpret_cast<uint4 *>(data) + (size_t)blockIdx.x * 1024 + threadIdx.x;
    uint4 v = *p;
    u32 t0 = v.x, t1 = t0 + v.y, t2 = t1 + v.z, t3 = t2 + v.w;
    u32 incl = t3;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 ws = warp_sums[lane];
        u32 wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += y;
        }
        warp_sums[lane] = wi - ws;  // exclusive
        if (lane == 31) totals[blockIdx.x] = wi;
    }
    __syncthreads();
    u32 base = warp_sums[wid] + incl - t3;
    uint4 o;
    o.x = base; o.y = base + t0; o.z = base + t1; o.w = base + t2;
    *p = o;
}

static __global__ void __launch_bounds__(1024) k_scan_totals(u32 *__restrict__ totals, u32 ntiles) {
    __shared__ u32 warp_sums[32];
    __shared__ u32 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 base = 0; base < ntiles; base += 1024) {
        u32 i = base + threadIdx.x;
        u32 v = i < ntiles ? totals[i] : 0;
        u32 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            u32 ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        u32 carry = carry_s;
        u32 excl = carry + warp_sums[wid] + incl - v;
        if (i < ntiles) totals[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(1024) k_scan_add(u32 *__restrict__ data, const u32 *__restrict__ totals,
                                                   u32 *__restrict__ copy) {
    uint4 *p = reinterpret_cast<uint4 *>(data) + (size_t)blockIdx.x * 1024 + threadIdx.x;
    u32 add = totals[blockIdx.x];
    uint4 v = *p;
    v.x += add; v.y += add; v.z += add; v.w += add;
    *p = v;
    reinterpret_cast<uint4 *>(copy)[(size_t)blockIdx.x * 1024 + threadIdx.x] = v;  // scatter cursors
}

// ------------------------------------------------------------------------------------------------
// 4: bucket accumulation over equal chunks of the sorted entry list
// ------------------------------------------------------------------------------------------------
template <class F, bool PREFETCH_REGS>
__global__ void __launch_bounds__(128) k_msm_accumulate(const Affine<F> *__restrict__ bases,
                                                          const u32 *__restrict__ entries,
                                                          const u32 *__restrict__ off, u32 NB, u32 T,
                                                          Xyzz<F> *__restrict__ buckets,
                                                          Xyzz<F> *__restrict__ partial, int add_existing) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 E = off[NB];
    const u64 start64 = (u64)t * T;
    if (start64 >= E) return;
    const u32 start = (u32)start64;
    const u32 end = (E - start < T) ? E : start + T;

    // bucket containing `start`: the largest b with off[b] <= start
    u32 lo = 0, hi = NB - 1;
    while (lo < hi) {
        u32 mid = (lo + hi + 1) >> 1;
        if (off[mid] <= start) lo = mid; else hi = mid - 1;
    }
    u32 b = lo;
    u32 b_end = off[b + 1];
    bool head = off[b] < start;      // this bucket began in an earlier chunk

    Xyzz<F> acc = Xyzz<F>::zero();
    u32 e_next = entries[start];
    Affine<F> p_next;
    if (PREFETCH_REGS) p_next = ldg_struct(bases + (e_next & 0x7fffffffu));

    for (u32 pos = start; pos < end; pos++) {
        if (pos >= b_end) {
            // bucket b is finished inside this chunk
            if (head) st_struct(partial + 2 * (size_t)t, acc);
            else {
                if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(acc, old); }
                st_struct(buckets + b, acc);
            }
            head = false;
            acc = Xyzz<F>::zero();
            do { b++; b_end = off[b + 1]; } while (pos >= b_end);
        }
        u32 e = e_next;
        Affine<F> p;
        if (PREFETCH_REGS) p = p_next; else p = ldg_struct(bases + (e & 0x7fffffffu));
        if (pos + 1 < end) {
            e_next = entries[pos + 1];
            if (PREFETCH_REGS) p_next = ldg_struct(bases + (e_next & 0x7fffffffu));
        }
        if (e >> 31) p.y = fneg(p.y);
        ec_madd(acc, p);
    }
    // last bucket of the chunk
    if (head) st_struct(partial + 2 * (size_t)t, acc);                  // spans the whole chunk or ends in it
    else if (b_end > end) st_struct(partial + 2 * (size_t)t + 1, acc);  // continues into the next chunk
    else {
        if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(acc, old); }
        st_struct(buckets + b, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// 5: fold the partials of buckets cut by chunk boundaries
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_msm_merge(const u32 *__restrict__ off, u32 NB, u32 T,
                                                     Xyzz<F> *__restrict__ buckets,
                                                     const Xyzz<F> *__restrict__ partial, int add_existing,
                                                     u32 *__restrict__ hot /* [0] = count, then bucket ids */) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    u32 o0 = off[b], o1 = off[b + 1];
    if (o1 == o0) return;
    u32 t0 = o0 / T, t1 = (o1 - 1) / T;
    if (t0 == t1) return;                      // written directly by the accumulate kernel
    if (t1 - t0 > (u32)MSM_HOT_SPAN) {
        u32 k = atomicAdd(&hot[0], 1u);
        hot[1 + k] = b;
        return;
    }
    Xyzz<F> acc = ld_struct(partial + 2 * (size_t)t0 + 1);
    for (u32 t = t0 + 1; t <= t1; t++) {
        Xyzz<F> q = ld_struct(partial + 2 * (size_t)t);
        ec_add(acc, q);
    }
    if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(acc, old); }
    st_struct(buckets + b, acc);
}

// CTA-wide tree sum of one XYZZ point per thread; result valid in thread 0.  smem: blockDim.x points
template <class F>
DEVFN void cta_tree_sum(Xyzz<F> &acc, Xyzz<F> *sm) {
    for (int stride = blockDim.x >> 1; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x >= stride && (int)threadIdx.x < 2 * stride) st_struct(sm + threadIdx.x, acc);
        __syncthreads();
        if ((int)threadIdx.x < stride) {
            Xyzz<F> q = ld_struct(sm + threadIdx.x + stride);
            ec_add(acc, q);
        }
        __syncthreads();
    }
}

template <class F>
__global__ void __launch_bounds__(128) k_msm_merge_hot(const u32 *__restrict__ off, u32 T,
                                                         Xyzz<F> *__restrict__ buckets,
                                                         const Xyzz<F> *__restrict__ partial, int add_existing,
                                                         const u32 *__restrict__ hot) {
    extern __shared__ uint4 smem_raw[];
    Xyzz<F> *sm = reinterpret_cast<Xyzz<F> *>(smem_raw);
    u32 nhot = hot[0];
    for (u32 h = blockIdx.x; h < nhot; h += gridDim.x) {
        u32 b = hot[1 + h];
        u32 o0 = off[b], o1 = off[b + 1];
        u32 t0 = o0 / T, t1 = (o1 - 1) / T;
        Xyzz<F> acc = Xyzz<F>::zero();
        if (threadIdx.x == 0) acc = ld_struct(partial + 2 * (size_t)t0 + 1);
        for (u32 t = t0 + 1 + threadIdx.x; t <= t1; t += blockDim.x) {
            Xyzz<F> q = ld_struct(partial + 2 * (size_t)t);
            ec_add(acc, q);
        }
        cta_tree_sum(acc, sm);
        if (threadIdx.x == 0) {
            if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(acc, old); }
            st_struct(buckets + b, acc);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// 6: sum_k (k+1) * B[w][k]  (reference: reduce, multiexp.cpp:62-96 - different recursion, same value)
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_msm_reduce_segments(const Xyzz<F> *__restrict__ buckets, u32 nbk, u32 L,
                                                               u32 nseg, u32 total_segs,
                                                               Xyzz<F> *__restrict__ seg_out) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_segs) return;
    u32 w = g / nseg, s = g % nseg;
    const Xyzz<F> *B = buckets + (size_t)w * nbk + (size_t)s * L;
    Xyzz<F> run = Xyzz<F>::zero(), acc = Xyzz<F>::zero();
    for (int k = (int)L - 1; k >= 0; k--) {
        Xyzz<F> q = ld_struct(B + k);
        ec_add(run, q);
        ec_add(acc, run);
    }
    // acc = sum (k+1) B_k over the segment; the segment starts at bucket s*L, so add (s*L) * run
    u32 mult = s * L;
    if (mult != 0 && !run.is_zero()) {
        Xyzz<F> m = ec_mul(run, &mult, 1);
        ec_add(acc, m);
    }
    st_struct(seg_out + g, acc);
}

template <class F>
__global__ void __launch_bounds__(128) k_msm_window_sum(const Xyzz<F> *__restrict__ seg_in, u32 nseg,
                                                          Xyzz<F> *__restrict__ win_out) {
    extern __shared__ uint4 smem_raw[];
    Xyzz<F> *sm = reinterpret_cast<Xyzz<F> *>(smem_raw);
    u32 w = blockIdx.x;
    Xyzz<F> acc = Xyzz<F>::zero();
    for (u32 s = threadIdx.x; s < nseg; s += blockDim.x) {
        Xyzz<F> q = ld_struct(seg_in + (size_t)w * nseg + s);
        ec_add(acc, q);
    }
    cta_tree_sum(acc, sm);
    if (threadIdx.x == 0) st_struct(win_out + w, acc);
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <class F>
int msm_run_impl(Ctx *ctx, const void *d_bases_v, const void *d_scalars_v, uint32_t scalar_size, uint64_t n,
                 Xyzz<F> *out_host) {
    typedef Xyzz<F> Pt;
    *out_host = Pt::zero();
    if (n == 0) return B200_OK;
    if (scalar_size == 0 || scalar_size > 32) { ctx->err = "msm: scalar_size must be 1..32 bytes"; return B200_ERR_ARG; }
    if (!d_bases_v || !d_scalars_v) { ctx->err = "msm: null input"; return B200_ERR_ARG; }
    const Affine<F> *d_bases = static_cast<const Affine<F> *>(d_bases_v);
    const uint8_t *d_scalars = static_cast<const uint8_t *>(d_scalars_v);

    MsmGeom g = msm_geometry(n, scalar_size, ctx->force_c);
    if (g.nwin > MSM_MAX_WIN) { ctx->err = "msm: too many windows"; return B200_ERR_ARG; }
    const u32 batch_max = n < MSM_MAX_BATCH ? (u32)n : MSM_MAX_BATCH;
    const size_t hist_len = ((size_t)g.NB + 1 + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    const u32 ntiles = (u32)(hist_len / SCAN_TILE);
    const size_t max_entries = (size_t)batch_max * g.nwin;
    const size_t max_chunks = (max_entries + g.T - 1) / g.T;
    const u32 total_segs = (u32)g.nwin * g.nseg;

    B200_TRY(ctx_reserve(ctx, ctx->w_hist, hist_len * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_cursor, hist_len * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_scan_totals, (size_t)ntiles * 4 + 16));
    B200_TRY(ctx_reserve(ctx, ctx->w_entries, max_entries * 4 + 16));
    B200_TRY(ctx_reserve(ctx, ctx->w_buckets, (size_t)g.NB * sizeof(Pt)));
    B200_TRY(ctx_reserve(ctx, ctx->w_partial, max_chunks * 2 * sizeof(Pt)));
    B200_TRY(ctx_reserve(ctx, ctx->w_hot, ((size_t)g.NB + 1) * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_segs, (size_t)total_segs * sizeof(Pt)));
    B200_TRY(ctx_reserve(ctx, ctx->w_win, (size_t)MSM_MAX_WIN * sizeof(Pt)));
    B200_TRY(ctx_pinned(ctx, (size_t)MSM_MAX_WIN * sizeof(G2Xyzz)));

    u32 *d_hist = (u32 *)ctx->w_hist.p, *d_cursor = (u32 *)ctx->w_cursor.p, *d_totals = (u32 *)ctx->w_scan_totals.p;
    u32 *d_entries = (u32 *)ctx->w_entries.p, *d_hot = (u32 *)ctx->w_hot.p;
    Pt *d_buckets = (Pt *)ctx->w_buckets.p, *d_partial = (Pt *)ctx->w_partial.p;
    Pt *d_segs = (Pt *)ctx->w_segs.p, *d_win = (Pt *)ctx->w_win.p;
    cudaStream_t st = ctx->stream;

    B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_buckets, 0, (size_t)g.NB * sizeof(Pt), st));

    int batch_idx = 0;
    for (uint64_t base = 0; base < n; base += batch_max, batch_idx++) {
        const u32 nb = (u32)((n - base < batch_max) ? (n - base) : batch_max);
        const uint8_t *sc = d_scalars + (size_t)base * scalar_size;
        const Affine<F> *bs = d_bases + base;
        const u32 dgrid = (nb + 255) / 256;

        phase_begin(ctx, PH_MSM_SORT);
        B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_hist, 0, hist_len * 4, st));
        B200_LAUNCH(ctx, k_msm_digits<false>, dgrid, 256, 0, sc, scalar_size, nb, g.c, g.nwin, g.nbk, d_hist, (u32 *)nullptr);
        B200_LAUNCH(ctx, k_scan_tile, ntiles, 1024, 0, d_hist, d_totals);
        B200_LAUNCH(ctx, k_scan_totals, 1, 1024, 0, d_totals, ntiles);
        B200_LAUNCH(ctx, k_scan_add, ntiles, 1024, 0, d_hist, d_totals, d_cursor);
        B200_LAUNCH(ctx, k_msm_digits<true>, dgrid, 256, 0, sc, scalar_size, nb, g.c, g.nwin, g.nbk, d_cursor, d_entries);
        phase_end(ctx);

        phase_begin(ctx, sizeof(F) == 32 ? PH_MSM_ACCUM : PH_MSM_ACCUM_G2);
        const size_t chunks = ((size_t)nb * g.nwin + g.T - 1) / g.T;
        const u32 agrid = (u32)((chunks + 127) / 128);
        auto kacc = k_msm_accumulate<F, (sizeof(F) == 32)>;
        B200_LAUNCH(ctx, kacc, agrid, 128, 0, bs, d_entries, d_hist, g.NB, g.T, d_buckets, d_partial, batch_idx > 0 ? 1 : 0);
        phase_end(ctx);

        phase_begin(ctx, PH_MSM_MERGE);
        B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_hot, 0, 4, st));
        B200_LAUNCH(ctx, k_msm_merge<F>, (g.NB + 127) / 128, 128, 0, d_hist, g.NB, g.T, d_buckets, d_partial, batch_idx > 0 ? 1 : 0, d_hot);
        B200_LAUNCH(ctx, k_msm_merge_hot<F>, 2 * ctx->sm_count, 128, 128 * sizeof(Pt), d_hist, g.T, d_buckets, d_partial, batch_idx > 0 ? 1 : 0, d_hot);
        phase_end(ctx);
    }

    phase_begin(ctx, PH_MSM_REDUCE);
    B200_LAUNCH(ctx, k_msm_reduce_segments<F>, (total_segs + 127) / 128, 128, 0, d_buckets, g.nbk, g.L, g.nseg, total_segs, d_segs);
    B200_LAUNCH(ctx, k_msm_window_sum<F>, g.nwin, 128, 128 * sizeof(Pt), d_segs, g.nseg, d_win);
    phase_end(ctx);

    phase_begin(ctx, PH_MSM_FINAL);
    Pt *h_win = (Pt *)ctx->pinned;
    B200_CUDA_CHECK(ctx, cudaMemcpyAsync(h_win, d_win, (size_t)g.nwin * sizeof(Pt), cudaMemcpyDeviceToHost, st));
    phase_end(ctx);
    B200_CUDA_CHECK(ctx, cudaStreamSynchronize(st));

    // Horner over the windows (multiexp.cpp:137-141) on the host's 4x64 field
    if (sizeof(F) == 32) host_horner_g1(h_win, g.nwin, g.c, out_host);
    else host_horner_g2(h_win, g.nwin, g.c, out_host);
    return B200_OK;
}

}  // namespace b200
