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
//   4. k_msm_plan_*         task list: one task per bucket, buckets larger than CAP entries are cut
//                           into CAP-sized sub-tasks; tasks are counting-sorted by length (longest
//                           first) so the 32 lanes of a warp run equally long loops and the long
//                           tasks start first
//   5. k_msm_accumulate     THE hot kernel, one thread per task: gathers its affine points (64 B /
//                           128 B each, read-only path, next point prefetched while the current
//                           mixed addition runs) into an XYZZ running sum and writes the bucket.
//                           Skewed witnesses (huge |digit| = 1 buckets) become many sub-tasks whose
//                           partial sums are folded by k_msm_merge_hot (one CTA per hot bucket:
//                           strided partial sums + shared-memory tree)
//   6. k_msm_reduce_segments / k_msm_plane_sum / k_msm_window_sum   sum_k k*B_k per bucket set:
//                           running sums over segments of L buckets give (W_seg, S_seg); the weights
//                           of the segment starts come from bit-plane sums of S over the segment index
//   7. host: Horner over the <= 21 planes and <= 65 window sums (a serial chain of ~270 group
//      operations is ~8x faster on one CPU core than on one GPU thread)
//
// Resident tables (zkey point sections are static): k_msm_precompute stores 2^(c*j) * P_i for every
// window j next to the original points, once, at upload time.  All windows then share ONE bucket
// set (digit j of point i is an entry for table point j*n + i), so steps 6-7 shrink by the window
// count and c can grow; only the gather in step 5 touches the larger table.
//
// Zero scalars/digits are skipped like the reference (multiexp.cpp:43); zero bases (0,0) are skipped
// inside the mixed add (multiexp.cpp:40).  Scalars are read as plain 8*scalar_size-bit integers and
// never reduced (multiexp.cpp:118).
#pragma once
#include <type_traits>
#include "ctx.cuh"
#include "memops.cuh"
#include "scan.cuh"

namespace b200 {

struct MsmGeom {
    int c;          // window bits
    int nwin;       // windows (covers nbits + 1 for the signed-digit carry)
    u32 nbk;        // buckets per window = 2^(c-1)
    u32 nwin_b;     // bucket sets: nwin, or 1 when all windows share one set (precomputed tables)
    u32 NB;         // nwin_b * nbk
    u32 CAP;        // max entries per accumulate task
    u32 L;          // buckets per reduce segment
    u32 nseg;       // segments per bucket set
    u32 nplanes;    // ceil(log2(nseg)): bit planes of the segment index
};

static const int MSM_MAX_WIN = 65;
static const u32 MSM_TARGET_TASKS = 1u << 18;
static const u32 MSM_WARM_MAX = 8;           // buckets cut into <= 8 tasks are folded by one thread (a serial chain), more by a
                                             // CTA tree (option "warm_max"; 64 made the 2-GPU shards' narrow top window a 64-add chain)
static const u32 MSM_MAX_CAP = 2048;         // task-length histogram has MSM_MAX_CAP + 1 <= 4096 bins

inline int msm_auto_c(uint64_t n) {
    int lg = 0;
    while ((2ull << lg) <= n) lg++;   // floor(log2 n)
    int c = lg - 5;                   // measured on B200: 2^20 points -> c = 15..16, 2^24 -> 19
    if (c < 4) c = 4;
    if (c > 20) c = 20;
    return c;
}

inline MsmGeom msm_geometry(uint64_t n_batch, uint32_t scalar_size, int c, bool shared_buckets, int target_log2 = 0,
                            bool tail = false, int tail_l = 0) {
    const uint64_t target_tasks = target_log2 > 0 ? (1ull << target_log2) : MSM_TARGET_TASKS;
    MsmGeom g;
    int nbits = (int)scalar_size * 8;
    g.c = c;
    g.nwin = (nbits + c) / c;         // ceil((nbits + 1) / c)
    g.nbk = 1u << (c - 1);
    g.nwin_b = shared_buckets ? 1u : (u32)g.nwin;
    g.NB = g.nwin_b * g.nbk;
    // task cap: with plenty of buckets (>= 2^18) a task is a whole bucket unless it is 2x the average (uniform digits
    // never get there - the bucket sizes are Poisson around the average - while the huge |digit| = 1 buckets of a
    // circom-style witness are cut into pieces short enough not to be the kernel's long pole);
    // with few, large buckets they are cut so that about MSM_TARGET_TASKS tasks exist (load balance)
    uint64_t ent = n_batch * (uint64_t)g.nwin;
    uint64_t avg = ent / g.NB + 1;
    u32 cap = 32;
    if ((ent < g.NB ? ent : g.NB) >= (1u << 18)) { while (cap < 2 * avg && cap < MSM_MAX_CAP) cap <<= 1; }
    else { while ((uint64_t)cap * 2 * target_tasks <= ent && cap < MSM_MAX_CAP) cap <<= 1; }
    g.CAP = cap;
    // reduce segments: longer for big bucket sets (amortises the per-segment multiplier, keeps the side-stream
    // kernel's footprint to a few dozen CTAs)
    // (k_msm_reduce_segments walks a segment with two lanes, L + 1 dependent additions: twice the round-1 lengths
    // for the same chain, half the segments for the plane sums - 2^20 proof 1 % faster with 128 than with 64, 256 is
    // 7 % slower, r02 runs 13 / 14)
    g.L = g.nbk >= (1u << 18) ? 128 : g.nbk >= (1u << 17) ? 32 : g.nbk >= 16 ? 16 : g.nbk;
    // one shared bucket set (resident tables) below 2^19 buckets = a shard of a multi-GPU zkey: the accumulations
    // are short, the reduction chains are what the proof waits for (round 1, 2 / 4 shards: 16 beats 64 by 7 %;
    // round 2 with the fused accumulation launch, rank 0 of 8 / of 4: 8 beats 16 by 8 % / 2 %, 4 is worse again)
    // round 2 with the two-lane reduction (r02 run 15, emulated rank 0 of 2 / 4 and rank 7 of 8): 16 beats 8 by
    // 0.5 % / 1.6 % / 0.7 %, 4 is 15 % slower
    if (shared_buckets && g.nbk < (1u << 19) && g.L > 16) g.L = 16;
    const u32 tl = (tail_l > 0 && (tail_l & (tail_l - 1)) == 0) ? (u32)tail_l : 32u;   // option "reduce_l_tail" (16 and 64: 1 % slower)
    if (tail && g.L > tl) g.L = tl;   // nothing left to overlap with: shortest chains, the whole GPU is free
    g.nseg = g.nbk / g.L;
    g.nplanes = 0;
    while ((1u << g.nplanes) < g.nseg) g.nplanes++;
    return g;
}

// ------------------------------------------------------------------------------------------------
// 1 + 3: signed digits, histogram / scatter
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const uint8_t *__restrict__ scalars, u32 scalar_size, u32 n,
                                                      int c, int nwin, u32 nbk, u32 shared_buckets, u32 table_stride,
                                                      u32 *__restrict__ counters, u32 *__restrict__ entries) {
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
            u32 b = (shared_buckets ? 0u : (u32)w * nbk) + mag - 1;
            if (!SCATTER) {
                atomicAdd(&counters[b], 1u);
            } else {
                u32 pos = atomicAdd(&counters[b], 1u);
                // table point of (window w, point i): w * table_stride + i  (table_stride = 0: plain bases)
                entries[pos] = ((u32)w * table_stride + i) | (neg << 31);
            }
        }
    }
}

// 2: exclusive scan of the u32 counters: scan.cuh (k_scan_tile / k_scan_totals / k_scan_add)

// ------------------------------------------------------------------------------------------------
// 4: task plan - buckets (or CAP-sized pieces of big buckets) sorted by length, longest first
// ------------------------------------------------------------------------------------------------
// plan[0] = partial slots allocated, plan[1] = hot buckets (> MSM_WARM_MAX tasks), plan[2] = tasks,
// plan[3] = warm buckets (2..MSM_WARM_MAX tasks)
// Both plan kernels aggregate per CTA in shared memory first: with uniform digits every bucket holds about the same
// number of entries, so all 2^19 buckets of a 2^20-point MSM hit the same ~40 length bins - one global atomic per
// bucket on those few addresses cost 0.1 ms per kernel (r02 launch list), i.e. 0.2 ms of the 0.5 ms digit sort that
// sits at the head of every proof.
static const u32 MSM_PLAN_THREADS = 1024;
static __global__ void __launch_bounds__(MSM_PLAN_THREADS) k_msm_plan_count(const u32 *__restrict__ off, u32 NB, u32 CAP,
                                                          u32 *__restrict__ lenhist /* [4096], index CAP_MAX - len */,
                                                          u32 *__restrict__ plan, u32 *__restrict__ hot_base,
                                                          u32 *__restrict__ hot_list, u32 *__restrict__ warm_list, u32 warm_max) {
    __shared__ u32 sh_hist[MSM_MAX_CAP];          // bins MSM_MAX_CAP - len, len = 1 .. CAP <= MSM_MAX_CAP
    for (u32 i = threadIdx.x; i < MSM_MAX_CAP; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u32 cnt = b < NB ? off[b + 1] - off[b] : 0;
    if (cnt) {
        u32 nfull = cnt / CAP, rem = cnt - nfull * CAP;
        if (nfull) atomicAdd(&sh_hist[MSM_MAX_CAP - CAP], nfull);
        if (rem) atomicAdd(&sh_hist[MSM_MAX_CAP - rem], 1u);
        u32 ntask = nfull + (rem ? 1u : 0u);
        if (ntask > 1) {                          // rare: a bucket of more than CAP entries
            hot_base[b] = atomicAdd(&plan[0], ntask);
            if (ntask > warm_max) hot_list[atomicAdd(&plan[1], 1u)] = b;
            else warm_list[atomicAdd(&plan[3], 1u)] = b;
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < MSM_MAX_CAP; i += blockDim.x) {
        u32 v = sh_hist[i];
        if (v) atomicAdd(&lenhist[i], v);
    }
}

static __global__ void __launch_bounds__(MSM_PLAN_THREADS) k_msm_plan_place(const u32 *__restrict__ off, u32 NB, u32 CAP,
                                                          u32 *__restrict__ cursor /* scanned lenhist */,
                                                          uint2 *__restrict__ tasks) {
    __shared__ u32 sh_cnt[MSM_MAX_CAP], sh_base[MSM_MAX_CAP];
    for (u32 i = threadIdx.x; i < MSM_MAX_CAP; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u32 cnt = b < NB ? off[b + 1] - off[b] : 0;
    u32 nfull = cnt / CAP, rem = cnt - nfull * CAP;
    u32 rank_full = 0, rank_rem = 0;              // this bucket's places inside the CTA's share of each length bin
    if (nfull) rank_full = atomicAdd(&sh_cnt[MSM_MAX_CAP - CAP], nfull);
    if (rem) rank_rem = atomicAdd(&sh_cnt[MSM_MAX_CAP - rem], 1u);
    __syncthreads();
    for (u32 i = threadIdx.x; i < MSM_MAX_CAP; i += blockDim.x) {
        u32 v = sh_cnt[i];
        if (v) sh_base[i] = atomicAdd(&cursor[i], v);
    }
    __syncthreads();
    if (nfull) {
        u32 pos = sh_base[MSM_MAX_CAP - CAP] + rank_full;
        for (u32 j = 0; j < nfull; j++) tasks[pos + j] = make_uint2(b, j);
    }
    if (rem) tasks[sh_base[MSM_MAX_CAP - rem] + rank_rem] = make_uint2(b, nfull);
}

// ------------------------------------------------------------------------------------------------
// 5: bucket accumulation, one thread per task
// ------------------------------------------------------------------------------------------------
// running sum in shared memory: field k of thread t lives at base[k * blockDim.x + t] (16-byte units), so a
// warp's 128-bit accesses are conflict-free; only one mixed addition's temporaries stay in registers
template <class F>
struct SmemAcc {
    uint4 *base;   // already offset by threadIdx.x
    static const int FV = sizeof(F) / 16;
    DEVFN F ld(int field) const {
        F r;
        uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
        for (int i = 0; i < FV; i++) d[i] = base[(field * FV + i) * blockDim.x];
        return r;
    }
    DEVFN void st(int field, const F &v) {
        const uint4 *s = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
        for (int i = 0; i < FV; i++) base[(field * FV + i) * blockDim.x] = s[i];
    }
    DEVFN F ld_x() const { return ld(0); }
    DEVFN F ld_y() const { return ld(1); }
    DEVFN F ld_zz() const { return ld(2); }
    DEVFN F ld_zzz() const { return ld(3); }
    DEVFN void st_x(const F &v) { st(0, v); }
    DEVFN void st_y(const F &v) { st(1, v); }
    DEVFN void st_zz(const F &v) { st(2, v); }
    DEVFN void st_zzz(const F &v) { st(3, v); }
    DEVFN void st_all(const Xyzz<F> &v) { st(0, v.x); st(1, v.y); st(2, v.zz); st(3, v.zzz); }
    DEVFN Xyzz<F> get() const { Xyzz<F> r; r.x = ld(0); r.y = ld(1); r.zz = ld(2); r.zzz = ld(3); return r; }
};

// ---- experimental variant (option "acc_smem" = 2, G2 only, NOT the default; to be measured) -----------------------
// The G2 loop keeps 250 registers alive (running sum 64, current + prefetched point 64, temporaries) and therefore
// runs 8 warps per SM; ncu shows its warps mostly in fixed-latency `wait` with the multiplier pipe 67-73 % busy
// (the G1 loop: 16 warps, 90 %).  Here the running sum AND the gathered points live in shared memory: the next
// point is fetched with cp.async straight into a double-buffered slot (no registers, no stall), coordinates are
// loaded where they are used.  512 bytes of shared memory per thread (64 KB per CTA, 3 CTAs per SM).
DEVFN void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
DEVFN void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DEVFN void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class F>
struct SmemPoint {            // one affine point of this thread in shared memory: unit k at base[k * blockDim.x]
    typedef F Field;
    uint4 *base;              // already offset by threadIdx.x
    bool neg;                 // the entry's sign bit: use -y
    static const int FV = sizeof(F) / 16;
    DEVFN F ld(int field) const {
        F r;
        uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
        for (int i = 0; i < FV; i++) d[i] = base[(field * FV + i) * blockDim.x];
        return r;
    }
    DEVFN F ld_x() const { return ld(0); }
    DEVFN F ld_y() const { F y = ld(1); return neg ? fneg(y) : y; }
    DEVFN bool is_zero() const {
        uint4 t = base[0];
#pragma unroll
        for (int i = 1; i < 2 * FV; i++) { uint4 u = base[i * blockDim.x]; t.x |= u.x; t.y |= u.y; t.z |= u.z; t.w |= u.w; }
        return (t.x | t.y | t.z | t.w) == 0;
    }
    DEVFN Affine<F> get() const { Affine<F> a; a.x = ld_x(); a.y = ld_y(); return a; }
    // once both coordinates have been consumed the slot is free: scratch element 0 / 1 of the running addition
    DEVFN void st_scratch(int field, const F &v) {
        const uint4 *s = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
        for (int i = 0; i < FV; i++) base[(field * FV + i) * blockDim.x] = s[i];
    }
    DEVFN F ld_scratch(int field) const { return ld(field); }
};

template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate_staged(const Affine<F> *__restrict__ bases,
                                                                       const u32 *__restrict__ entries,
                                                                       const u32 *__restrict__ off,
                                                                       const uint2 *__restrict__ tasks,
                                                                       const u32 *__restrict__ ntasks_ptr, u32 CAP,
                                                                       const u32 *__restrict__ hot_base,
                                                                       Xyzz<F> *__restrict__ buckets,
                                                                       Xyzz<F> *__restrict__ partial, int add_existing) {
    extern __shared__ uint4 smem_raw[];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *ntasks_ptr) return;
    const uint2 task = tasks[t];
    const u32 b = task.x;
    const u32 o0 = off[b], cnt = off[b + 1] - o0;
    const u32 start = o0 + task.y * CAP;
    const u32 len = (cnt - task.y * CAP < CAP) ? cnt - task.y * CAP : CAP;
    constexpr int FV = sizeof(F) / 16;        // 16-byte units per field element
    constexpr int PU = 2 * FV;                // units per affine point

    SmemAcc<F> acc;
    acc.base = smem_raw + threadIdx.x;                                   // units [0, 4 FV) x blockDim
    uint4 *pt = smem_raw + (size_t)4 * FV * blockDim.x + threadIdx.x;    // two point buffers of PU units each
    acc.st_all(Xyzz<F>::zero());

    u32 e_next = entries[start];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(bases + (e_next & 0x7fffffffu));
#pragma unroll
        for (int k = 0; k < PU; k++) cp_async16(pt + (size_t)k * blockDim.x, src + k);
        cp_async_commit();
    }
    int buf = 0;
    for (u32 k = 0; k < len; k++) {
        const u32 e = e_next;
        cp_async_wait_all();                                             // the point of entry k is in buffer `buf`
        if (k + 1 < len) {
            e_next = entries[start + k + 1];
            const uint4 *src = reinterpret_cast<const uint4 *>(bases + (e_next & 0x7fffffffu));
            uint4 *dst = pt + (size_t)(buf ^ 1) * PU * blockDim.x;
#pragma unroll
            for (int u = 0; u < PU; u++) cp_async16(dst + (size_t)u * blockDim.x, src + u);
            cp_async_commit();
        }
        SmemPoint<F> q;
        q.base = pt + (size_t)buf * PU * blockDim.x;
        q.neg = (e >> 31) != 0;
        ec_madd_acc_pt(acc, q);
        buf ^= 1;
    }
    Xyzz<F> r = acc.get();
    if (cnt <= CAP) {
        if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(r, old); }
        st_struct(buckets + b, r);
    } else {
        st_struct(partial + hot_base[b] + task.y, r);
    }
}

#ifndef B200_G2_PREFETCH_L2
#define B200_G2_PREFETCH_L2 1      // 1: prefetch.global.L2, 2: prefetch.global.L1 (only the 168-register G2 kernel uses it)
#endif
#ifndef B200_G2_PREFETCH_L2_MINB
#define B200_G2_PREFETCH_L2_MINB 3   // from how many CTAs per SM on; 2 = also the default 255-register kernel (A/B builds)
#endif
DEVFN void prefetch_l2(const void *p) {
#if B200_G2_PREFETCH_L2 == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// CTA-wide maximum of one u32 per thread (NT threads, every thread of the CTA calls it)
template <int NT>
DEVFN u32 cta_max_u32(u32 v) {
    __shared__ u32 s_wmax[NT / 32];
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = v;
    __syncthreads();
    u32 m = s_wmax[0];
#pragma unroll
    for (int i = 1; i < NT / 32; i++) m = s_wmax[i] > m ? s_wmax[i] : m;
    return m;
}

// LOCK (option "lockstep"): the warps of a CTA walk the loop together - one barrier per mixed addition, trip count =
// the longest task of the CTA (tasks are sorted by length, so the CTA's tasks are equally long give or take one).  The
// fully inlined G2 addition is ~100 KB of SASS, three times the SM's 32 KB instruction cache (B300_MICROARCH.md), and
// ncu charges the loop one `no_instruction` stall cycle per issued instruction: warps that have drifted apart each
// stream the whole body from L2; warps at the same place share every fetched line.  NT = threads per CTA.
template <class F, bool ACC_SMEM, int MINB, bool LOCK = false, int NT = 128>
__global__ void __launch_bounds__(NT, MINB) k_msm_accumulate(const Affine<F> *__restrict__ bases,
                                                                const u32 *__restrict__ entries,
                                                                const u32 *__restrict__ off,
                                                                const uint2 *__restrict__ tasks,
                                                                const u32 *__restrict__ ntasks_ptr, u32 CAP,
                                                                const u32 *__restrict__ hot_base,
                                                                Xyzz<F> *__restrict__ buckets,
                                                                Xyzz<F> *__restrict__ partial, int add_existing) {
    extern __shared__ uint4 smem_raw[];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < *ntasks_ptr;
    if (!LOCK && !live) return;
    const uint2 task = live ? tasks[t] : make_uint2(0, 0);
    const u32 b = task.x;
    const u32 o0 = off[b], cnt = off[b + 1] - o0;
    const u32 start = o0 + task.y * CAP;
    const u32 len = !live ? 0 : (cnt - task.y * CAP < CAP) ? cnt - task.y * CAP : CAP;
    u32 trips = len;
    if constexpr (LOCK) {
        trips = cta_max_u32<NT>(len);
        if (trips == 0) return;             // a CTA past the last task
    }

    typename std::conditional<ACC_SMEM, SmemAcc<F>, RegAcc<F>>::type acc;
    if constexpr (ACC_SMEM) acc.base = smem_raw + threadIdx.x;
    acc.st_all(Xyzz<F>::zero());

    // next point: held in registers while the current addition runs (default), or - PFL2, the G2 kernel compiled for
    // three CTAs per SM - only pulled into L2 (prefetch.global.L2: one 128-byte line = one G2 point) and loaded where it
    // is used: 32 registers fewer across the whole addition
    constexpr bool PFL2 = B200_G2_PREFETCH_L2 && sizeof(F) != 32 && MINB >= B200_G2_PREFETCH_L2_MINB;
    u32 e_next = len ? entries[start] : 0;
    Affine<F> p_next;
    if constexpr (PFL2) prefetch_l2(bases + (e_next & 0x7fffffffu));
    else p_next = ldg_struct(bases + (e_next & 0x7fffffffu));
    for (u32 k = 0; k < trips; k++) {
        if constexpr (LOCK) __syncthreads();
        if (!LOCK || k < len) {
            u32 e = e_next;
            Affine<F> p;
            if constexpr (PFL2) p = ldg_struct(bases + (e & 0x7fffffffu));
            else p = p_next;
            if (k + 1 < len) {
                e_next = entries[start + k + 1];
                if constexpr (PFL2) prefetch_l2(bases + (e_next & 0x7fffffffu));
                else p_next = ldg_struct(bases + (e_next & 0x7fffffffu));
            }
            if (e >> 31) p.y = fneg(p.y);
            ec_madd_acc(acc, p);
        }
    }
    if (!live) return;
    Xyzz<F> r = acc.get();
    if (cnt <= CAP) {
        if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(r, old); }
        st_struct(buckets + b, r);
    } else {
        st_struct(partial + hot_base[b] + task.y, r);
    }
}

// Several MSMs in ONE accumulation launch (blockIdx.y = MSM): the witness MSMs A, B1, C read the same sorted entry
// list with different tables, H brings its own list.  One grid of all their tasks packs the SMs better than one grid
// per MSM - a shard of a multi-GPU zkey has only 2^16 tasks per MSM, less than one wave of CTAs - and there is one
// drain at the end instead of one per MSM.
template <class F>
struct MsmAccSet {
    const Affine<F> *bases;
    const u32 *entries, *off, *ntasks_ptr, *hot_base;
    const uint2 *tasks;
    Xyzz<F> *buckets, *partial;
    u32 CAP;
};
static const int MSM_MAX_FUSE = 4;
template <class F>
struct MsmAccSets { MsmAccSet<F> s[MSM_MAX_FUSE]; };

template <class F, int MINB, bool LOCK = false>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate_sets(const __grid_constant__ MsmAccSets<F> sets) {
    const MsmAccSet<F> &a = sets.s[blockIdx.y];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < *a.ntasks_ptr;
    if (!LOCK && !live) return;
    const uint2 task = live ? a.tasks[t] : make_uint2(0, 0);
    const u32 b = task.x, CAP = a.CAP;
    const u32 o0 = a.off[b], cnt = a.off[b + 1] - o0;
    const u32 start = o0 + task.y * CAP;
    const u32 len = !live ? 0 : (cnt - task.y * CAP < CAP) ? cnt - task.y * CAP : CAP;
    const u32 *__restrict__ entries = a.entries;
    const Affine<F> *__restrict__ bases = a.bases;
    u32 trips = len;
    if constexpr (LOCK) {                   // see k_msm_accumulate
        trips = cta_max_u32<128>(len);
        if (trips == 0) return;
    }

    RegAcc<F> acc;
    acc.st_all(Xyzz<F>::zero());
    u32 e_next = len ? entries[start] : 0;
    Affine<F> p_next = ldg_struct(bases + (e_next & 0x7fffffffu));
    for (u32 k = 0; k < trips; k++) {
        if constexpr (LOCK) __syncthreads();
        if (!LOCK || k < len) {
            u32 e = e_next;
            Affine<F> p = p_next;
            if (k + 1 < len) {
                e_next = entries[start + k + 1];
                p_next = ldg_struct(bases + (e_next & 0x7fffffffu));
            }
            if (e >> 31) p.y = fneg(p.y);
            ec_madd_acc(acc, p);
        }
    }
    if (!live) return;
    Xyzz<F> r = acc.get();
    if (cnt <= CAP) st_struct(a.buckets + b, r);
    else st_struct(a.partial + a.hot_base[b] + task.y, r);
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
__global__ void __launch_bounds__(128) k_msm_merge_warm(const u32 *__restrict__ off, u32 CAP,
                                                          Xyzz<F> *__restrict__ buckets,
                                                          const Xyzz<F> *__restrict__ partial, int add_existing,
                                                          const u32 *__restrict__ plan, const u32 *__restrict__ hot_base,
                                                          const u32 *__restrict__ warm_list) {
    const u32 nwarm = plan[3];
    for (u32 h = blockIdx.x * blockDim.x + threadIdx.x; h < nwarm; h += gridDim.x * blockDim.x) {
        const u32 b = warm_list[h];
        const u32 cnt = off[b + 1] - off[b];
        const u32 ntask = (cnt + CAP - 1) / CAP;
        const Xyzz<F> *src = partial + hot_base[b];
        Xyzz<F> acc = ld_struct(src);
        for (u32 j = 1; j < ntask; j++) {
            Xyzz<F> q = ld_struct(src + j);
            ec_add(acc, q);
        }
        if (add_existing) { Xyzz<F> old = ld_struct(buckets + b); ec_add(acc, old); }
        st_struct(buckets + b, acc);
    }
}

template <class F>
__global__ void __launch_bounds__(128) k_msm_merge_hot(const u32 *__restrict__ off, u32 CAP,
                                                         Xyzz<F> *__restrict__ buckets,
                                                         const Xyzz<F> *__restrict__ partial, int add_existing,
                                                         const u32 *__restrict__ plan, const u32 *__restrict__ hot_base,
                                                         const u32 *__restrict__ hot_list) {
    extern __shared__ uint4 smem_raw[];
    Xyzz<F> *sm = reinterpret_cast<Xyzz<F> *>(smem_raw);
    const u32 nhot = plan[1];
    for (u32 h = blockIdx.x; h < nhot; h += gridDim.x) {
        const u32 b = hot_list[h];
        const u32 cnt = off[b + 1] - off[b];
        const u32 ntask = (cnt + CAP - 1) / CAP;
        const Xyzz<F> *src = partial + hot_base[b];
        Xyzz<F> acc = Xyzz<F>::zero();
        for (u32 j = threadIdx.x; j < ntask; j += blockDim.x) {
            Xyzz<F> q = ld_struct(src + j);
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
// Launched with 32-thread CTAs and <= 168 registers: ~5 K registers per CTA slip in next to the three resident
// accumulation CTAs of an SM (3 x 18 K of 64 K) instead of evicting two of them, which is what a 128-thread /
// 255-register CTA did (ncu launch list, profiles/r01_*: the reductions were 28% of the serialised time).
// TWO lanes per segment.  The recursion  run += B_k; acc += run  is two chains of L dependent full additions of
// which step k of the second needs only step k of the first: lane 2i walks the `run` chain and hands every new value
// to lane 2i + 1 (warp shuffle), which adds it to `acc` one step behind - L + 1 dependent additions instead of 2 L.
// These kernels are pure latency (10 us per dependent full addition for a lone warp; the per-MSM side streams and
// the tail of a proof wait for exactly this chain), so halving the chain halves their duration; the number of
// additions executed is the same.
template <class F>
DEVFN Xyzz<F> shfl_point(const Xyzz<F> &v, int src_lane) {
    Xyzz<F> r;
    const u32 *s = reinterpret_cast<const u32 *>(&v);
    u32 *d = reinterpret_cast<u32 *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(Xyzz<F>) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], src_lane);
    return r;
}

template <class F>
__global__ void __launch_bounds__(128, 3) k_msm_reduce_segments(const Xyzz<F> *__restrict__ buckets, u32 nbk, u32 L,
                                                               u32 nseg, u32 total_segs,
                                                               Xyzz<F> *__restrict__ seg_out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 g = t >> 1;                         // segment; both lanes of a pair are in the same warp (t even/odd)
    const bool second = (t & 1u) != 0;            // false: the `run` chain, true: the `acc` chain
    const bool live = g < total_segs;             // dead pairs still take part in the shuffles
    const u32 w = live ? g / nseg : 0, sidx = live ? g % nseg : 0;
    const Xyzz<F> *B = buckets + (size_t)w * nbk + (size_t)sidx * L;
    const int lane = (int)(threadIdx.x & 31u);
    Xyzz<F> mine = Xyzz<F>::zero();               // run (first lane) or acc (second lane)
    Xyzz<F> handed = Xyzz<F>::zero();             // second lane: `run` as it was after the previous step
    for (int k = (int)L - 1; k >= -1; k--) {
        Xyzz<F> q;
        if (second) q = handed;
        else if (live && k >= 0) q = ld_struct(B + k);
        else q = Xyzz<F>::zero();
        ec_add(mine, q);                          // first lane: run += B_k;  second lane: acc += run (one step behind)
        handed = shfl_point(mine, lane & ~1);
    }
    // acc = sum (k+1) B_k over the segment (local weights), run = plain sum.  The segment starts at bucket s*L:
    // its true contribution is acc + (s*L) * run; the second term is assembled from bit-plane sums of `run`
    // over the segment index (k_msm_plane_sum) instead of a per-thread scalar multiplication.
    if (live) st_struct(seg_out + 2 * (size_t)g + (second ? 0 : 1), mine);
}

// Bit-plane sums over the segments of one bucket set: CTA (chunk, plane, set).
//   plane t < nplanes : sum of S_seg over the chunk's segments whose index has bit t set
//   plane == nplanes  : sum of W_seg over the chunk's segments
// out[(set * (nplanes + 1) + plane) * nchunk + chunk].  sum_seg seg * S_seg = sum_t 2^t * plane_t.
template <class F>
__global__ void __launch_bounds__(128) k_msm_plane_sum(const Xyzz<F> *__restrict__ segs, u32 nseg, u32 nchunk, u32 nplanes,
                                                         Xyzz<F> *__restrict__ out) {
    extern __shared__ uint4 smem_raw[];
    Xyzz<F> *sm = reinterpret_cast<Xyzz<F> *>(smem_raw);
    const u32 chunk = blockIdx.x, plane = blockIdx.y, w = blockIdx.z;
    const u32 per = (nseg + nchunk - 1) / nchunk;
    const u32 lo = chunk * per, hi = (lo + per < nseg) ? lo + per : nseg;
    const Xyzz<F> *base = segs + 2 * (size_t)w * nseg;
    Xyzz<F> acc = Xyzz<F>::zero();
    for (u32 s = lo + threadIdx.x; s < hi; s += blockDim.x) {
        if (plane == nplanes) {
            Xyzz<F> q = ld_struct(base + 2 * (size_t)s);
            ec_add(acc, q);
        } else if ((s >> plane) & 1u) {
            Xyzz<F> q = ld_struct(base + 2 * (size_t)s + 1);
            ec_add(acc, q);
        }
    }
    cta_tree_sum(acc, sm);
    if (threadIdx.x == 0) st_struct(out + ((size_t)w * (nplanes + 1) + plane) * nchunk + chunk, acc);
}

// plain sums: CTA (chunk, set) adds its slice of the set's nitems points; out[set * nchunk + chunk]
template <class F>
__global__ void __launch_bounds__(128) k_msm_window_sum(const Xyzz<F> *__restrict__ in, u32 nitems, u32 nchunk,
                                                          Xyzz<F> *__restrict__ out) {
    extern __shared__ uint4 smem_raw[];
    Xyzz<F> *sm = reinterpret_cast<Xyzz<F> *>(smem_raw);
    const u32 chunk = blockIdx.x, w = blockIdx.y;
    const u32 per = (nitems + nchunk - 1) / nchunk;
    const u32 lo = chunk * per, hi = (lo + per < nitems) ? lo + per : nitems;
    Xyzz<F> acc = Xyzz<F>::zero();
    for (u32 s = lo + threadIdx.x; s < hi; s += blockDim.x) {
        Xyzz<F> q = ld_struct(in + (size_t)w * nitems + s);
        ec_add(acc, q);
    }
    cta_tree_sum(acc, sm);
    if (threadIdx.x == 0) st_struct(out + (size_t)w * nchunk + chunk, acc);
}

// ------------------------------------------------------------------------------------------------
// resident tables: tbl[j * n + i] = 2^(c*j) * P_i (affine), j < nwin; row 0 is a copy of the points
// ------------------------------------------------------------------------------------------------
static const int MSM_PRE_MAX_WIN = 24;   // tables are only built for c >= 11

template <class F>
__global__ void __launch_bounds__(128) k_msm_precompute(const Affine<F> *__restrict__ pts, u32 n, int c, int nwin,
                                                          Affine<F> *__restrict__ tbl) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = ldg_struct(pts + i);
    st_struct(tbl + i, p);
    Xyzz<F> q[MSM_PRE_MAX_WIN];   // local memory: 2^(c*j) * P, j = 1..nwin-1, still in XYZZ
    F pref[MSM_PRE_MAX_WIN];      // running products of zz*zzz for one shared inversion (Montgomery's trick)
    Xyzz<F> cur = Xyzz<F>::from_affine(p);
    F run = F::one();
    for (int j = 1; j < nwin; j++) {
        for (int k = 0; k < c; k++) cur = ec_dbl(cur);
        q[j] = cur;
        if (!cur.is_zero()) run = fmul(run, fmul(cur.zz, cur.zzz));
        pref[j] = run;
    }
    F inv = finv(run);
    for (int j = nwin - 1; j >= 1; j--) {
        Affine<F> a;
        if (q[j].is_zero()) { a.x = F::zero(); a.y = F::zero(); }
        else {
            F before = (j > 1) ? pref[j - 1] : F::one();
            F iz = fmul(inv, before);                       // 1 / (zz_j * zzz_j)
            inv = fmul(inv, fmul(q[j].zz, q[j].zzz));
            a.x = fmul(q[j].x, fmul(iz, q[j].zzz));
            a.y = fmul(q[j].y, fmul(iz, q[j].zz));
        }
        st_struct(tbl + (size_t)j * n + i, a);
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <class F>
struct MsmTable {          // resident precomputed table (nullptr tbl: plain bases, no precomputation)
    const Affine<F> *tbl = nullptr;
    u32 n = 0;             // points per row
    int c = 0, nwin = 0;   // geometry the table was built for (scalar_size = 32)
};

template <class F>
int msm_precompute_table(Ctx *ctx, const Affine<F> *d_pts, u32 n, int c, Affine<F> *d_tbl) {
    MsmGeom g = msm_geometry(n, 32, c, true);
    if (n == 0) return B200_OK;
    if (g.nwin > MSM_PRE_MAX_WIN) { ctx->err = "msm: window too small for a precomputed table"; return B200_ERR_ARG; }
    B200_LAUNCH(ctx, k_msm_precompute<F>, (n + 127) / 128, 128, 0, d_pts, n, c, g.nwin, d_tbl);
    return B200_OK;
}

// What is left to do for one MSM once its accumulation kernel has been launched (or queued for a fused launch)
template <class F>
struct MsmPending {
    MsmGeom g;
    int slot = 0, ws = 0;
    u32 tree_threads = 0, agrid = 0, npoints = 0;
    int add_existing = 0;
    bool last_batch = true;
    const Affine<F> *bases = nullptr;
    u32 *d_hist = nullptr, *d_plan = nullptr, *d_hot_base = nullptr, *d_hot_list = nullptr, *d_warm_list = nullptr, *d_entries = nullptr;
    uint2 *d_tasks = nullptr;
    Xyzz<F> *d_buckets = nullptr, *d_partial = nullptr, *d_segs = nullptr, *d_win = nullptr, *h_win = nullptr;
    cudaStream_t side = nullptr;
};

// MSMs whose accumulations go into one launch (msm_fuse_flush)
template <class F>
struct MsmFuse {
    int n = 0;
    MsmPending<F> job[MSM_MAX_FUSE];
};

// folding of split buckets, then - for the last batch - bucket reduction, window sums and the read-back, on the MSM's
// side stream: overlaps whatever the main stream does next
template <class F>
int msm_post_impl(Ctx *ctx, const MsmPending<F> &p) {
    typedef Xyzz<F> Pt;
    const MsmGeom &g = p.g;
    const int slot = p.slot, ws = p.ws;
    cudaStream_t st = ctx->stream, side = p.side;
    B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_ws_acc[ws], st));
    ctx->ws_acc_pending[ws] = true;
    cudaStream_t ms = p.last_batch ? side : st;
    if (p.last_batch) {
        B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_acc[slot], st));
        B200_CUDA_CHECK(ctx, cudaStreamWaitEvent(side, ctx->ev_acc[slot], 0));
    }
    phase_begin(ctx, PH_MSM_MERGE, ms);
    // side-stream kernels keep a CTA's register footprint at or below one G1 accumulation CTA (128 x 128
    // registers), otherwise they only get onto an SM when two of those retire together: 32-thread CTAs for the
    // per-thread folds, 64-thread trees for G2 (168 registers per thread)
    B200_LAUNCH_ON(ctx, ms, k_msm_merge_warm<F>, 16 * ctx->sm_count, 32, 0, p.d_hist, g.CAP, p.d_buckets, p.d_partial, p.add_existing, p.d_plan, p.d_hot_base, p.d_warm_list);
    B200_LAUNCH_ON(ctx, ms, k_msm_merge_hot<F>, 2 * ctx->sm_count, p.tree_threads, p.tree_threads * sizeof(Pt), p.d_hist, g.CAP, p.d_buckets, p.d_partial, p.add_existing, p.d_plan, p.d_hot_base, p.d_hot_list);
    phase_end(ctx, ms);
    if (!p.last_batch) return B200_OK;
    B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_merge[slot], side));
    ctx->sort_readers[ws] |= 1u << slot;

    // bucket reduction + window sums + D2H on the side stream: overlaps the next MSM's sort / accumulation
    const u32 total_segs = g.nwin_b * g.nseg, npl1 = g.nplanes + 1;
    phase_begin(ctx, PH_MSM_REDUCE, side);
    B200_LAUNCH_ON(ctx, side, k_msm_reduce_segments<F>, (2 * total_segs + 31) / 32, 32, 0, p.d_buckets, g.nbk, g.L, g.nseg, total_segs, p.d_segs);
    {
        // plane sums in two passes: nchunk CTAs per (set, plane), then one CTA per (set, plane) over the chunk sums
        // chunk size: 1024 segments per CTA = 8 (G1) / 16 (G2) per thread, added one after the other before the CTA
        // tree.  Shorter chains (option "plane_items" = segments per thread; 2 and 4 measured on B200, r02 run 14) mean
        // four times as many high-priority CTAs squeezing in beside the accumulation: 2 % slower on one GPU, no gain
        // on an emulated shard
        const u32 per_cta = ctx->opt_plane_items > 0 ? (u32)ctx->opt_plane_items * p.tree_threads : 1024u;
        u32 nchunk = (g.nseg + per_cta - 1) / per_cta;
        if (nchunk > 128) nchunk = 128;
        if (nchunk < 1) nchunk = 1;
        Pt *d_chunk = p.d_segs + 2 * (size_t)total_segs;
        B200_LAUNCH_ON(ctx, side, k_msm_plane_sum<F>, dim3(nchunk, npl1, g.nwin_b), p.tree_threads, p.tree_threads * sizeof(Pt), p.d_segs, g.nseg, nchunk, g.nplanes, d_chunk);
        B200_LAUNCH_ON(ctx, side, k_msm_window_sum<F>, dim3(1, npl1 * g.nwin_b), p.tree_threads, p.tree_threads * sizeof(Pt), d_chunk, nchunk, 1u, p.d_win);
    }
    phase_end(ctx, side);
    B200_CUDA_CHECK(ctx, cudaMemcpyAsync(p.h_win, p.d_win, (size_t)g.nwin_b * npl1 * sizeof(Pt), cudaMemcpyDeviceToHost, side));
    B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_done[slot], side));
    ctx->slot_busy[slot] = true;
    Ctx::SlotInfo &si = ctx->slot_info[slot];
    si.nwin_b = (int)g.nwin_b; si.nwin = g.nwin; si.c = g.c; si.nplanes = (int)g.nplanes; si.L = (int)g.L; si.used = true;
    return B200_OK;
}

// G1 accumulation in lockstep (k_msm_accumulate, LOCK)?  Option "lockstep_g1": 0 off, 1 on, -1 (default) by size.
// Measured on B200 (profiles/r02_g2_experiments.md): the kernel alone gets 1.6 % slower (barrier waits), a whole 2^20
// proof 1.5 % faster - the latency-bound bucket reductions running beside it on the side streams finish sooner; no
// difference at 2^22, 0.7 % slower at 2^24, where the reductions are a small part of the work.  Shards (<= 2^19
// points) and single MSM calls (no reductions beside them) were not measured with it and stay as they were.
static inline bool msm_lockstep_g1(const Ctx *ctx, u32 npoints) {
    if (ctx->opt_lockstep_g1 >= 0) return ctx->opt_lockstep_g1 > 0;
    return npoints > (1u << 19) && npoints <= (1u << 21);
}

// one accumulation launch for all queued MSMs, then each one's post-processing on its own side stream
template <class F>
int msm_fuse_flush(Ctx *ctx, MsmFuse<F> *fuse) {
    if (fuse->n == 0) return B200_OK;
    MsmAccSets<F> sets;
    u32 agrid = 0;
    for (int i = 0; i < fuse->n; i++) {
        const MsmPending<F> &p = fuse->job[i];
        MsmAccSet<F> &a = sets.s[i];
        a.bases = p.bases; a.entries = p.d_entries; a.off = p.d_hist; a.ntasks_ptr = p.d_plan + 2; a.hot_base = p.d_hot_base;
        a.tasks = p.d_tasks; a.buckets = p.d_buckets; a.partial = p.d_partial; a.CAP = p.g.CAP;
        if (p.agrid > agrid) agrid = p.agrid;
    }
    for (int i = fuse->n; i < MSM_MAX_FUSE; i++) sets.s[i] = sets.s[0];
    const bool g2 = sizeof(F) != 32;
    phase_begin(ctx, g2 ? PH_MSM_ACCUM_G2 : PH_MSM_ACCUM);
    u32 npts = 0;
    for (int i = 0; i < fuse->n; i++) npts = fuse->job[i].npoints > npts ? fuse->job[i].npoints : npts;
    const bool lock = g2 ? ctx->opt_lockstep_g2 > 0 : msm_lockstep_g1(ctx, npts);
    if (g2 && lock) B200_LAUNCH(ctx, (k_msm_accumulate_sets<F, 2, true>), dim3(agrid, fuse->n), 128, 0, sets);
    else if (g2) B200_LAUNCH(ctx, (k_msm_accumulate_sets<F, 2>), dim3(agrid, fuse->n), 128, 0, sets);
    else if (lock) B200_LAUNCH(ctx, (k_msm_accumulate_sets<F, 4, true>), dim3(agrid, fuse->n), 128, 0, sets);
    else B200_LAUNCH(ctx, (k_msm_accumulate_sets<F, 4>), dim3(agrid, fuse->n), 128, 0, sets);
    phase_end(ctx);
    for (int i = 0; i < fuse->n; i++) B200_TRY(msm_post_impl(ctx, fuse->job[i]));
    fuse->n = 0;
    return B200_OK;
}

template <class F>
int msm_enqueue_impl(Ctx *ctx, const void *d_bases_v, const void *d_scalars_v, uint32_t scalar_size, uint64_t n, int slot,
                     const MsmTable<F> *table, bool reuse_sort, bool tail, int ws, cudaStream_t sort_stream,
                     MsmFuse<F> *fuse = nullptr) {
    typedef Xyzz<F> Pt;
    if (slot < 0 || slot >= Ctx::MSM_SLOTS || ws < 0 || ws >= Ctx::SORT_WS) { ctx->err = "msm: bad result slot / workspace"; return B200_ERR_ARG; }
    Ctx::SlotInfo &si = ctx->slot_info[slot];
    si.used = false;
    if (n == 0) { si.nwin_b = 0; si.used = true; return B200_OK; }
    if (scalar_size == 0 || scalar_size > 32) { ctx->err = "msm: scalar_size must be 1..32 bytes"; return B200_ERR_ARG; }
    if (!d_bases_v || !d_scalars_v) { ctx->err = "msm: null input"; return B200_ERR_ARG; }
    const bool pre = table && table->tbl;
    const Affine<F> *d_bases = pre ? table->tbl : static_cast<const Affine<F> *>(d_bases_v);
    const uint8_t *d_scalars = static_cast<const uint8_t *>(d_scalars_v);
    if (pre && (n != table->n || scalar_size != 32 || n > MSM_MAX_BATCH || (ctx->opt_max_batch_log2 >= 4 && n > (1ull << ctx->opt_max_batch_log2)))) { ctx->err = "msm: table geometry mismatch"; return B200_ERR_ARG; }

    const u32 batch_cap = (ctx->opt_max_batch_log2 >= 4 && ctx->opt_max_batch_log2 <= 24) ? (1u << ctx->opt_max_batch_log2) : MSM_MAX_BATCH;
    const u32 batch_max = n < batch_cap ? (u32)n : batch_cap;
    int c = pre ? table->c : msm_auto_c(n);
    if (!pre && ctx->force_c >= 4 && ctx->force_c <= 20) c = ctx->force_c;
    MsmGeom g = msm_geometry(batch_max, scalar_size, c, pre, ctx->opt_target_tasks_log2, tail, ctx->opt_reduce_l_tail);
    {   // experiments: reduce-segment length override ("reduce_l" both groups, "reduce_l_g2" G2 only); power of two
        int ov = (sizeof(F) != 32 && ctx->opt_reduce_l_g2 > 0) ? ctx->opt_reduce_l_g2 : ctx->opt_reduce_l;
        if (ov > 0 && (ov & (ov - 1)) == 0 && (u32)ov <= g.nbk && !(tail && ov > 16)) {
            g.L = (u32)ov;
            g.nseg = g.nbk / g.L;
            g.nplanes = 0;
            while ((1u << g.nplanes) < g.nseg) g.nplanes++;
        }
    }
    if (g.nwin > MSM_MAX_WIN) { ctx->err = "msm: too many windows"; return B200_ERR_ARG; }
    if (pre && (uint64_t)g.nwin * n >= (1ull << 31)) { ctx->err = "msm: table too large for 31-bit entries"; return B200_ERR_ARG; }
    if (reuse_sort && (n > batch_max)) { ctx->err = "msm: reuse_sort needs a single batch"; return B200_ERR_ARG; }
    const size_t hist_len = ((size_t)g.NB + 1 + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    const u32 ntiles = (u32)(hist_len / SCAN_TILE);
    const size_t max_entries = (size_t)batch_max * g.nwin;
    const size_t max_tasks = (max_entries < g.NB ? max_entries : g.NB) + max_entries / g.CAP + 1;
    const size_t max_partials = 2 * (max_entries / g.CAP) + 2;
    const u32 total_segs = g.nwin_b * g.nseg;
    const size_t PT_MAX = sizeof(G2Xyzz);   // G1 and G2 calls share the buffers: size for the larger point

    const int bb = slot;   // every result slot owns its bucket / partial / segment buffers
    B200_TRY(ctx_reserve(ctx, ctx->w_hist[ws], hist_len * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_cursor[ws], hist_len * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_scan_totals[ws], (size_t)ntiles * 4 + 16));
    B200_TRY(ctx_reserve(ctx, ctx->w_entries[ws], max_entries * 4 + 16));
    B200_TRY(ctx_reserve(ctx, ctx->w_buckets[bb], (size_t)g.NB * PT_MAX));
    B200_TRY(ctx_reserve(ctx, ctx->w_partial[bb], max_partials * PT_MAX));
    B200_TRY(ctx_reserve(ctx, ctx->w_hot[ws], ((size_t)3 * g.NB + 8) * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_plan[ws], (size_t)(2 * SCAN_TILE + 16) * 4));
    B200_TRY(ctx_reserve(ctx, ctx->w_tasks[ws], max_tasks * sizeof(uint2)));
    const u32 npl1 = g.nplanes + 1;
    const size_t MSM_SLOT_PTS = 512;   // (planes + 1) x bucket sets per result slot
    if ((size_t)npl1 * g.nwin_b > MSM_SLOT_PTS) { ctx->err = "msm: too many window x plane sums"; return B200_ERR_ARG; }
    B200_TRY(ctx_reserve(ctx, ctx->w_segs[bb], (2 * (size_t)total_segs + (size_t)128 * npl1 * g.nwin_b) * PT_MAX));
    B200_TRY(ctx_reserve(ctx, ctx->w_win, (size_t)Ctx::MSM_SLOTS * MSM_SLOT_PTS * PT_MAX));
    B200_TRY(ctx_pinned(ctx, (size_t)Ctx::MSM_SLOTS * MSM_SLOT_PTS * PT_MAX));

    u32 *d_hist = (u32 *)ctx->w_hist[ws].p, *d_cursor = (u32 *)ctx->w_cursor[ws].p, *d_totals = (u32 *)ctx->w_scan_totals[ws].p;
    u32 *d_entries = (u32 *)ctx->w_entries[ws].p;
    u32 *d_hot_base = (u32 *)ctx->w_hot[ws].p, *d_hot_list = d_hot_base + g.NB, *d_warm_list = d_hot_list + g.NB;
    // plan buffer: [0..4095] task-length histogram (then its scan), [4096..8191] scan cursors, [8192..] counters
    u32 *d_lenhist = (u32 *)ctx->w_plan[ws].p, *d_lencur = d_lenhist + SCAN_TILE, *d_plan = d_lenhist + 2 * SCAN_TILE;
    uint2 *d_tasks = (uint2 *)ctx->w_tasks[ws].p;
    Pt *d_buckets = (Pt *)ctx->w_buckets[bb].p, *d_partial = (Pt *)ctx->w_partial[bb].p;
    Pt *d_segs = (Pt *)ctx->w_segs[bb].p;
    Pt *d_win = (Pt *)((uint8_t *)ctx->w_win.p + (size_t)slot * MSM_SLOT_PTS * PT_MAX);
    Pt *h_win = (Pt *)((uint8_t *)ctx->pinned + (size_t)slot * MSM_SLOT_PTS * PT_MAX);
    cudaStream_t st = ctx->stream, side = ctx_side_stream(ctx, slot);
    if (!side) { ctx->err = "msm: cannot create the side stream"; return B200_ERR_CUDA; }
    const bool g2 = sizeof(F) != 32;
    const u32 tree_threads = ctx->opt_tree_threads > 0 ? (u32)ctx->opt_tree_threads : (g2 ? 64u : 128u);
    const bool acc_smem = ctx->opt_acc_smem == 1;  // measured on B200: registers win for both groups (2 = staged G2 variant)
    const size_t acc_smem_bytes = 128 * sizeof(Pt);
    static const int env_g2_minb = getenv("B200_G2_MINB") ? atoi(getenv("B200_G2_MINB")) : 0;   // experiments
    const int g2_minb = ctx->opt_g2_minb ? ctx->opt_g2_minb : env_g2_minb ? env_g2_minb : (B200_G2_HOT_CALLS ? 3 : 2);

    // this slot's buffers may still be read by the side-stream reduction of its previous MSM
    if (ctx->slot_busy[slot]) B200_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ctx->ev_done[slot], 0));
    B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_buckets, 0, (size_t)g.NB * sizeof(Pt), st));

    int batch_idx = 0;
    for (uint64_t base = 0; base < n; base += batch_max, batch_idx++) {
        const u32 nb = (u32)((n - base < batch_max) ? (n - base) : batch_max);
        const uint8_t *sc = d_scalars + (size_t)base * scalar_size;
        const Affine<F> *bs = pre ? d_bases : d_bases + base;
        const u32 dgrid = (nb + 255) / 256;
        const u32 tstride = pre ? table->n : 0;
        const int add_existing = batch_idx > 0 ? 1 : 0;

        if (!reuse_sort) {
            cudaStream_t ss = sort_stream ? sort_stream : st;
            // side-stream merges of earlier MSMs may still read this workspace's offsets / plan
            for (int r = 0; r < Ctx::MSM_SLOTS; r++)
                if (ctx->sort_readers[ws] & (1u << r)) B200_CUDA_CHECK(ctx, cudaStreamWaitEvent(ss, ctx->ev_merge[r], 0));
            ctx->sort_readers[ws] = 0;
            // ... and so may the last accumulation that used it, if that is still queued on the main stream
            if (ss != st && ctx->ws_acc_pending[ws]) B200_CUDA_CHECK(ctx, cudaStreamWaitEvent(ss, ctx->ev_ws_acc[ws], 0));
            phase_begin(ctx, PH_MSM_SORT, ss);
            B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_hist, 0, hist_len * 4, ss));
            B200_CUDA_CHECK(ctx, cudaMemsetAsync(d_lenhist, 0, (size_t)(2 * SCAN_TILE + 16) * 4, ss));
            B200_LAUNCH_ON(ctx, ss, k_msm_digits<false>, dgrid, 256, 0, sc, scalar_size, nb, g.c, g.nwin, g.nbk, pre ? 1u : 0u, tstride, d_hist, (u32 *)nullptr);
            B200_LAUNCH_ON(ctx, ss, k_scan_tile, ntiles, 1024, 0, d_hist, d_totals);
            B200_LAUNCH_ON(ctx, ss, k_scan_totals, 1, 1024, 0, d_totals, ntiles);
            B200_LAUNCH_ON(ctx, ss, k_scan_add, ntiles, 1024, 0, d_hist, d_totals, d_cursor);
            B200_LAUNCH_ON(ctx, ss, k_msm_digits<true>, dgrid, 256, 0, sc, scalar_size, nb, g.c, g.nwin, g.nbk, pre ? 1u : 0u, tstride, d_cursor, d_entries);
            // task plan: histogram of task lengths (descending), scan, placement
            B200_LAUNCH_ON(ctx, ss, k_msm_plan_count, (g.NB + MSM_PLAN_THREADS - 1) / MSM_PLAN_THREADS, MSM_PLAN_THREADS, 0, d_hist, g.NB, g.CAP, d_lenhist, d_plan, d_hot_base, d_hot_list, d_warm_list, ctx->opt_warm_max > 0 ? (u32)ctx->opt_warm_max : MSM_WARM_MAX);
            B200_LAUNCH_ON(ctx, ss, k_scan_tile, 1, 1024, 0, d_lenhist, d_plan + 2);      // d_plan[2] = number of tasks
            B200_CUDA_CHECK(ctx, cudaMemcpyAsync(d_lencur, d_lenhist, SCAN_TILE * 4, cudaMemcpyDeviceToDevice, ss));
            B200_LAUNCH_ON(ctx, ss, k_msm_plan_place, (g.NB + MSM_PLAN_THREADS - 1) / MSM_PLAN_THREADS, MSM_PLAN_THREADS, 0, d_hist, g.NB, g.CAP, d_lencur, d_tasks);
            phase_end(ctx, ss);
            B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_sort[ws], ss));   // "sorted": also lets callers chain other streams
            if (ss != st) B200_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ctx->ev_sort[ws], 0));
        }

        const size_t ent = (size_t)nb * g.nwin;
        const size_t tasks_ub = (ent < g.NB ? ent : g.NB) + ent / g.CAP + 1;
        const u32 agrid = (u32)((tasks_ub + 127) / 128);
        MsmPending<F> pend;
        pend.g = g; pend.slot = slot; pend.ws = ws; pend.tree_threads = tree_threads; pend.agrid = agrid;
        pend.add_existing = add_existing; pend.last_batch = base + batch_max >= n; pend.bases = bs; pend.npoints = nb;
        pend.d_hist = d_hist; pend.d_plan = d_plan; pend.d_hot_base = d_hot_base; pend.d_hot_list = d_hot_list;
        pend.d_warm_list = d_warm_list; pend.d_entries = d_entries; pend.d_tasks = d_tasks;
        pend.d_buckets = d_buckets; pend.d_partial = d_partial; pend.d_segs = d_segs; pend.d_win = d_win; pend.h_win = h_win;
        pend.side = side;
        // a single-batch MSM with the default kernel variant can join a fused launch (msm_fuse_flush)
        if (fuse && n <= batch_max && ctx->opt_acc_smem <= 0 && fuse->n < MSM_MAX_FUSE && !(g2 && g2_minb == 3)) {
            fuse->job[fuse->n++] = pend;
            si.used = false;
            return B200_OK;
        }
        phase_begin(ctx, g2 ? PH_MSM_ACCUM_G2 : PH_MSM_ACCUM);
        {
            // variants: running sum in registers (default) or in shared memory
            void (*kacc)(const Affine<F> *, const u32 *, const u32 *, const uint2 *, const u32 *, u32, const u32 *, Pt *, Pt *, int);
            size_t smem = 0;
            u32 nthr = 128;
            if (ctx->opt_acc_smem == 2 && g2) {   // experimental: running sum + points staged in shared memory
                kacc = k_msm_accumulate_staged<F, 3>;
                smem = (size_t)128 * (sizeof(Pt) + 2 * sizeof(Affine<F>));
                if (!ctx->attr_acc_staged) {   // per device, hence per context
                    B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(kacc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    ctx->attr_acc_staged = true;
                }
            } else if (acc_smem) { kacc = g2 ? k_msm_accumulate<F, true, 3> : k_msm_accumulate<F, true, 4>; smem = acc_smem_bytes; }
            else if (g2) {
                if (ctx->opt_lockstep_g2 == 2) { kacc = k_msm_accumulate<F, false, 1, true, 256>; nthr = 256; }   // one 8-warp CTA per SM
                else if (ctx->opt_lockstep_g2 == 1) kacc = g2_minb == 3 ? k_msm_accumulate<F, false, 3, true> : k_msm_accumulate<F, false, 2, true>;
                else kacc = g2_minb == 3 ? k_msm_accumulate<F, false, 3> : k_msm_accumulate<F, false, 2>;
            } else if (ctx->opt_lockstep_g1 > 0) { kacc = k_msm_accumulate<F, false, 4, true>; }   // a single MSM: only on request
            else { kacc = k_msm_accumulate<F, false, 4>; }   // capped at 128 registers: four CTAs (16 warps) per SM
            B200_LAUNCH(ctx, kacc, (agrid * 128 + nthr - 1) / nthr, nthr, smem, bs, d_entries, d_hist, d_tasks, d_plan + 2, g.CAP, d_hot_base, d_buckets, d_partial, add_existing);
        }
        phase_end(ctx);
        B200_TRY(msm_post_impl(ctx, pend));
    }
    return B200_OK;
}

template <class F>
int msm_collect_impl(Ctx *ctx, int slot, Xyzz<F> *out_host) {
    typedef Xyzz<F> Pt;
    if (slot < 0 || slot >= Ctx::MSM_SLOTS || !ctx->slot_info[slot].used) { ctx->err = "msm: nothing enqueued in this slot"; return B200_ERR_ARG; }
    Ctx::SlotInfo &si = ctx->slot_info[slot];
    si.used = false;
    if (si.nwin_b == 0) { *out_host = Pt::zero(); return B200_OK; }
    B200_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->ev_done[slot]));
    ctx->slot_busy[slot] = false;
    const Pt *h_win = (const Pt *)((const uint8_t *)ctx->pinned + (size_t)slot * 512 * sizeof(G2Xyzz));
    // per bucket set: W + L * sum_t 2^t plane_t, then Horner over the windows (multiexp.cpp:137-141), all on the
    // host's 4x64 field
    if (sizeof(F) == 32) host_msm_finish_g1(h_win, si.nwin_b, si.nplanes, si.L, si.nwin, si.c, out_host);
    else host_msm_finish_g2(h_win, si.nwin_b, si.nplanes, si.L, si.nwin, si.c, out_host);
    return B200_OK;
}

}  // namespace b200
