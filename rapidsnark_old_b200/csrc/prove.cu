// Resident zkey + the fused prove path, NTT C-ABI front ends and the fixed-base table generator.
//
// b200_zkey_upload mirrors Groth16::makeProver (src/groth16.cpp:9-46): it takes the raw zkey section
// pointers, but instead of borrowing them it copies this GPU's point-range shard of every table to
// HBM once, and repacks the 44-byte coefficient records (groth16.hpp:27-35) into CSR by (matrix, row)
// so that the a/b build (groth16.cpp:62-85, 1024 striped locks on the CPU) is a lock-free row gather.
// b200_prove_msms is groth16.cpp:52-207: a,b,c -> 3 x (ifft, coset twist, fft) -> h -> five MSMs.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include "ctx.cuh"
#include "memops.cuh"
#include "scan.cuh"

namespace b200 {
int ntt_natural(Ctx *ctx, Fr *d_a, uint64_t n, bool inverse);
int h_pipeline(Ctx *ctx, Fr *d_a, Fr *d_b, Fr *d_c, uint64_t n, int nstreams);
int h_transforms(Ctx *ctx, Fr *d_a, Fr *d_b, Fr *d_c, uint64_t n, int nstreams, unsigned poly_mask);
int h_combine(Ctx *ctx, Fr *d_a, const Fr *d_b, const Fr *d_c, uint64_t n);
int build_abc(Ctx *ctx, const Fr *d_wtns, const u32 *d_row_a, const u32 *d_row_b, const u32 *d_sig, const Fr *d_coef,
              u32 n, Fr *d_a, Fr *d_b, Fr *d_c);
}  // namespace b200

using namespace b200;


struct Range { uint64_t lo, hi; };

struct b200_zkey {
    Ctx *ctx;
    u32 n_vars, n_public, domain_size;
    u64 n_coefs;
    Range rA, rC, rH;       // this shard's index ranges (A, B1, B2 share rA)
    bool stage1_done = false, stage1_combined = false;   // b200_prove_begin / b200_prove_finish pairing
    bool sharded = false;   // shard_count > 1: small MSMs, the (replicated) H pipeline is the critical path
    bool has_coefs = true;  // false: uploaded without the coefficient section (a shard that only receives a, b, c)
    u32 *d_row_a = nullptr, *d_row_b = nullptr, *d_sig = nullptr;
    Fr *d_coef = nullptr;
    G1Affine *d_A = nullptr, *d_B1 = nullptr, *d_C = nullptr, *d_H = nullptr;   // plain shard slices, or
    G2Affine *d_B2 = nullptr;                                                   // per-window tables (t*.tbl)
    MsmTableRaw tA, tB1, tB2, tC, tH;
    Fr *d_wtns = nullptr, *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;
    // b200_zkey_share: a view borrows the parent's resident tables / CSR and owns only its per-proof buffers
    b200_zkey *parent = nullptr;
    int views = 0;          // live views of this (parent) zkey
    bool released = false;  // b200_zkey_free was called on a parent that still has views: freed with the last view
};

static Range shard_range(uint64_t len, u32 idx, u32 cnt) {
    Range r;
    r.lo = len * idx / cnt;
    r.hi = len * (idx + 1) / cnt;
    return r;
}

// Window width of the resident tables for `len` points: the zkey's scalars (witness, h) are field elements below
// r < 2^254, so ceil(254 / c) windows carry digits; cost = one mixed addition per digit plus two full additions
// (~2.8 mixed) per bucket of the shared set.  Widths that leave only a few bits for the top window are penalised:
// they pile len / 2^bits entries on each of a handful of buckets, whose partial sums then fold through long
// serial chains (c = 18: 2 bits, c = 19: 7 bits - measured at 4 and 2 shards of a 2^20 circuit).  2^20 points -> 20
// (measured best of 16..20 on B200), 2^17..2^19 -> 17.  Any width gives the same result.
static int table_window_bits(uint64_t len) {
    int best = 11;
    double best_cost = 0;
    for (int cb = 11; cb <= 20; cb++) {
        const int eff = (254 + cb - 1) / cb, top_bits = 254 - cb * (eff - 1);
        double cost = (double)eff * (double)len + 2.8 * (double)(1u << (cb - 1));
        if (top_bits <= 8) cost += 0.1 * (double)len;
        if (cb == 11 || cost < best_cost) { best = cb; best_cost = cost; }
    }
    return best;
}

// ---- ingestion (reference: src/binfile_utils.cpp:14-64 reads the file into a malloc'd copy) ------------
// The zkey sections arrive as pointers into the caller's mmap of the file: pageable memory.  Two ways to the device:
// plain cudaMemcpyAsync (the driver stages through its own pinned buffers), or - B200_STAGING=1 - a few host threads
// that copy 8 MB chunks into their own pinned buffers and issue the H2D copies from those.  Measured on the B200 box
// at 2^20 (0.54 GB zkey, page cache warm, profiles/r02_cli_bench.json): plain 58 ms for the whole upload including
// the device-side CSR build, staged 93 ms - allocating the pinned buffers costs more than the overlap saves - so the
// plain path is the default.
static const size_t STAGE_CHUNK = 8u << 20;
static const int STAGE_THREADS = 4;

struct Stager {
    void *buf[STAGE_THREADS] = {nullptr};
    cudaStream_t st[STAGE_THREADS] = {nullptr};
    cudaEvent_t ev[STAGE_THREADS] = {nullptr};
    bool ok = false;
    bool init() {
        if (ok) return true;
        for (int i = 0; i < STAGE_THREADS; i++) {
            if (cudaHostAlloc(&buf[i], STAGE_CHUNK, cudaHostAllocDefault) != cudaSuccess) return false;
            if (cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        ok = true;
        return true;
    }
    void release() {
        for (int i = 0; i < STAGE_THREADS; i++) {
            if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
            if (ev[i]) cudaEventDestroy(ev[i]);
            if (buf[i]) cudaFreeHost(buf[i]);
            buf[i] = nullptr; st[i] = nullptr; ev[i] = nullptr;
        }
        ok = false;
    }
};

struct CopyJob { void *dst; const void *src; size_t bytes; };

// all jobs, chunk by chunk, over the staging threads; returns when every byte is on the device
static int staged_upload(Ctx *c, const std::vector<CopyJob> &jobs) {
    size_t total = 0;
    for (auto &j : jobs) total += j.bytes;
    static const bool on = getenv("B200_STAGING") != nullptr && atoi(getenv("B200_STAGING")) != 0;
    Stager sg;
    if (!on || total < (64u << 20) || !sg.init()) {
        sg.release();
        cudaGetLastError();
        for (auto &j : jobs)
            if (j.bytes) B200_CUDA_CHECK(c, cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyHostToDevice, c->stream));
        B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        return B200_OK;
    }
    struct Chunk { uint8_t *dst; const uint8_t *src; size_t bytes; };
    std::vector<Chunk> chunks;
    for (auto &j : jobs)
        for (size_t o = 0; o < j.bytes; o += STAGE_CHUNK)
            chunks.push_back({(uint8_t *)j.dst + o, (const uint8_t *)j.src + o, std::min(STAGE_CHUNK, j.bytes - o)});
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    const int dev = c->device;
    auto work = [&](int t) {
        cudaSetDevice(dev);
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= chunks.size() || failed.load()) break;
            if (cudaEventSynchronize(sg.ev[t]) != cudaSuccess) { failed = 1; break; }   // this buffer's last copy is out
            memcpy(sg.buf[t], chunks[i].src, chunks[i].bytes);
            if (cudaMemcpyAsync(chunks[i].dst, sg.buf[t], chunks[i].bytes, cudaMemcpyHostToDevice, sg.st[t]) != cudaSuccess ||
                cudaEventRecord(sg.ev[t], sg.st[t]) != cudaSuccess) { failed = 1; break; }
        }
        cudaStreamSynchronize(sg.st[t]);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < STAGE_THREADS; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    sg.release();
    if (failed.load()) { c->err = "zkey_upload: staged host-to-device copy failed"; cudaGetLastError(); return B200_ERR_CUDA; }
    return B200_OK;
}

template <class T>
static int alloc_slice(Ctx *c, T **dst, const void *src, Range r, std::vector<CopyJob> &jobs) {
    size_t cnt = (size_t)(r.hi - r.lo);
    B200_CUDA_CHECK(c, cudaMalloc((void **)dst, std::max<size_t>(cnt, 1) * sizeof(T)));
    if (cnt) jobs.push_back({*dst, (const uint8_t *)src + (size_t)r.lo * sizeof(T), cnt * sizeof(T)});
    return B200_OK;
}

// ---- CSR build of the coefficient records on the device (reference: groth16.cpp:62-85 walks the packed 44-byte
// records {u32 m, u32 c, u32 s, Fr coef} with 1024 striped locks).  Rows of A then rows of B share one offset array:
// off[m * n + row]; the order of the entries inside a row is arbitrary - their sum is exact field arithmetic.
static __global__ void __launch_bounds__(256) k_coef_count(const u32 *__restrict__ rec, u64 n_coefs, u32 n, u32 n_vars,
                                                            u32 *__restrict__ cnt, u32 *__restrict__ bad) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_coefs) return;
    const u32 *r = rec + i * 11;
    u32 m = r[0], row = r[1], sg = r[2];
    if (m > 1 || row >= n || sg >= n_vars) { atomicOr(bad, 1u); return; }
    atomicAdd(&cnt[(size_t)m * n + row], 1u);
}

static __global__ void __launch_bounds__(256) k_coef_place(const u32 *__restrict__ rec, u64 n_coefs, u32 n,
                                                            u32 *__restrict__ cursor, u32 *__restrict__ sig, Fr *__restrict__ coef) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_coefs) return;
    const u32 *r = rec + i * 11;
    u32 m = r[0], row = r[1];
    u32 k = atomicAdd(&cursor[(size_t)m * n + row], 1u);
    sig[k] = r[2];
    Fr v;
#pragma unroll
    for (int j = 0; j < 8; j++) v.v[j] = r[3 + j];
    st_struct(coef + k, v);
}

// ---- fixed-base tables (synthetic zkey generation; not on the prove path) ---------------------------
template <class F>
__global__ void k_fb_window_bases(Affine<F> base, Xyzz<F> *__restrict__ wb) {  // wb[w] = 2^(8w) * base
    u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    Xyzz<F> p = Xyzz<F>::from_affine(base);
    for (u32 i = 0; i < 8 * w; i++) p = ec_dbl(p);
    st_struct(wb + w, p);
}

template <class F>
DEVFN Affine<F> to_affine_1inv(const Xyzz<F> &p) {
    Affine<F> r;
    if (p.is_zero()) { r.x = F::zero(); r.y = F::zero(); return r; }
    F i = finv(fmul(p.zz, p.zzz));
    r.x = fmul(p.x, fmul(i, p.zzz));
    r.y = fmul(p.y, fmul(i, p.zz));
    return r;
}

template <class F>
__global__ void k_fb_table(const Xyzz<F> *__restrict__ wb, Affine<F> *__restrict__ tbl) {  // tbl[w*256+d] = d * wb[w]
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 32 * 256) return;
    u32 w = g >> 8, d = g & 255;
    Xyzz<F> p = ld_struct(wb + w);
    Xyzz<F> r = ec_mul(p, &d, 1);
    st_struct(tbl + g, to_affine_1inv(r));
}

template <class F>
__global__ void __launch_bounds__(128) k_fb_mul(const Affine<F> *__restrict__ tbl, const uint8_t *__restrict__ scalars,
                                                u64 n, Affine<F> *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *s = scalars + i * 32;
    Xyzz<F> acc = Xyzz<F>::zero();
    for (int w = 0; w < 32; w++) {
        u32 d = s[w];
        if (d) {
            Affine<F> p = ldg_struct(tbl + w * 256 + d);
            ec_madd(acc, p);
        }
    }
    st_struct(out + i, to_affine_1inv(acc));
}

template <class F>
static int fixed_base_run(Ctx *c, const void *base_affine, const void *scalars, uint64_t n, void *out) {
    if (!base_affine || (n && (!scalars || !out))) { c->err = "fixed_base: null argument"; return B200_ERR_ARG; }
    if (n == 0) return B200_OK;
    Affine<F> base;
    memcpy(&base, base_affine, sizeof base);
    Xyzz<F> *d_wb = nullptr;
    Affine<F> *d_tbl = nullptr, *d_out = nullptr;
    uint8_t *d_sc = nullptr;
    B200_CUDA_CHECK(c, cudaMalloc(&d_wb, 32 * sizeof(Xyzz<F>)));
    B200_CUDA_CHECK(c, cudaMalloc(&d_tbl, 32 * 256 * sizeof(Affine<F>)));
    B200_CUDA_CHECK(c, cudaMalloc(&d_sc, n * 32));
    B200_CUDA_CHECK(c, cudaMalloc(&d_out, n * sizeof(Affine<F>)));
    B200_CUDA_CHECK(c, cudaMemcpyAsync(d_sc, scalars, n * 32, cudaMemcpyHostToDevice, c->stream));
    B200_LAUNCH(c, k_fb_window_bases<F>, 1, 32, 0, base, d_wb);
    B200_LAUNCH(c, k_fb_table<F>, 32 * 256 / 64, 64, 0, d_wb, d_tbl);
    B200_LAUNCH(c, k_fb_mul<F>, (u32)((n + 127) / 128), 128, 0, d_tbl, d_sc, n, d_out);
    B200_CUDA_CHECK(c, cudaMemcpyAsync(out, d_out, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_wb); cudaFree(d_tbl); cudaFree(d_sc); cudaFree(d_out);
    return B200_OK;
}

extern "C" {

// ---------------------------------------------------------------------------------------------- NTT
int b200_ntt_fr_dev(b200_ctx *h, void *d_a, uint64_t n, int inverse) {
    if (!h || !d_a) return B200_ERR_ARG;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    phase_reset(c);
    phase_begin(c, PH_NTT);
    int rc = ntt_natural(c, (Fr *)d_a, n, inverse != 0);
    phase_end(c);
    if (rc != B200_OK) return rc;
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    phase_collect(c);
    return B200_OK;
}

int b200_ntt_fr(b200_ctx *h, void *a_host, uint64_t n, int inverse) {
    if (!h || !a_host) return B200_ERR_ARG;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    if (n == 0 || (n & (n - 1))) { c->err = "ntt: n must be a power of two"; return B200_ERR_ARG; }
    B200_TRY(ctx_reserve(c, c->w_ntt, n * 32));
    phase_reset(c);
    phase_begin(c, PH_H2D);
    B200_CUDA_CHECK(c, cudaMemcpyAsync(c->w_ntt.p, a_host, n * 32, cudaMemcpyHostToDevice, c->stream));
    phase_end(c);
    phase_begin(c, PH_NTT);
    int rc = ntt_natural(c, (Fr *)c->w_ntt.p, n, inverse != 0);
    phase_end(c);
    if (rc != B200_OK) return rc;
    B200_CUDA_CHECK(c, cudaMemcpyAsync(a_host, c->w_ntt.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    phase_collect(c);
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------- zkey
static void zkey_free_tables(b200_zkey *zk) {
    void *all[] = {zk->d_row_a /* d_row_b is part of it */, zk->d_sig, zk->d_coef, zk->d_A, zk->d_B1, zk->d_C, zk->d_H, zk->d_B2};
    for (void *p : all) if (p) cudaFree(p);
    delete zk;
}

void b200_zkey_free(b200_zkey *zk) {
    if (!zk || zk->released) return;
    cudaSetDevice(zk->ctx->device);
    cudaStreamSynchronize(zk->ctx->stream);
    void *work[] = {zk->d_wtns, zk->d_a, zk->d_b, zk->d_c};
    for (void *p : work) if (p) cudaFree(p);
    zk->d_wtns = zk->d_a = zk->d_b = zk->d_c = nullptr;
    if (zk->parent) {                       // a view: give the tables back
        b200_zkey *par = zk->parent;
        delete zk;
        if (--par->views == 0 && par->released) zkey_free_tables(par);
        return;
    }
    if (zk->views > 0) { zk->released = true; return; }   // tables stay until the last view is gone
    zkey_free_tables(zk);
}

int b200_zkey_share(b200_ctx *h, b200_zkey *src, b200_zkey **out) {
    if (!h || !src || !out) return B200_ERR_ARG;
    Ctx *c = &h->c;
    *out = nullptr;
    b200_zkey *root = src->parent ? src->parent : src;
    if (root->released) { c->err = "zkey_share: the source zkey has been freed"; return B200_ERR_ARG; }
    if (c->device != root->ctx->device) { c->err = "zkey_share: the view's context must be on the source's device"; return B200_ERR_ARG; }
    cudaSetDevice(c->device);
    b200_zkey *zk = new b200_zkey(*root);
    zk->ctx = c;
    zk->parent = root;
    zk->views = 0;
    zk->released = false;
    zk->stage1_done = zk->stage1_combined = false;
    zk->d_wtns = zk->d_a = zk->d_b = zk->d_c = nullptr;
    const size_t n = root->domain_size;
    cudaError_t e = cudaMalloc(&zk->d_wtns, (size_t)root->n_vars * sizeof(Fr));
    if (e == cudaSuccess) e = cudaMalloc(&zk->d_a, n * sizeof(Fr));
    if (e == cudaSuccess) e = cudaMalloc(&zk->d_b, n * sizeof(Fr));
    if (e == cudaSuccess) e = cudaMalloc(&zk->d_c, n * sizeof(Fr));
    if (e != cudaSuccess) {
        c->err = std::string("zkey_share: ") + cudaGetErrorString(e);
        void *work[] = {zk->d_wtns, zk->d_a, zk->d_b, zk->d_c};
        for (void *p : work) if (p) cudaFree(p);
        delete zk;
        return B200_ERR_CUDA;
    }
    root->views++;
    *out = zk;
    return B200_OK;
}

static bool cnt_is_one(const b200_zkey_desc *d) { return d->shard_count <= 1; }

int b200_zkey_upload(b200_ctx *h, const b200_zkey_desc *d, b200_zkey **out) {
    if (!h || !d || !out) return B200_ERR_ARG;
    Ctx *c = &h->c;
    *out = nullptr;
    cudaSetDevice(c->device);
    // coefs may be NULL for a shard that never builds a, b, c itself (it receives them: b200_prove_begin with
    // poly_mask = 0) - five of eight ranks then skip 44 bytes per coefficient of host reads, upload and CSR build
    if ((!d->coefs && cnt_is_one(d)) || !d->points_a || !d->points_b1 || !d->points_b2 || !d->points_c || !d->points_h) {
        c->err = "zkey_upload: null section pointer"; return B200_ERR_ARG;
    }
    const u32 n = d->domain_size;
    if (n == 0 || (n & (n - 1))) { c->err = "zkey_upload: domain_size must be a power of two"; return B200_ERR_ARG; }
    if ((uint64_t)n * 2 > (1ull << 28)) { c->err = "Domain size too big for the curve"; return B200_ERR_RANGE; }
    if (d->n_vars == 0 || d->n_public + 1 > d->n_vars) { c->err = "zkey_upload: bad n_vars / n_public"; return B200_ERR_ARG; }
    const u32 cnt = d->shard_count ? d->shard_count : 1;
    if (d->shard_index >= cnt) { c->err = "zkey_upload: shard_index >= shard_count"; return B200_ERR_ARG; }

    b200_zkey *zk = new b200_zkey();
    zk->ctx = c;
    zk->n_vars = d->n_vars; zk->n_public = d->n_public; zk->domain_size = n; zk->n_coefs = d->n_coefs;
    zk->sharded = cnt > 1;
    if (d->shard_den) {   // explicit (uneven) bounds
        if (d->shard_lo_num > d->shard_hi_num || d->shard_hi_num > d->shard_den) { delete zk; c->err = "zkey_upload: bad shard bounds"; return B200_ERR_ARG; }
        zk->rA.lo = (uint64_t)d->n_vars * d->shard_lo_num / d->shard_den;
        zk->rA.hi = (uint64_t)d->n_vars * d->shard_hi_num / d->shard_den;
        zk->rH.lo = (uint64_t)n * d->shard_lo_num / d->shard_den;
        zk->rH.hi = (uint64_t)n * d->shard_hi_num / d->shard_den;
    } else {
        zk->rA = shard_range(d->n_vars, d->shard_index, cnt);
        zk->rH = shard_range(n, d->shard_index, cnt);
    }
    zk->rC = zk->rA;   // the C table is padded to the witness indexing (see below)
    int rc = B200_OK;
    auto fail = [&](int code) { b200_zkey_free(zk); return code; };
#define ZK_TRY(x) do { rc = (x); if (rc != B200_OK) return fail(rc); } while (0)
#define ZK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { c->err = std::string("zkey_upload: ") + cudaGetErrorString(e_); return fail(B200_ERR_CUDA); } } while (0)
    // every section goes to the device through the staged uploader; the coefficient records as they are in the file
    // (the CSR arrays are built from them on the device, below)
    std::vector<CopyJob> jobs;
    u32 *d_rec = nullptr;                       // raw 44-byte records (section 4 after the u32 count: groth16.cpp:38)
    const bool have_coefs = d->coefs != nullptr && d->n_coefs > 0;
    if (have_coefs) {
        ZK_CUDA(cudaMalloc(&d_rec, (size_t)d->n_coefs * 44));
        jobs.push_back({d_rec, (const uint8_t *)d->coefs + 4, (size_t)d->n_coefs * 44});
    }
    ZK_TRY(alloc_slice(c, &zk->d_A, d->points_a, zk->rA, jobs));
    ZK_TRY(alloc_slice(c, &zk->d_B1, d->points_b1, zk->rA, jobs));
    ZK_TRY(alloc_slice(c, &zk->d_B2, d->points_b2, zk->rA, jobs));
    {   // C is stored aligned with the witness: (n_public + 1) leading points at infinity, so the A, B1, B2 and C
        // MSMs all read the same scalars and can share one digit sort (zero bases are skipped by the mixed add)
        const uint64_t lo = zk->rA.lo, hi = zk->rA.hi, skip = (uint64_t)d->n_public + 1;
        ZK_CUDA(cudaMalloc((void **)&zk->d_C, std::max<size_t>(hi - lo, 1) * sizeof(G1Affine)));
        ZK_CUDA(cudaMemsetAsync(zk->d_C, 0, std::max<size_t>(hi - lo, 1) * sizeof(G1Affine), c->stream));
        ZK_CUDA(cudaStreamSynchronize(c->stream));       // the staged copies below run on other streams
        const uint64_t from = std::max(lo, skip);
        if (hi > from)
            jobs.push_back({zk->d_C + (from - lo), (const uint8_t *)d->points_c + (from - skip) * sizeof(G1Affine),
                            (size_t)(hi - from) * sizeof(G1Affine)});
    }
    ZK_TRY(alloc_slice(c, &zk->d_H, d->points_h, zk->rH, jobs));
    if (int up = staged_upload(c, jobs)) { if (d_rec) cudaFree(d_rec); return fail(up); }
    // CSR arrays: off[0 .. n] = row starts of A, off[n .. 2n] = row starts of B (one exclusive scan over the 2n counts)
    {
        const size_t rows = 2 * (size_t)n + 1, rows_pad = (rows + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
        const u32 ntiles = (u32)(rows_pad / SCAN_TILE);
        const size_t nc = have_coefs ? (size_t)d->n_coefs : 1;
        u32 *d_cur = nullptr, *d_tot = nullptr, *d_bad = nullptr;
        ZK_CUDA(cudaMalloc(&zk->d_row_a, rows_pad * 4));
        zk->d_row_b = zk->d_row_a + n;              // same allocation
        ZK_CUDA(cudaMalloc(&zk->d_sig, nc * 4));
        ZK_CUDA(cudaMalloc(&zk->d_coef, nc * sizeof(Fr)));
        ZK_CUDA(cudaMemsetAsync(zk->d_row_a, 0, rows_pad * 4, c->stream));
        if (have_coefs) {
            cudaError_t e = cudaMalloc(&d_cur, rows_pad * 4);
            if (e == cudaSuccess) e = cudaMalloc(&d_tot, (size_t)ntiles * 4 + 16);
            if (e == cudaSuccess) e = cudaMalloc(&d_bad, 4);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, c->stream);
            u32 bad = 0;
            if (e == cudaSuccess) {
                const u32 grid = (u32)((d->n_coefs + 255) / 256);
                k_coef_count<<<grid, 256, 0, c->stream>>>(d_rec, d->n_coefs, n, d->n_vars, zk->d_row_a, d_bad);
                k_scan_tile<<<ntiles, 1024, 0, c->stream>>>(zk->d_row_a, d_tot);
                k_scan_totals<<<1, 1024, 0, c->stream>>>(d_tot, ntiles);
                k_scan_add<<<ntiles, 1024, 0, c->stream>>>(zk->d_row_a, d_tot, d_cur);
                c->launches += 4;
                e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
                if (e == cudaSuccess && !bad) {
                    k_coef_place<<<grid, 256, 0, c->stream>>>(d_rec, d->n_coefs, n, d_cur, zk->d_sig, zk->d_coef);
                    c->launches++;
                    e = cudaStreamSynchronize(c->stream);
                }
                if (e == cudaSuccess) e = cudaGetLastError();
            }
            if (d_cur) cudaFree(d_cur);
            if (d_tot) cudaFree(d_tot);
            if (d_bad) cudaFree(d_bad);
            cudaFree(d_rec);
            d_rec = nullptr;
            if (e != cudaSuccess) { c->err = std::string("zkey_upload: ") + cudaGetErrorString(e); return fail(B200_ERR_CUDA); }
            if (bad) { c->err = "zkey_upload: coefficient record out of range"; return fail(B200_ERR_ARG); }
        }
        zk->has_coefs = have_coefs || d->n_coefs == 0;
    }
    // resident per-window tables (msm.cuh): all windows of an MSM share one bucket set
    {
        // window width of the tables: about log2(points) (measured on B200: 2^20 points -> c = 20 beats 16..19),
        // no tables at all for tiny shards where the plain multi-window path is just as fast
        uint64_t len_max = std::max<uint64_t>(zk->rA.hi - zk->rA.lo, zk->rH.hi - zk->rH.lo);
        int lg = 0;
        while ((2ull << lg) <= len_max) lg++;
        int pc = c->opt_precomp_c > 0 ? c->opt_precomp_c : table_window_bits(len_max);
        if (c->opt_precomp_c <= 0 && lg < 11) pc = 0;
        int rows = pc > 0 ? msm_table_windows(pc) : 0;
        uint64_t lenA = zk->rA.hi - zk->rA.lo, lenC = lenA, lenH = zk->rH.hi - zk->rH.lo;
        uint64_t need = (uint64_t)rows * ((2 * lenA + lenC + lenH) * 64 + lenA * 128);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        bool want = c->opt_precomp != 0 && pc >= 11 && pc <= 22 && rows <= 24 && need < (uint64_t)(free_b * 0.6) &&
                    (uint64_t)rows * lenH < (1ull << 31) && (uint64_t)rows * lenA < (1ull << 31) &&
                    // a table MSM is a single sort batch (msm_enqueue_impl): longer shards keep the plain tables and
                    // take the multi-batch path
                    lenA <= MSM_MAX_BATCH && lenH <= MSM_MAX_BATCH &&
                    !(c->opt_max_batch_log2 >= 4 && len_max > (1ull << c->opt_max_batch_log2));
        if (want) {
            auto build1 = [&](G1Affine **slice, uint64_t len, MsmTableRaw *t) -> int {
                if (len == 0) return B200_OK;
                G1Affine *tbl = nullptr;
                if (cudaMalloc(&tbl, (size_t)rows * len * sizeof(G1Affine)) != cudaSuccess) { cudaGetLastError(); return B200_OK; }
                int r = msm_g1_precompute(c, *slice, (u32)len, pc, tbl);
                if (r != B200_OK) { cudaFree(tbl); return r; }
                cudaStreamSynchronize(c->stream);
                cudaFree(*slice);
                *slice = tbl;                         // row 0 of the table is the original slice
                t->tbl = tbl; t->n = (u32)len; t->c = pc; t->nwin = rows;
                return B200_OK;
            };
            ZK_TRY(build1(&zk->d_A, lenA, &zk->tA));
            ZK_TRY(build1(&zk->d_B1, lenA, &zk->tB1));
            ZK_TRY(build1(&zk->d_C, lenC, &zk->tC));
            ZK_TRY(build1(&zk->d_H, lenH, &zk->tH));
            if (lenA) {
                G2Affine *tbl = nullptr;
                if (cudaMalloc(&tbl, (size_t)rows * lenA * sizeof(G2Affine)) == cudaSuccess) {
                    ZK_TRY(msm_g2_precompute(c, zk->d_B2, (u32)lenA, pc, tbl));
                    cudaStreamSynchronize(c->stream);
                    cudaFree(zk->d_B2);
                    zk->d_B2 = tbl;
                    zk->tB2.tbl = tbl; zk->tB2.n = (u32)lenA; zk->tB2.c = pc; zk->tB2.nwin = rows;
                } else cudaGetLastError();
            }
        }
    }
    ZK_CUDA(cudaMalloc(&zk->d_wtns, (size_t)d->n_vars * sizeof(Fr)));
    ZK_CUDA(cudaMalloc(&zk->d_a, (size_t)n * sizeof(Fr)));
    ZK_CUDA(cudaMalloc(&zk->d_b, (size_t)n * sizeof(Fr)));
    ZK_CUDA(cudaMalloc(&zk->d_c, (size_t)n * sizeof(Fr)));
    ZK_CUDA(cudaStreamSynchronize(c->stream));
#undef ZK_TRY
#undef ZK_CUDA
    *out = zk;
    return B200_OK;
}

// witness upload, then a,b,c -> h.  With `overlap` the H pipeline runs on its own stream (c->hstream) so that
// the witness MSMs, which do not depend on it, can start right after the upload; the caller makes the H MSM
// wait on c->ev_h.
// `slice_only`: a shard that builds none of a, b, c (it receives them) only reads its own range of the witness
static int wtns_upload(Ctx *c, b200_zkey *zk, const void *wtns, bool wtns_on_device, bool slice_only = false) {
    const size_t lo = slice_only ? (size_t)zk->rA.lo * 32 : 0;
    const size_t hi = slice_only ? (size_t)zk->rA.hi * 32 : (size_t)zk->n_vars * 32;
    phase_begin(c, PH_H2D);
    if (hi > lo)
        B200_CUDA_CHECK(c, cudaMemcpyAsync((uint8_t *)zk->d_wtns + lo, (const uint8_t *)wtns + lo, hi - lo,
                                            wtns_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    phase_end(c);
    return B200_OK;
}

// `after`: event on another stream the H pipeline has to wait for (witness uploaded / witness digits sorted).
// poly_mask / combine: see h_transforms (ntt.cu); the single-GPU path runs all three chains and the combine.
static int h_on_device(Ctx *c, b200_zkey *zk, bool overlap = false, cudaEvent_t after = nullptr, unsigned poly_mask = 7u,
                       bool combine = true) {
    cudaStream_t main_stream = c->stream;
    if (overlap) {
        if (after) B200_CUDA_CHECK(c, cudaStreamWaitEvent(c->hstream, after, 0));
        c->stream = c->hstream;           // build_abc / h_pipeline launch on c->stream
    }
    int rc = B200_OK;
    if (poly_mask) {                      // a rank that owns no chain receives all three polynomials
        phase_begin(c, PH_BUILD_AB);
        rc = build_abc(c, zk->d_wtns, zk->d_row_a, zk->d_row_b, zk->d_sig, zk->d_coef, zk->domain_size, zk->d_a, zk->d_b, zk->d_c);
        phase_end(c);
    }
    if (rc == B200_OK && (poly_mask || combine)) {
        phase_begin(c, PH_NTT);
        rc = h_transforms(c, zk->d_a, zk->d_b, zk->d_c, zk->domain_size, c->opt_h_streams ? c->opt_h_streams : 1, poly_mask);
        if (rc == B200_OK && combine) rc = h_combine(c, zk->d_a, zk->d_b, zk->d_c, zk->domain_size);
        phase_end(c);
    }
    if (overlap) {
        if (rc == B200_OK && cudaEventRecord(c->ev_h, c->hstream) != cudaSuccess) rc = B200_ERR_CUDA;
        c->stream = main_stream;
    }
    return rc;
}

int b200_h_scalars(b200_ctx *h, b200_zkey *zk, const void *wtns_host, void *h_out_host) {
    if (!h || !zk || !wtns_host || !h_out_host) return B200_ERR_ARG;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    if (!zk->has_coefs) { c->err = "h_scalars: this zkey was uploaded without the coefficient section"; return B200_ERR_ARG; }
    phase_reset(c);
    B200_TRY(wtns_upload(c, zk, wtns_host, false));
    B200_TRY(h_on_device(c, zk));
    B200_CUDA_CHECK(c, cudaMemcpyAsync(h_out_host, zk->d_a, (size_t)zk->domain_size * 32, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    phase_collect(c);
    return B200_OK;
}

// Stage 1: witness upload, the four witness MSMs enqueued, a/b/c built and the transform chains in poly_mask run on
// the H stream.  Stage 2 (prove_stage2): combine, H MSM, collect.  A single GPU runs both back to back; with the
// zkey sharded over several GPUs the caller exchanges the transformed polynomials between the stages (on the H
// stream), each rank having run only the chains it owns.
// After a failure part-way through a prove call work may still be queued on the main, H and side streams and result
// slots stay marked used / busy: wait for everything and forget it, so that the next call starts from a clean state
// instead of racing with stale kernels over d_wtns, d_a and the bucket buffers.
static void prove_drain(Ctx *c) {
    const std::string keep = c->err;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->hstream);
    for (int i = 0; i < 2; i++) if (c->hstream_bc[i]) cudaStreamSynchronize(c->hstream_bc[i]);
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) {
        if (c->side[i]) cudaStreamSynchronize(c->side[i]);
        c->slot_info[i].used = false;
        c->slot_busy[i] = false;
    }
    for (int w = 0; w < Ctx::SORT_WS; w++) { c->sort_readers[w] = 0; c->ws_acc_pending[w] = false; }
    msm_g1_fuse_free(c);               // MSMs still queued for a fused launch are dropped
    cudaGetLastError();
    c->segs.clear();
    c->ev_used = 0;
    c->err = keep;
}

static int prove_stage1_impl(Ctx *c, b200_zkey *zk, const void *wtns, bool wtns_on_device, unsigned poly_mask, bool combine);
static int prove_stage1(Ctx *c, b200_zkey *zk, const void *wtns, bool wtns_on_device, unsigned poly_mask, bool combine) {
    if (zk->stage1_done) { c->err = "prove_begin: the previous b200_prove_begin has not been finished"; return B200_ERR_ARG; }
    int rc = prove_stage1_impl(c, zk, wtns, wtns_on_device, poly_mask, combine);
    if (rc != B200_OK) prove_drain(c);
    return rc;
}

static int prove_stage1_impl(Ctx *c, b200_zkey *zk, const void *wtns, bool wtns_on_device, unsigned poly_mask, bool combine) {
    cudaSetDevice(c->device);
    if (poly_mask && !zk->has_coefs) { c->err = "prove: this zkey was uploaded without the coefficient section (it cannot build a, b, c)"; return B200_ERR_ARG; }
    phase_reset(c);
    B200_TRY(wtns_upload(c, zk, wtns, wtns_on_device, poly_mask == 0 && !combine));
    B200_CUDA_CHECK(c, cudaEventRecord(c->ev_h, c->stream));   // "witness uploaded"
    const uint8_t *w = (const uint8_t *)zk->d_wtns;
    // groth16.cpp:173 / :183 / :190 / :197 / :204, restricted to this shard's point range.  All five MSMs are
    // enqueued back to back (bucket reductions overlap the next accumulation on a side stream), then collected.
    // The four witness MSMs read the same scalars with the same geometry: sorted once, reused three times.
    const uint64_t lenA = zk->rA.hi - zk->rA.lo;
    const bool same_geom = (!zk->tA.tbl == !zk->tB1.tbl) && (!zk->tA.tbl == !zk->tB2.tbl) && (!zk->tA.tbl == !zk->tC.tbl) &&
                           lenA <= (1u << 24);
    // order: the G2 MSM first (its long bucket reduction then hides behind the G1 accumulations), H last
    B200_TRY(msm_g2_enqueue(c, zk->d_B2, w + zk->rA.lo * 32, 32, lenA, 1, &zk->tB2, false, false));
    // H pipeline on its own stream.  One GPU: started once the witness digits are sorted (the sort is atomics-bound
    // and would only be slowed down by the NTT kernels; the G2 accumulation that follows absorbs them).  A shard of
    // several: the transform chains, the exchange and the H MSM behind them are the critical path of the proof, the
    // MSMs are short - the chains start right after the witness upload, beside the sort.
    const bool h_early = c->opt_h_early > 0 || (c->opt_h_early < 0 && zk->sharded);     // option "h_early"
    B200_TRY(h_on_device(c, zk, true, (lenA && !h_early) ? c->ev_sort[0] : c->ev_h, poly_mask, combine));
    // The three witness G1 MSMs read one sorted entry list: with resident tables their accumulations go into ONE
    // launch (msm.cuh k_msm_accumulate_sets).  A single GPU that also combines here (no exchange in between) adds the
    // H MSM to the same launch in prove_enqueue_h: the H pipeline finishes under the G2 accumulation, so nothing waits.
    // Measured on B200 (profiles/r02_fuse_ab.md): on one GPU the three witness MSMs in one launch and H in a second one
    // is best (the witness MSMs' bucket reductions hide under the H accumulation; all four at once leaves four
    // reductions for the tail); on a shard the reductions are latency-bound and equally long alone or together, so all
    // four accumulations share one launch.  Option "fuse_g1": -1 auto, 0 one launch per MSM (round 1), 3 witness MSMs
    // only, 4 all four.
    const bool fuse = c->opt_fuse_g1 != 0 && same_geom && zk->tA.tbl;
    const bool with_h = fuse && (c->opt_fuse_g1 == 4 || (c->opt_fuse_g1 != 3 && zk->sharded));
    // with H in the same launch nothing follows the four accumulations: every bucket reduction is a "tail" one
    B200_TRY(msm_g1_enqueue(c, zk->d_A, w + zk->rA.lo * 32, 32, lenA, 2, &zk->tA, same_geom, with_h, 0, nullptr, fuse));
    B200_TRY(msm_g1_enqueue(c, zk->d_B1, w + zk->rA.lo * 32, 32, lenA, 3, &zk->tB1, same_geom, with_h, 0, nullptr, fuse));
    B200_TRY(msm_g1_enqueue(c, zk->d_C, w + zk->rA.lo * 32, 32, lenA, 4, &zk->tC, same_geom, with_h, 0, nullptr, fuse));
    if (fuse && !with_h) B200_TRY(msm_g1_flush(c));   // else: launched together with H by prove_enqueue_h
    zk->stage1_done = true;
    zk->stage1_combined = combine;
    return B200_OK;
}

static int prove_stage2_impl(Ctx *c, b200_zkey *zk, void *out768);
static int prove_stage2(Ctx *c, b200_zkey *zk, void *out768) {
    if (!zk->stage1_done) { c->err = "prove_finish without prove_begin"; return B200_ERR_ARG; }
    zk->stage1_done = false;
    int rc = prove_stage2_impl(c, zk, out768);
    if (rc != B200_OK) prove_drain(c);
    return rc;
}

// combine (when not done in stage 1) + the H MSM, all enqueued, nothing waited for
static int prove_enqueue_h(Ctx *c, b200_zkey *zk) {
    cudaSetDevice(c->device);
    if (!zk->stage1_combined) {     // the exchanged a, b, c -> h, on the H stream behind the caller's exchange
        cudaStream_t main_stream = c->stream;
        c->stream = c->hstream;
        phase_begin(c, PH_NTT);
        // only this shard's slice of h is ever read (its range of the H table): the exchange before this call need
        // not deliver more than that slice of a, b and c
        const uint64_t lo = zk->rH.lo, len = zk->rH.hi - zk->rH.lo;
        int rc = len ? h_combine(c, zk->d_a + lo, zk->d_b + lo, zk->d_c + lo, len) : B200_OK;
        phase_end(c);
        c->stream = main_stream;
        B200_TRY(rc);
    }
    // the digit sort of h runs on the H-pipeline stream right behind the NTTs (own sort workspace), i.e. under the
    // witness accumulations; only the H accumulation itself waits for it on the main stream
    B200_TRY(msm_g1_enqueue(c, zk->d_H, (const uint8_t *)zk->d_a + zk->rH.lo * 32, 32, zk->rH.hi - zk->rH.lo, 0, &zk->tH, false, true,
                            1, c->hstream, c->opt_fuse_g1 != 0));
    return msm_g1_flush(c);     // H alone, or H together with the witness MSMs still queued by stage 1
}

static int prove_sync_all(Ctx *c) {
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < Ctx::MSM_SLOTS; i++) if (c->side[i]) B200_CUDA_CHECK(c, cudaStreamSynchronize(c->side[i]));
    B200_CUDA_CHECK(c, cudaStreamSynchronize(c->hstream));
    phase_collect(c);
    return B200_OK;
}

static int prove_stage2_impl(Ctx *c, b200_zkey *zk, void *out768) {
    B200_TRY(prove_enqueue_h(c, zk));
    uint8_t *o = (uint8_t *)out768;
    G1Xyzz pih, pia, pib1, pic;
    G2Xyzz pib;
    B200_TRY(msm_g1_collect(c, 0, &pih));
    B200_TRY(msm_g2_collect(c, 1, &pib));
    B200_TRY(msm_g1_collect(c, 2, &pia));
    B200_TRY(msm_g1_collect(c, 3, &pib1));
    B200_TRY(msm_g1_collect(c, 4, &pic));
    B200_TRY(prove_sync_all(c));
    memcpy(o, &pih, 128);
    memcpy(o + 128, &pia, 128);
    memcpy(o + 256, &pib1, 128);
    memcpy(o + 384, &pib, 256);
    memcpy(o + 640, &pic, 128);
    return B200_OK;
}

static int prove_msms_impl(b200_ctx *h, b200_zkey *zk, const void *wtns_host, bool wtns_on_device, void *out768) {
    if (!h || !zk || !wtns_host || !out768) return B200_ERR_ARG;
    Ctx *c = &h->c;
    B200_TRY(prove_stage1(c, zk, wtns_host, wtns_on_device, 7u, true));
    return prove_stage2(c, zk, out768);
}

// The whole of Prover::prove (groth16.cpp:48-253) in one call: everything is enqueued first, then the host does the
// blinding work in the order the results become available - the key-only part (r*delta1, s*delta1, rs*delta1,
// s*delta2: most of it) right away, s*A + r*B1 as soon as pi_a and pib1 are collected (the GPU is still busy with the
// C and H accumulations then), B when the G2 reduction is through, and only C = ... + pih after the last kernel.
static int groth16_prove_impl(Ctx *c, b200_zkey *zk, const void *wtns, bool on_device, const b200_vkey *vk, const uint8_t *r32,
                              const uint8_t *s32, uint8_t *proof256, uint8_t *msms768) {
    if (zk->stage1_done) { c->err = "groth16_prove: a b200_prove_begin is pending on this zkey"; return B200_ERR_ARG; }
    B200_TRY(prove_stage1_impl(c, zk, wtns, on_device, 7u, true));
    zk->stage1_done = false;
    B200_TRY(prove_enqueue_h(c, zk));
    uint8_t prep[640], T[128];
    groth16_blind_prepare(vk->delta1, vk->delta2, r32, s32, prep);
    G1Xyzz pih, pia, pib1, pic;
    G2Xyzz pib;
    B200_TRY(msm_g1_collect(c, 2, &pia));
    B200_TRY(msm_g1_collect(c, 3, &pib1));
    groth16_blind_ab(&pia, &pib1, vk->alpha1, vk->beta1, prep, r32, s32, proof256, T);
    B200_TRY(msm_g2_collect(c, 1, &pib));
    groth16_blind_b(&pib, vk->beta2, prep, proof256 + 64);
    B200_TRY(msm_g1_collect(c, 4, &pic));
    B200_TRY(msm_g1_collect(c, 0, &pih));
    groth16_blind_c(&pic, &pih, T, prep, proof256 + 192);
    B200_TRY(prove_sync_all(c));
    if (msms768) {
        memcpy(msms768, &pih, 128); memcpy(msms768 + 128, &pia, 128); memcpy(msms768 + 256, &pib1, 128);
        memcpy(msms768 + 384, &pib, 256); memcpy(msms768 + 640, &pic, 128);
    }
    return B200_OK;
}

int b200_groth16_prove(b200_ctx *h, b200_zkey *zk, const void *wtns, int wtns_on_device, const b200_vkey *vk, const void *r32,
                       const void *s32, void *out_proof256, void *out_msms768) {
    if (!h || !zk || !wtns || !vk || !r32 || !s32 || !out_proof256) return B200_ERR_ARG;
    if (!vk->alpha1 || !vk->beta1 || !vk->beta2 || !vk->delta1 || !vk->delta2) { h->c.err = "groth16_prove: null key point"; return B200_ERR_ARG; }
    if (zk->sharded) { h->c.err = "groth16_prove: the zkey is one shard of several (use b200_prove_msms + fold + finalize)"; return B200_ERR_ARG; }
    Ctx *c = &h->c;
    int rc = groth16_prove_impl(c, zk, wtns, wtns_on_device != 0, vk, (const uint8_t *)r32, (const uint8_t *)s32,
                                (uint8_t *)out_proof256, (uint8_t *)out_msms768);
    if (rc != B200_OK) prove_drain(c);
    return rc;
}

int b200_prove_begin(b200_ctx *h, b200_zkey *zk, const void *wtns, int wtns_on_device, uint32_t poly_mask,
                     void **d_abc3, void **h_stream) {
    if (!h || !zk || !wtns || !d_abc3 || !h_stream) return B200_ERR_ARG;
    Ctx *c = &h->c;
    B200_TRY(prove_stage1(c, zk, wtns, wtns_on_device != 0, poly_mask & 7u, false));
    d_abc3[0] = zk->d_a; d_abc3[1] = zk->d_b; d_abc3[2] = zk->d_c;
    *h_stream = (void *)c->hstream;
    return B200_OK;
}

int b200_exchange_polys(b200_ctx *const *hs, b200_zkey *const *zks, int n) {
    if (!hs || !zks || n < 1) return B200_ERR_ARG;
    for (int g = 0; g < n; g++)
        if (!hs[g] || !zks[g] || !zks[g]->stage1_done || zks[g]->stage1_combined ||
            zks[g]->domain_size != zks[0]->domain_size) {
            if (hs[g]) hs[g]->c.err = "exchange_polys: every shard needs a pending b200_prove_begin of the same circuit";
            return B200_ERR_ARG;
        }
    for (int i = 0; i < 3; i++) {
        const int o = i % n;
        Ctx *co = &hs[o]->c;
        Fr *src = i == 0 ? zks[o]->d_a : i == 1 ? zks[o]->d_b : zks[o]->d_c;
        for (int r = 0; r < n; r++) {
            if (r == o) continue;
            Ctx *cr = &hs[r]->c;
            Fr *dst = i == 0 ? zks[r]->d_a : i == 1 ? zks[r]->d_b : zks[r]->d_c;
            // shard r combines (and reads) only its own range of h: that slice of the polynomial is all it needs
            const size_t lo = (size_t)zks[r]->rH.lo, bytes = (size_t)(zks[r]->rH.hi - zks[r]->rH.lo) * sizeof(Fr);
            if (bytes == 0) continue;
            cudaSetDevice(cr->device);
            if (cr->device != co->device && !(cr->peer_enabled & (1ull << co->device)) && co->device < 64) {
                cudaError_t e = cudaDeviceEnablePeerAccess(co->device, 0);   // direct NVLink path when available
                if (e != cudaSuccess) cudaGetLastError();                    // already enabled / unsupported: staged copy
                cr->peer_enabled |= 1ull << co->device;
            }
            // co->ev_h: recorded on the owner's H stream behind its transform chains (h_on_device)
            B200_CUDA_CHECK(cr, cudaStreamWaitEvent(cr->hstream, co->ev_h, 0));
            B200_CUDA_CHECK(cr, cudaMemcpyPeerAsync(dst + lo, cr->device, src + lo, co->device, bytes, cr->hstream));
        }
    }
    // Write-after-read: the combine of b200_prove_finish overwrites d_a IN PLACE on the owner's H stream, while the
    // peers' copies above read the owner's buffers on THEIR H streams.  Every receiver records "my copies are done"
    // and every owner's H stream waits for all of them before anything later (the combine) may run.
    if (n > 1) {
        for (int r = 0; r < n; r++) {
            Ctx *cr = &hs[r]->c;
            cudaSetDevice(cr->device);
            B200_CUDA_CHECK(cr, cudaEventRecord(cr->ev_xchg, cr->hstream));
        }
        for (int i = 0; i < 3 && i < n; i++) {   // owners are shards 0 .. min(3, n) - 1
            Ctx *co = &hs[i]->c;
            cudaSetDevice(co->device);
            for (int r = 0; r < n; r++)
                if (r != i) B200_CUDA_CHECK(co, cudaStreamWaitEvent(co->hstream, hs[r]->c.ev_xchg, 0));
        }
    }
    return B200_OK;
}

int b200_prove_finish(b200_ctx *h, b200_zkey *zk, void *out768) {
    if (!h || !zk || !out768) return B200_ERR_ARG;
    return prove_stage2(&h->c, zk, out768);
}

int b200_prove_msms(b200_ctx *h, b200_zkey *zk, const void *wtns_host, void *out768) {
    return prove_msms_impl(h, zk, wtns_host, false, out768);
}
int b200_prove_msms_dev(b200_ctx *h, b200_zkey *zk, const void *d_wtns, void *out768) {
    return prove_msms_impl(h, zk, d_wtns, true, out768);
}

int b200_fixed_base_g1(b200_ctx *h, const void *base_affine64, const void *scalars32, uint64_t n, void *out_affine) {
    if (!h) return B200_ERR_ARG;
    cudaSetDevice(h->c.device);
    return fixed_base_run<Fq>(&h->c, base_affine64, scalars32, n, out_affine);
}
int b200_fixed_base_g2(b200_ctx *h, const void *base_affine128, const void *scalars32, uint64_t n, void *out_affine) {
    if (!h) return B200_ERR_ARG;
    cudaSetDevice(h->c.device);
    return fixed_base_run<Fq2>(&h->c, base_affine128, scalars32, n, out_affine);
}

}  // extern "C"
