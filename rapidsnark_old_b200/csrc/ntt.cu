// Fr-domain number-theoretic transforms and the H-polynomial pipeline on sm_100a.
//
// Replaces FFT<Fr>::fft / ifft (depends/ffiasm/c/fft.cpp:175-212; roots built by the ctor :32-115:
// w_{2^s} = 5^((r-1)/2^s)).  The reference runs log2(n) full-array radix-2 passes after a bit-reverse
// permutation (fft.cpp:159-171).  Here a transform of 2^k points is 2 (k <= 20) or 3 passes over HBM:
// each pass keeps a tile of 2^S rows x 2^q adjacent columns in shared memory (two 16-byte planes per
// element so 128-bit shared loads are conflict-free), runs S radix-2 butterfly stages on it with a
// shared-memory copy of the 2^(S-1) local twiddles, and applies the inter-pass twiddle
// w_M^(low * bitrev(row)) on the way in (DIT) or out (DIF).  Inter-pass twiddles and the coset twist
// come from a two-level table of w_{2n} (2 x 2^14 entries instead of the reference's 2n-entry table,
// fft.cpp:74-102).
//   DIF passes: natural order in  -> bit-reversed out   (inverse transform of the H pipeline)
//   DIT passes: bit-reversed in   -> natural out        (forward transform of the H pipeline)
// so prove()'s ifft -> coset twist -> fft (src/groth16.cpp:101-155) needs no permutation pass at all;
// the stand-alone b200_ntt_fr (natural in/out like the reference) adds one bit-reverse kernel.
#include <cuda.h>
#include <map>
#include <tuple>
#include "ctx.cuh"
#include "memops.cuh"

namespace b200 {

static const int NTT_LB = 14;        // low-table bits of the two-level root table
static const int NTT_S_LAST = 11;    // max stages of the contiguous (lo = 0) pass: 2^11 x 32 B = 64 KB tile
static const int NTT_S_STRIDED = 9;  // max stages of a strided pass (x 4 columns = 64 KB tile)
static const int NTT_Q = 2;          // log2 columns of a strided tile: 4 x 32 B = one 128-byte line

struct NttTable {
    int s = 0;             // tables are powers of W = w_{2^s}
    int loc_log = 0;       // loc[i] = w_{2^loc_log}^i, i < 2^(loc_log-1)
    Fr *t_lo = nullptr;    // W^i, i < 2^LB
    Fr *t_hi = nullptr;    // W^(i << LB)
    Fr *loc = nullptr;
};

struct Ctx::Twiddles {
    // index [s][inverse]
    NttTable tab[29][2];
    // twist_br[k][p] = n^-1 * w_{2n}^bitrev_k(p), n = 2^k: the factor the fused last inverse pass applies at position p
    Fr *twist_br[29] = {nullptr};
    // tensor maps of the TMA passes, per (array, log2 size, lo)
    std::map<std::tuple<const void *, int, int>, CUtensorMap> tmaps;
};

DEVFN Fr lds_fr(const uint4 *p0, const uint4 *p1, u32 i) {
    Fr r;
    uint4 a = p0[i], b = p1[i];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
DEVFN void sts_fr(uint4 *p0, uint4 *p1, u32 i, const Fr &r) {
    p0[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p1[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// W^e from the two-level table
DEVFN Fr root_pow(const Fr *__restrict__ t_lo, const Fr *__restrict__ t_hi, int s, u32 e) {
    Fr lo = ldg_struct(t_lo + (e & ((1u << NTT_LB) - 1)));
    if (s <= NTT_LB) return lo;
    Fr hi = ldg_struct(t_hi + (e >> NTT_LB));
    return fp_mul(lo, hi);
}

// build tables: out[i] = base^(i * stride_mul) for i < count, base given in Montgomery form
__global__ void __launch_bounds__(256) k_ntt_build_powers(Fr base, u32 count, Fr *__restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr r = Fr::one(), b = base;
    for (u32 e = i; e != 0; e >>= 1) {
        if (e & 1) r = fp_mul(r, b);
        b = fp_sqr(b);
    }
    st_struct(out + i, r);
}

// G butterfly stages on 2^G elements held in registers (radix-2^G step): one shared-memory round trip and one
// barrier per G stages instead of per stage.
//   DIF: stages with half = H, H/2, .., H/2^(G-1); the thread owns rows  blk*2H + j + m*(2H >> G), m < 2^G
//   DIT: stages with half = h0, 2*h0, .., h0*2^(G-1); the thread owns rows  blk*(h0 << G) + j + m*h0
// Twiddle of a butterfly whose lower row is r at a stage of half-size `half`: w_R^((r mod half) * (R/2) / half).
// The tile in shared memory, two layouts:
//   TilePlanes  element e = 16-byte halves at x0[e] and x1[e] (two planes: 128-bit accesses of a warp are conflict-free)
//   TileSwz     what a TMA load with CU_TENSOR_MAP_SWIZZLE_128B leaves: rows of 128 bytes = 4 elements, the 16-byte
//               chunk c of row r stored at chunk c ^ (r & 7); lanes on the same column of 8 consecutive rows, or on
//               the 4 columns of 2 rows, hit 8 different chunks: conflict-free as well
struct TilePlanes {
    uint4 *x0, *x1;
    DEVFN Fr ld(u32 e) const { return lds_fr(x0, x1, e); }
    DEVFN void st(u32 e, const Fr &v) const { sts_fr(x0, x1, e, v); }
};
struct TileSwz {
    uint4 *t;   // 1024-byte aligned
    DEVFN u32 at(u32 e, u32 half) const { const u32 row = e >> 2; return row * 8 + ((((e & 3) << 1) | half) ^ (row & 7)); }
    DEVFN Fr ld(u32 e) const {
        Fr r;
        uint4 a = t[at(e, 0)], b = t[at(e, 1)];
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    DEVFN void st(u32 e, const Fr &r) const {
        t[at(e, 0)] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        t[at(e, 1)] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    }
};

template <bool DIT, int G, class Tile>
DEVFN void ntt_group(const Tile &X, const uint4 *w0, const uint4 *w1, u32 R, int q, u32 hsel, u32 tid, u32 nthreads) {
    const u32 Q = 1u << q;
    const u32 stride = DIT ? hsel : ((2 * hsel) >> G);      // row distance between consecutive m
    const u32 span = stride << G;                            // rows covered by one group instance
    const u32 units = (R >> G) << q;                         // (group instance, j, column) triples in the tile
    for (u32 id = tid; id < units; id += nthreads) {
        const u32 c = id & (Q - 1), t = id >> q;
        const u32 j = t % stride, blk = t / stride;
        const u32 r0 = blk * span + j;
        Fr v[1 << G];
#pragma unroll
        for (int m = 0; m < (1 << G); m++) v[m] = X.ld(((r0 + m * stride) << q) + c);
#pragma unroll
        for (int s = 0; s < G; s++) {
            const int bit = DIT ? s : (G - 1 - s);           // partner distance in m is 2^bit
            const u32 half = stride << bit;
            const u32 tw_mul = (R >> 1) / half;
#pragma unroll
            for (int m = 0; m < (1 << G); m++) {
                if (m & (1 << bit)) continue;
                const int m2 = m | (1 << bit);
                const u32 pos = j + (u32)(m & ((1 << bit) - 1)) * stride;   // (row of m) mod half
                if (DIT) {
                    Fr tv = v[m2];
                    if (pos) tv = fp_mul(tv, lds_fr(w0, w1, pos * tw_mul));
                    Fr u = v[m];
                    v[m] = fp_add(u, tv);
                    v[m2] = fp_sub(u, tv);
                } else {
                    Fr u = v[m], w = v[m2];
                    v[m] = fp_add(u, w);
                    Fr d = fp_sub(u, w);
                    if (pos) d = fp_mul(d, lds_fr(w0, w1, pos * tw_mul));
                    v[m2] = d;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < (1 << G); m++) X.st(((r0 + m * stride) << q) + c, v[m]);
    }
}

// all S butterfly stages of a tile, radix-8 steps first; every step ends with a barrier
template <bool DIT, class Tile>
DEVFN void ntt_stages(const Tile &X, const uint4 *w0, const uint4 *w1, u32 R, int S, int q) {
    if (DIT) {
        int done = 0;
        while (S - done >= 3) { ntt_group<true, 3>(X, w0, w1, R, q, 1u << done, threadIdx.x, blockDim.x); done += 3; __syncthreads(); }
        if (S - done == 2) { ntt_group<true, 2>(X, w0, w1, R, q, 1u << done, threadIdx.x, blockDim.x); done += 2; __syncthreads(); }
        if (S - done == 1) { ntt_group<true, 1>(X, w0, w1, R, q, 1u << done, threadIdx.x, blockDim.x); done += 1; __syncthreads(); }
    } else {
        int left = S;   // stages left; the next stage has half = 2^(left-1)
        while (left >= 3) { ntt_group<false, 3>(X, w0, w1, R, q, 1u << (left - 1), threadIdx.x, blockDim.x); left -= 3; __syncthreads(); }
        if (left == 2) { ntt_group<false, 2>(X, w0, w1, R, q, 2u, threadIdx.x, blockDim.x); left -= 2; __syncthreads(); }
        if (left == 1) { ntt_group<false, 1>(X, w0, w1, R, q, 1u, threadIdx.x, blockDim.x); left -= 1; __syncthreads(); }
    }
}

// One pass over HBM: tile of 2^S rows x 2^q columns in shared memory, S stages in radix-8 (then radix-4/2) steps.
// FUSE (DIF, lo == 0 only): the ifft tail and the coset twist of groth16.cpp:107-110 are applied on the way out:
// position p holds coefficient bitrev(p), multiplied by n^-1 * w_2n^bitrev(p).
// twist_br[p] = n_inv * W^bitrev(p) (two-level table product), one entry per position, in the order the fused pass
// reads it (coalesced): one product per element in that pass instead of three
__global__ void __launch_bounds__(256) k_ntt_build_twist(Fr *__restrict__ out, int k, Fr n_inv, NttTable tb) {
    u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >> k) return;
    u32 i = k ? (u32)(__brevll(p) >> (64 - k)) : 0;
    u32 ex = i << (tb.s - (k + 1));
    Fr x = n_inv;
    if (ex) x = fp_mul(x, root_pow(tb.t_lo, tb.t_hi, tb.s, ex));
    st_struct(out + p, x);
}

template <bool DIT, bool FUSE>
__global__ void __launch_bounds__(256, 2) k_ntt_pass(Fr *__restrict__ a, int lo, int S, int q, NttTable tb, int k,
                                                      const Fr *__restrict__ twist_br) {
    extern __shared__ uint4 smem_raw[];
    const u32 R = 1u << S, Q = 1u << q, tile_elems = R << q;
    uint4 *x0 = smem_raw, *x1 = x0 + tile_elems;
    uint4 *w0 = x1 + tile_elems, *w1 = w0 + (R >> 1);

    const u32 tiles_per_group = 1u << (lo - q);
    const u32 tile = blockIdx.x;
    const u64 base = (u64)(tile >> (lo - q)) << (lo + S);
    const u32 l0 = (tile & (tiles_per_group - 1)) << q;
    const int tw_shift = tb.s - (lo + S);   // w_M = W^(2^tw_shift)

    for (u32 i = threadIdx.x; i < (R >> 1); i += blockDim.x) {
        Fr w = ldg_struct(tb.loc + ((size_t)i << (tb.loc_log - S)));
        sts_fr(w0, w1, i, w);
    }
    for (u32 e = threadIdx.x; e < tile_elems; e += blockDim.x) {
        u32 r = e >> q, c = e & (Q - 1);
        Fr x = ld_struct(a + base + ((u64)r << lo) + l0 + c);
        if (DIT && lo > 0) {
            u32 k1 = __brev(r) >> (32 - S);
            u32 ex = ((l0 + c) * k1) << tw_shift;
            if (ex) x = fp_mul(x, root_pow(tb.t_lo, tb.t_hi, tb.s, ex));
        }
        sts_fr(x0, x1, e, x);
    }
    __syncthreads();

    ntt_stages<DIT>(TilePlanes{x0, x1}, w0, w1, R, S, q);

    for (u32 e = threadIdx.x; e < tile_elems; e += blockDim.x) {
        u32 r = e >> q, c = e & (Q - 1);
        Fr x = lds_fr(x0, x1, e);
        if (!DIT && lo > 0) {
            u32 k1 = __brev(r) >> (32 - S);
            u32 ex = ((l0 + c) * k1) << tw_shift;
            if (ex) x = fp_mul(x, root_pow(tb.t_lo, tb.t_hi, tb.s, ex));
        }
        if (FUSE) x = fp_mul(x, ldg_struct(twist_br + base + r));   // lo == 0, q == 0: position p = base + r
        st_struct(a + base + ((u64)r << lo) + l0 + c, x);
    }
}


// ---- the same pass with the tile moved by the TMA engine ---------------------------------------------------------------
// A strided tile (2^S rows of 4 adjacent elements, 2^lo elements apart) is a box of the array seen as a matrix with
// rows of 2^lo elements; the contiguous tile of the lo == 0 pass is a box of the array seen as rows of 4 elements.
// Either way: inner extent 128 bytes, up to 256 rows per cp.async.bulk.tensor, 128-byte swizzle.  One thread arms an
// mbarrier with the tile's byte count and issues the loads; the others fetch the local twiddles meanwhile; after the
// butterflies the tile goes back with cp.async.bulk.tensor stores.  The inter-pass twiddles, which the plain kernel
// applies while copying, are applied in shared memory here.
DEVFN u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
DEVFN void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
DEVFN void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEVFN void mbar_wait(u64 *bar, u32 phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NTT_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NTT_TMA_DONE;\n"
        "bra NTT_TMA_WAIT;\n"
        "NTT_TMA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
DEVFN void tma_load_2d(void *dst, const CUtensorMap *map, u64 *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
DEVFN void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <bool DIT, bool FUSE>
__global__ void __launch_bounds__(256, 2) k_ntt_pass_tma(const __grid_constant__ CUtensorMap map, int lo, int S, int q,
                                                          NttTable tb, int k, const Fr *__restrict__ twist_br) {
    extern __shared__ uint4 smem_raw[];
    const u32 R = 1u << S, tile_elems = R << q, nrows = tile_elems >> 2;      // smem rows of 128 bytes
    uint4 *t = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint4 *w0 = t + 2 * tile_elems, *w1 = w0 + (R >> 1);
    u64 *bar = reinterpret_cast<u64 *>(w1 + (R >> 1));
    const TileSwz X{t};

    const u32 tiles_per_group = 1u << (lo - q);
    const u32 tile = blockIdx.x;
    const u64 base = (u64)(tile >> (lo - q)) << (lo + S);
    const u32 l0 = (tile & (tiles_per_group - 1)) << q;
    const int tw_shift = tb.s - (lo + S);
    const u32 bh = nrows < 256 ? nrows : 256, nbox = nrows / bh;
    // box coordinates: (u32 column, matrix row)
    const int c0 = lo ? (int)(l0 * 8) : 0;
    const int row0 = lo ? (int)(base >> lo) : (int)(base >> 2);

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, tile_elems * 32);
        for (u32 j = 0; j < nbox; j++) tma_load_2d(t + (size_t)j * bh * 8, &map, bar, c0, row0 + (int)(j * bh));
    }
    for (u32 i = threadIdx.x; i < (R >> 1); i += blockDim.x) {
        Fr w = ldg_struct(tb.loc + ((size_t)i << (tb.loc_log - S)));
        sts_fr(w0, w1, i, w);
    }
    mbar_wait(bar, 0);
    if (DIT && lo > 0) {
        for (u32 e = threadIdx.x; e < tile_elems; e += blockDim.x) {
            u32 r = e >> q, c = e & ((1u << q) - 1);
            u32 k1 = __brev(r) >> (32 - S);
            u32 ex = ((l0 + c) * k1) << tw_shift;
            if (ex) X.st(e, fp_mul(X.ld(e), root_pow(tb.t_lo, tb.t_hi, tb.s, ex)));
        }
    }
    __syncthreads();

    ntt_stages<DIT>(X, w0, w1, R, S, q);

    if ((!DIT && lo > 0) || FUSE) {
        for (u32 e = threadIdx.x; e < tile_elems; e += blockDim.x) {
            u32 r = e >> q, c = e & ((1u << q) - 1);
            Fr x = X.ld(e);
            if (!DIT && lo > 0) {
                u32 k1 = __brev(r) >> (32 - S);
                u32 ex = ((l0 + c) * k1) << tw_shift;
                if (ex) x = fp_mul(x, root_pow(tb.t_lo, tb.t_hi, tb.s, ex));
            }
            if (FUSE) x = fp_mul(x, ldg_struct(twist_br + base + e));      // lo == 0: position p = base + e
            X.st(e, x);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
        for (u32 j = 0; j < nbox; j++) tma_store_2d(&map, t + (size_t)j * bh * 8, c0, row0 + (int)(j * bh));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) k_ntt_bitrev(Fr *__restrict__ a, int k) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> k) return;
    u64 j = (u64)(__brevll(i) >> (64 - k));
    if (i < j) {
        Fr x = ld_struct(a + i), y = ld_struct(a + j);
        st_struct(a + i, y);
        st_struct(a + j, x);
    }
}

// a[i] *= f   (ifft tail: x powTwoInv, fft.cpp:206-211)
__global__ void __launch_bounds__(256) k_ntt_scale(Fr *__restrict__ a, u64 n, Fr f) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_struct(a + i, fp_mul(ld_struct(a + i), f));
}

// after the inverse DIF passes position p holds coefficient bitrev(p): scale by 1/n and apply the coset
// twist w_{2n}^bitrev(p)  (groth16.cpp:107-110 with fft.hpp:28 root(domainPower+1, i))
__global__ void __launch_bounds__(256) k_ntt_scale_twist_brev(Fr *__restrict__ a, int k, Fr n_inv, NttTable tb) {
    u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >> k) return;
    u32 i = k ? (u32)(__brevll(p) >> (64 - k)) : 0;
    u32 ex = i << (tb.s - (k + 1));
    Fr x = fp_mul(ld_struct(a + p), n_inv);
    if (ex) x = fp_mul(x, root_pow(tb.t_lo, tb.t_hi, tb.s, ex));
    st_struct(a + p, x);
}

// h[i] = fromMontgomery(a[i]*b[i] - c[i])   (groth16.cpp:158-163), written over a
__global__ void __launch_bounds__(256) k_h_combine(Fr *__restrict__ a, const Fr *__restrict__ b,
                                                   const Fr *__restrict__ c, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr t = fp_sub(fp_mul(ld_struct(a + i), ld_struct(b + i)), ld_struct(c + i));
    st_struct(a + i, fp_from_mont(t));
}

// a[row], b[row] from the CSR-repacked coefficient records, c = a*b   (groth16.cpp:62-96).
// wtns is in normal form and coef = value*R^2, so one Montgomery product gives Montgomery form.
__global__ void __launch_bounds__(128) k_build_abc(const Fr *__restrict__ wtns, const u32 *__restrict__ row_ptr_a,
                                                   const u32 *__restrict__ row_ptr_b, const u32 *__restrict__ sig,
                                                   const Fr *__restrict__ coef, u32 n, Fr *__restrict__ a,
                                                   Fr *__restrict__ b, Fr *__restrict__ c) {
    u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    Fr sa = Fr::zero(), sb = Fr::zero();
    for (u32 k = row_ptr_a[row], e = row_ptr_a[row + 1]; k < e; k++)
        sa = fp_add(sa, fp_mul(ldg_struct(wtns + sig[k]), ldg_struct(coef + k)));
    for (u32 k = row_ptr_b[row], e = row_ptr_b[row + 1]; k < e; k++)
        sb = fp_add(sb, fp_mul(ldg_struct(wtns + sig[k]), ldg_struct(coef + k)));
    st_struct(a + row, sa);
    st_struct(b + row, sb);
    st_struct(c + row, fp_mul(sa, sb));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static Fr host_root_of_unity(int s, bool inverse) {
    // 5^((r-1)/2^s) (fft.cpp:52-83: nqr = 5 for BN254 r, 2-adicity 28)
    u32 e[8];
    for (int i = 0; i < 8; i++) e[i] = FrParams::mod(i);
    e[0] -= 1;
    // e >>= s
    for (int k = 0; k < s; k++) {
        for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? (e[i + 1] << 31) : 0);
    }
    Fr five = Fr::zero();
    five.v[0] = 5;
    five = fp_to_mont(five);
    Fr w = fp_pow(five, e);
    return inverse ? fp_inv(w) : w;
}

static Fr host_pow2k(Fr w, int k) {  // w^(2^k)
    for (int i = 0; i < k; i++) w = fp_sqr(w);
    return w;
}

void ntt_free_tables(Ctx *ctx) {
    if (!ctx->tw) return;
    for (int s = 0; s < 29; s++)
        for (int inv = 0; inv < 2; inv++) {
            NttTable &t = ctx->tw->tab[s][inv];
            if (t.t_lo) cudaFree(t.t_lo);
            if (t.t_hi) cudaFree(t.t_hi);
            if (t.loc) cudaFree(t.loc);
        }
    for (int k = 0; k < 29; k++) if (ctx->tw->twist_br[k]) cudaFree(ctx->tw->twist_br[k]);
    delete ctx->tw;
    ctx->tw = nullptr;
}

static int ntt_get_table(Ctx *ctx, int k, bool inverse, NttTable *out) {
    const int s = k + 1;   // tables of w_{2n}: also serve the coset twist
    if (s > 28) { ctx->err = "Domain size too big for the curve"; return B200_ERR_RANGE; }
    if (!ctx->tw) ctx->tw = new Ctx::Twiddles();
    NttTable &t = ctx->tw->tab[s][inverse ? 1 : 0];
    if (!t.t_lo) {
        t.s = s;
        t.loc_log = k < NTT_S_LAST ? (k < 1 ? 1 : k) : NTT_S_LAST;
        Fr W = host_root_of_unity(s, inverse);
        const u32 n_lo = 1u << NTT_LB, n_hi = s > NTT_LB ? (1u << (s - NTT_LB)) : 1u, n_loc = 1u << (t.loc_log - 1);
        B200_CUDA_CHECK(ctx, cudaMalloc(&t.t_lo, (size_t)n_lo * sizeof(Fr)));
        B200_CUDA_CHECK(ctx, cudaMalloc(&t.t_hi, (size_t)n_hi * sizeof(Fr)));
        B200_CUDA_CHECK(ctx, cudaMalloc(&t.loc, (size_t)n_loc * sizeof(Fr)));
        B200_LAUNCH(ctx, k_ntt_build_powers, (n_lo + 255) / 256, 256, 0, W, n_lo, t.t_lo);
        B200_LAUNCH(ctx, k_ntt_build_powers, (n_hi + 255) / 256, 256, 0, host_pow2k(W, NTT_LB), n_hi, t.t_hi);
        B200_LAUNCH(ctx, k_ntt_build_powers, (n_loc + 255) / 256, 256, 0, host_pow2k(W, s - t.loc_log), n_loc, t.loc);
    }
    *out = t;
    return B200_OK;
}

struct NttPass { int lo, S, q; };

static int ntt_plan(int k, NttPass *p) {   // passes ordered from the high bits down (DIF order)
    int np = 0;
    int last = k < NTT_S_LAST ? k : NTT_S_LAST;
    int rest = k - last;
    int nstr = (rest + NTT_S_STRIDED - 1) / NTT_S_STRIDED;
    int hi = k;
    for (int i = 0; i < nstr; i++) {
        int S = (rest + (nstr - i) - 1) / (nstr - i);
        rest -= S;
        p[np].lo = hi - S;
        p[np].S = S;
        p[np].q = p[np].lo < NTT_Q ? p[np].lo : NTT_Q;
        hi -= S;
        np++;
    }
    p[np].lo = 0; p[np].S = last; p[np].q = 0;
    np++;
    return np;
}

// 2-D view of the array for the TMA passes: rows of 2^lo elements (strided passes) or of 4 elements (lo == 0); the box
// is 128 bytes x up to 256 rows, 128-byte swizzle.  cuTensorMapEncodeTiled comes from the driver through the runtime's
// entry-point query, so the library does not link libcuda.
static int ntt_tensor_map(Ctx *ctx, Fr *d_a, int k, int lo, u32 box_rows, CUtensorMap *out) {
    auto key = std::make_tuple((const void *)d_a, k, lo * 1024 + (int)box_rows);
    auto it = ctx->tw->tmaps.find(key);
    if (it != ctx->tw->tmaps.end()) { *out = it->second; return B200_OK; }
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
            cudaGetLastError();
            ctx->err = "ntt: cuTensorMapEncodeTiled is not available";
            return B200_ERR_CUDA;
        }
        encode = (EncodeFn)fn;
    }
    const int rl = lo ? lo : 2;                                   // log2 elements per matrix row
    cuuint64_t dims[2] = {(cuuint64_t)8 << rl, (cuuint64_t)1 << (k - rl)};
    cuuint64_t strides[1] = {(cuuint64_t)32 << rl};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "ntt: cuTensorMapEncodeTiled failed"; return B200_ERR_CUDA; }
    if (ctx->tw->tmaps.size() > 64) ctx->tw->tmaps.clear();
    ctx->tw->tmaps[key] = m;
    *out = m;
    return B200_OK;
}

template <bool DIT, bool FUSE>
static int ntt_launch_pass(Ctx *ctx, Fr *d_a, int k, const NttPass &ps, const NttTable &tb, const Fr *twist_br) {
    if (ps.S == 0) return B200_OK;
    // TMA variant (option "ntt_tma"): tiles whose shared-memory rows are whole 128-byte lines
    if (ctx->opt_ntt_tma > 0 && k >= 12 && ((ps.lo == 0 && ps.S >= 5) || (ps.lo >= 2 && ps.q == 2))) {
        const u32 tile_elems = 1u << (ps.S + ps.q), nrows = tile_elems >> 2, bh = nrows < 256 ? nrows : 256;
        CUtensorMap map;
        B200_TRY(ntt_tensor_map(ctx, d_a, k, ps.lo, bh, &map));
        const size_t smem = (size_t)tile_elems * 32 + (size_t)(1u << (ps.S - 1)) * 32 + 16 + 1024;
        const u32 grid = (u32)(((u64)1 << k) >> (ps.S + ps.q));
        if (!ctx->attr_ntt_tma) {
            B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            ctx->attr_ntt_tma = true;
        }
        B200_LAUNCH(ctx, (k_ntt_pass_tma<DIT, FUSE>), grid, 256, smem, map, ps.lo, ps.S, ps.q, tb, k, twist_br);
        return B200_OK;
    }
    const u32 tile_elems = 1u << (ps.S + ps.q);
    const size_t smem = (size_t)tile_elems * 32 + (size_t)(1u << (ps.S - 1)) * 32;
    const u32 grid = (u32)(((u64)1 << k) >> (ps.S + ps.q));
    u32 threads = tile_elems / 8;
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    auto kern = k_ntt_pass<DIT, FUSE>;
    if (!ctx->attr_ntt) {   // a function attribute belongs to the current device, not to the process
        B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_ntt_pass<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        ctx->attr_ntt = true;
    }
    B200_LAUNCH(ctx, kern, grid, threads, smem, d_a, ps.lo, ps.S, ps.q, tb, k, twist_br);
    return B200_OK;
}

static int ilog2_exact(uint64_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int k = 0;
    while ((1ull << k) < n) k++;
    return k;
}

// natural -> bit-reversed (DIF).  fuse_scale_twist: multiply the result by n^-1 * w_2n^(natural index) on the way
// out of the last pass (the forward table `tw` supplies the coset roots).
int ntt_dif(Ctx *ctx, Fr *d_a, int k, bool inverse_roots, const Fr *n_inv = nullptr) {
    if (k == 0) return B200_OK;
    NttTable tb;
    B200_TRY(ntt_get_table(ctx, k, inverse_roots, &tb));
    NttPass p[8];
    int np = ntt_plan(k, p);
    const Fr *twist = nullptr;
    if (n_inv) {   // table of n_inv * (FORWARD root)^bitrev(p), built once per size
        Fr *&t = ctx->tw->twist_br[k];
        if (!t) {
            NttTable tf;
            B200_TRY(ntt_get_table(ctx, k, false, &tf));
            B200_CUDA_CHECK(ctx, cudaMalloc(&t, ((size_t)1 << k) * sizeof(Fr)));
            B200_LAUNCH(ctx, k_ntt_build_twist, (u32)((((u64)1 << k) + 255) / 256), 256, 0, t, k, *n_inv, tf);
        }
        twist = t;
    }
    for (int i = 0; i < np; i++) {
        if (n_inv && i == np - 1) B200_TRY((ntt_launch_pass<false, true>(ctx, d_a, k, p[i], tb, twist)));
        else B200_TRY((ntt_launch_pass<false, false>(ctx, d_a, k, p[i], tb, nullptr)));
    }
    return B200_OK;
}

// bit-reversed -> natural (DIT)
int ntt_dit(Ctx *ctx, Fr *d_a, int k, bool inverse_roots) {
    if (k == 0) return B200_OK;
    NttTable tb;
    B200_TRY(ntt_get_table(ctx, k, inverse_roots, &tb));
    NttPass p[8];
    int np = ntt_plan(k, p);
    for (int i = np - 1; i >= 0; i--) B200_TRY((ntt_launch_pass<true, false>(ctx, d_a, k, p[i], tb, nullptr)));
    return B200_OK;
}

static Fr host_n_inv(int k) {
    Fr two = Fr::zero();
    two.v[0] = 2;
    two = fp_to_mont(two);
    Fr inv2 = fp_inv(two), r = Fr::one();
    for (int i = 0; i < k; i++) r = fp_mul(r, inv2);
    return r;
}

// reference-order transform: natural in, natural out (fft.cpp:175-212)
int ntt_natural(Ctx *ctx, Fr *d_a, uint64_t n, bool inverse) {
    int k = ilog2_exact(n);
    if (k < 0) { ctx->err = "ntt: n must be a power of two"; return B200_ERR_ARG; }
    if (k + 1 > 28) { ctx->err = "Domain size too big for the curve"; return B200_ERR_RANGE; }
    if (k == 0) return B200_OK;
    B200_LAUNCH(ctx, k_ntt_bitrev, (u32)((n + 255) / 256), 256, 0, d_a, k);
    B200_TRY(ntt_dit(ctx, d_a, k, inverse));
    if (inverse) B200_LAUNCH(ctx, k_ntt_scale, (u32)((n + 255) / 256), 256, 0, d_a, n, host_n_inv(k));
    return B200_OK;
}

static int h_scale_twist_k0(Ctx *ctx, Fr *d, const Fr &ninv, const NttTable &tf) {
    B200_LAUNCH(ctx, k_ntt_scale_twist_brev, 1, 256, 0, d, 0, ninv, tf);
    return B200_OK;
}

// a, b, c (natural order, Montgomery) -> h scalars in d_a (normal form)   (groth16.cpp:101-163)
// nstreams = 3: the three transform chains are independent until the final combine; b and c run on their own
// streams beside a's (forked from / joined into ctx->stream by events).  Same kernels, same results; it only
// matters when the GPU is not already full (sharded zkeys: small MSMs, the H pipeline is the critical path).
// poly_mask: bit i set = run the transform chain of arr[i] here (multi-GPU: the chains are spread over the ranks
// and exchanged before the combine); combine = false leaves a, b, c as coset evaluations for that exchange.
int h_transforms(Ctx *ctx, Fr *d_a, Fr *d_b, Fr *d_c, uint64_t n, int nstreams, unsigned poly_mask) {
    int k = ilog2_exact(n);
    if (k < 0) { ctx->err = "h pipeline: domain size must be a power of two"; return B200_ERR_ARG; }
    if (k + 1 > 28) { ctx->err = "Domain size too big for the curve"; return B200_ERR_RANGE; }
    NttTable tf, ti;
    B200_TRY(ntt_get_table(ctx, k, false, &tf));
    if (k > 0) B200_TRY(ntt_get_table(ctx, k, true, &ti));   // built on ctx->stream before any fork
    Fr ninv = host_n_inv(k);
    Fr *arr[3] = {d_a, d_b, d_c};
    cudaStream_t main_stream = ctx->stream;
    if (nstreams >= 3)
        for (int i = 0; i < 2; i++)
            if (!ctx->hstream_bc[i]) cudaStreamCreateWithPriority(&ctx->hstream_bc[i], cudaStreamNonBlocking, ctx->hi_prio);
    const bool fork = nstreams >= 3 && ctx->hstream_bc[0] && ctx->hstream_bc[1];
    if (fork) B200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_h_fork, main_stream));
    int rc = B200_OK;
    for (int i = 0; i < 3 && rc == B200_OK; i++) {
        if (!((poly_mask >> i) & 1u)) continue;
        if (fork && i > 0) {
            ctx->stream = ctx->hstream_bc[i - 1];
            if (cudaStreamWaitEvent(ctx->stream, ctx->ev_h_fork, 0) != cudaSuccess) rc = B200_ERR_CUDA;
        }
        if (rc == B200_OK) {
            if (k == 0) rc = h_scale_twist_k0(ctx, arr[i], ninv, tf);
            else rc = ntt_dif(ctx, arr[i], k, true, &ninv);    // ifft + 1/n + coset twist, output bit-reversed
        }
        if (rc == B200_OK) rc = ntt_dit(ctx, arr[i], k, false);
        if (fork && i > 0) {
            if (rc == B200_OK && cudaEventRecord(ctx->ev_h_join[i - 1], ctx->stream) != cudaSuccess) rc = B200_ERR_CUDA;
            ctx->stream = main_stream;
            if (rc == B200_OK && cudaStreamWaitEvent(main_stream, ctx->ev_h_join[i - 1], 0) != cudaSuccess) rc = B200_ERR_CUDA;
        }
    }
    ctx->stream = main_stream;
    return rc;
}

// h[i] = fromMontgomery(a[i] * b[i] - c[i]) into d_a   (groth16.cpp:158-163)
int h_combine(Ctx *ctx, Fr *d_a, const Fr *d_b, const Fr *d_c, uint64_t n) {
    B200_LAUNCH(ctx, k_h_combine, (u32)((n + 255) / 256), 256, 0, d_a, d_b, d_c, n);
    return B200_OK;
}

int h_pipeline(Ctx *ctx, Fr *d_a, Fr *d_b, Fr *d_c, uint64_t n, int nstreams) {
    B200_TRY(h_transforms(ctx, d_a, d_b, d_c, n, nstreams, 7u));
    return h_combine(ctx, d_a, d_b, d_c, n);
}

int build_abc(Ctx *ctx, const Fr *d_wtns, const u32 *d_row_a, const u32 *d_row_b, const u32 *d_sig, const Fr *d_coef,
              u32 n, Fr *d_a, Fr *d_b, Fr *d_c) {
    B200_LAUNCH(ctx, k_build_abc, (n + 127) / 128, 128, 0, d_wtns, d_row_a, d_row_b, d_sig, d_coef, n, d_a, d_b, d_c);
    return B200_OK;
}

}  // namespace b200
