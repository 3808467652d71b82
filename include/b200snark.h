/*
 * b200snark.h - C-ABI of the B200-native Groth16 hot path (BN254).
 *
 * This is the drop-in boundary for the two hot paths inside the reference's
 * Groth16::Prover<Engine>::prove (src/groth16.cpp:48-254):
 *   - the multi-scalar multiplications   Curve::multiMulByScalar  (depends/ffiasm/c/curve.hpp:118-121
 *                                        -> ParallelMultiexp::multiexp, depends/ffiasm/c/multiexp.cpp:98-144)
 *   - the Fr-domain transforms           FFT<Field>::fft / ifft   (depends/ffiasm/c/fft.hpp:24-28,
 *                                        fft.cpp:175-212)
 *   - the fused H-polynomial pipeline + five MSMs of prove()      (src/groth16.cpp:52-207)
 * Everything is plain pointers and sizes; no exceptions cross the boundary.
 *
 * Byte conventions (identical to the reference):
 *   field element      32 B little-endian (4 x u64 == 8 x u32 limbs)
 *   Fq / Fq2 / NTT data in Montgomery form, R = 2^256
 *   G1 affine 64 B {x,y}; G2 affine 128 B {x.a,x.b,y.a,y.b}; (0,0) = point at infinity (curve.cpp:534-537)
 *   G1 XYZZ 128 B {x,y,zz,zzz}; G2 XYZZ 256 B; zz == 0 = point at infinity (curve.hpp:11-16)
 *   MSM scalars: plain (non-Montgomery) little-endian integers of `scalar_size` bytes, not reduced
 *
 * Return value: 0 = OK, non-zero = error (B200_ERR_*); text via b200_last_error().  All calls block
 * until the result is in the caller's buffer.  One ctx is used by one host thread at a time (the
 * reference's Prover is driven the same way: src/fullprover.cpp:84-99).
 */
#ifndef B200SNARK_H
#define B200SNARK_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG 1     /* bad argument (null pointer, scalar_size > 32, n not a power of two, ...) */
#define B200_ERR_CUDA 2    /* a CUDA runtime call or kernel failed */
#define B200_ERR_NO_GPU 3  /* no usable CUDA device: the product path has NO CPU fallback */
#define B200_ERR_RANGE 4   /* domain too big for the curve (fft.cpp:70-72) */

typedef struct b200_ctx b200_ctx;
typedef struct b200_zkey b200_zkey;

/* ---- context ------------------------------------------------------------------------------------ */
int b200_init(int device, b200_ctx **out);
void b200_free(b200_ctx *ctx);
const char *b200_last_error(b200_ctx *ctx); /* ctx may be NULL: error of the last failed b200_init */
/* kernels launched by this ctx since creation (bench.py's "gpu_launches") */
uint64_t b200_launch_count(b200_ctx *ctx);
/* the cudaStream_t every kernel of this ctx is launched on (for callers that time with their own events) */
void *b200_stream(b200_ctx *ctx);
/* device milliseconds (CUDA events on the ctx stream) of the last MSM / NTT / prove call, per phase;
 * fills up to `cap` floats, returns how many; names via b200_phase_name(i) */
int b200_last_phase_ms(b200_ctx *ctx, float *out, int cap);
const char *b200_phase_name(int i);
/* with option "timeline" = 1: every timed segment of the last call as a triple (phase index, start ms, end ms), times
 * relative to the first segment's start, in issue order - segments of different streams overlap, so this, not the
 * per-phase sums above, shows the critical path.  Fills up to cap floats (3 per segment); returns the segment count */
int b200_last_timeline(b200_ctx *ctx, float *out, int cap);

/* ---- MSM: replaces Curve::multiMulByScalar (curve.hpp:118-121) ---------------------------------- */
/* host buffers in, XYZZ Montgomery out (any representative of the reference's result point) */
int b200_msm_g1(b200_ctx *ctx, const void *bases_affine, const void *scalars, uint32_t scalar_size,
                uint64_t n, void *out_xyzz128);
int b200_msm_g2(b200_ctx *ctx, const void *bases_affine, const void *scalars, uint32_t scalar_size,
                uint64_t n, void *out_xyzz256);
/* same, bases and scalars already resident in device memory (raw device pointers) */
int b200_msm_g1_dev(b200_ctx *ctx, const void *d_bases_affine, const void *d_scalars,
                    uint32_t scalar_size, uint64_t n, void *out_xyzz128);
int b200_msm_g2_dev(b200_ctx *ctx, const void *d_bases_affine, const void *d_scalars,
                    uint32_t scalar_size, uint64_t n, void *out_xyzz256);
/* window size override for experiments (0 = automatic) */
void b200_set_msm_window(b200_ctx *ctx, int c_bits);
/* tuning knobs for experiments: "msm_window", "acc_smem" (-1 auto / 0 registers / 1 running sum in shared memory /
 * 2 G2 only: running sum and cp.async-staged points in shared memory),
 * "precomp" (-1 auto / 0 off / 1 on: per-window precomputed tables for resident zkeys), "precomp_c",
 * "h_streams" (1 / 3: a, b, c transform chains on one or three streams), "g2_minb", "warm_max", "reduce_l",
 * "reduce_l_g2", "tree_threads", "timeline" (see b200_last_timeline), "fuse_g1" (G1 accumulations sharing a launch: -1 automatic - the three
 * witness MSMs on a single GPU, all four on a shard; 3: witness MSMs; 4: all four; 0: one launch per MSM) */
int b200_set_option(b200_ctx *ctx, const char *name, int value);

/* ---- NTT: replaces FFT<Fr>::fft / ifft (fft.hpp:24-25), natural order in and out ------------------ */
int b200_ntt_fr(b200_ctx *ctx, void *a_host, uint64_t n, int inverse);
int b200_ntt_fr_dev(b200_ctx *ctx, void *d_a, uint64_t n, int inverse);

/* ---- resident zkey + fused prove path (groth16.cpp:9-46 makeProver, :48-207 prove) ---------------- */
typedef struct b200_zkey_desc {
    uint32_t n_vars, n_public, domain_size;
    uint64_t n_coefs;
    const void *coefs;    /* zkey section 4 payload: u32 count, then n_coefs 44-byte records (groth16.hpp:27-35).
                           * May be NULL when shard_count > 1 for a shard that never builds a, b, c itself
                           * (b200_prove_begin with poly_mask = 0: it receives its slices from the owners) */
    const void *points_a; /* n_vars G1 */
    const void *points_b1;/* n_vars G1 */
    const void *points_b2;/* n_vars G2 */
    const void *points_c; /* (n_vars - n_public - 1) G1 */
    const void *points_h; /* domain_size G1 */
    /* point-range shard owned by this ctx (multi-GPU): indices [shard_index*len/shard_count, ...) of
     * every table; shard_count = 1 for a single GPU */
    uint32_t shard_index, shard_count;
    /* optional uneven split (shard_den != 0): this shard owns [len*shard_lo_num/shard_den, len*shard_hi_num/shard_den)
     * of every table instead - ranks that also run a transform chain of the H pipeline get a smaller share of the
     * MSMs (rapidsnark_old_b200/dist.py shard_plan).  The shards of one zkey must tile [0, shard_den). */
    uint32_t shard_lo_num, shard_hi_num, shard_den;
} b200_zkey_desc;

int b200_zkey_upload(b200_ctx *ctx, const b200_zkey_desc *desc, b200_zkey **out);
void b200_zkey_free(b200_zkey *zk);
/* A second handle on the SAME resident tables for another context on the same device: the view owns only its
 * per-proof buffers (witness, a, b, c), so two contexts - two host threads - can prove against one zkey at the same
 * time and the GPU overlaps one proof's tail (last bucket reduction, result read-back, host finalisation) with the
 * next proof's start: the reference's server use (src/fullprover.cpp:84-99) as a throughput mode.  Views and the
 * source may be freed in any order; the tables go with the last of them. */
int b200_zkey_share(b200_ctx *ctx, b200_zkey *src, b200_zkey **out);
/* a, b, c -> h scalars of groth16.cpp:52-163 (normal form, domain_size x 32 B) to a host buffer */
int b200_h_scalars(b200_ctx *ctx, b200_zkey *zk, const void *wtns_host, void *h_out_host);
/* H pipeline + this shard's part of the five MSMs of groth16.cpp:165-207.
 * out768 = pih(128) pi_a(128) pib1(128) pi_b(256) pi_c(128), XYZZ Montgomery, pre-blinding */
int b200_prove_msms(b200_ctx *ctx, b200_zkey *zk, const void *wtns_host, void *out768);
/* same with the witness already in device memory (bench: inputs resident in HBM) */
int b200_prove_msms_dev(b200_ctx *ctx, b200_zkey *zk, const void *d_wtns, void *out768);

/* The whole proof in ONE call - Prover::prove, groth16.cpp:48-253: b200_prove_msms + blinding + to-affine, with the
 * host-side blinding done while the GPU works and in the order the five results arrive (what is left after the last
 * kernel is two group additions and one inversion).  r32, s32: the blinding factors (32-byte little-endian; the
 * reference draws 31 random bytes each, groth16.cpp:213-217).  out_proof256 = A (G1 affine 64 B) | B (G2 affine 128 B)
 * | C (G1 affine 64 B), Montgomery; out_msms768 (may be NULL) = the five pre-blinding points as b200_prove_msms.
 * Single-GPU zkeys only (shard_count = 1). */
typedef struct b200_vkey {
    const void *alpha1, *beta1; /* G1 affine 64 B each (zkey header, zkey_utils.cpp:41-46) */
    const void *beta2;          /* G2 affine 128 B */
    const void *delta1;         /* G1 */
    const void *delta2;         /* G2 */
} b200_vkey;
int b200_groth16_prove(b200_ctx *ctx, b200_zkey *zk, const void *wtns, int wtns_on_device, const b200_vkey *vk,
                       const void *r32, const void *s32, void *out_proof256, void *out_msms768);

/* Two-stage form of b200_prove_msms for zkeys sharded over several GPUs (one ctx per GPU, one process per GPU).
 * The H pipeline of groth16.cpp:101-163 is three independent transform chains (a, b, c) up to the final combine;
 * b200_prove_begin runs only the chains in poly_mask (bit 0 = a, 1 = b, 2 = c) besides uploading the witness and
 * enqueueing this shard's four witness MSMs, and returns WITHOUT synchronising: d_abc3[0..2] = device pointers of the
 * domain_size x 32-byte a, b, c buffers (coset evaluations, Montgomery), *h_stream = the cudaStream_t they are
 * produced on.  The caller exchanges the buffers so that every rank holds, of all three, at least the slice
 * [domain_size * shard_index / shard_count, domain_size * (shard_index + 1) / shard_count) - the only part of h its
 * H MSM reads (e.g. NCCL sends of those slices from the rank that owns the polynomial, enqueued on / ordered after
 * h_stream), then calls b200_prove_finish: combine of that slice -> h, this shard's H MSM, collection of the five
 * partial results (out768 as b200_prove_msms).
 * poly_mask = 7 and no exchange is exactly b200_prove_msms. */
int b200_prove_begin(b200_ctx *ctx, b200_zkey *zk, const void *wtns, int wtns_on_device, uint32_t poly_mask,
                     void **d_abc3, void **h_stream);
int b200_prove_finish(b200_ctx *ctx, b200_zkey *zk, void *out768);

/* Exchange step for ONE process driving n GPUs (ctxs[g] / zks[g] = shard g, all after b200_prove_begin with
 * poly_mask = the polynomials i with i % n == g): of every transformed polynomial each shard receives the slice it
 * will combine from the owner, device to device (cudaMemcpyPeerAsync over NVLink), ordered on the H streams - the
 * in-process counterpart of the NCCL broadcasts of dist.py.  No host synchronisation. */
int b200_exchange_polys(b200_ctx *const *ctxs, b200_zkey *const *zks, int n);

/* ---- synthetic tables: k_i * G for known k_i (fixed-base, device side), affine Montgomery out ------ */
int b200_fixed_base_g1(b200_ctx *ctx, const void *base_affine64, const void *scalars32, uint64_t n, void *out_affine);
int b200_fixed_base_g2(b200_ctx *ctx, const void *base_affine128, const void *scalars32, uint64_t n, void *out_affine);

/* Scalars of the synthetic chain circuit used by bench.py and the tests (the same values as
 * rapidsnark_old_b200/synth.py, computed on the host in C++): inputs are 32-byte little-endian integers below r
 * (toxic waste tau, alpha, beta, gamma, delta and the free wire w_1), outputs normal-form 32-byte integers:
 * wtns[V], a_tau[V], b_tau[V] (A_s(tau), B_s(tau)), c_scalars[V - P - 1], ic_scalars[P + 1], h_tbl[n] (the known
 * discrete logs of the five point tables), dlogs96 = sum w_s A_s | sum w_s B_s | sum_{s <= P} w_s K_s.
 * V = 2^log_n - 6, P = n_public.  CPU only (no ctx). */
int b200_synth_chain(uint32_t log_n, uint32_t n_public, const void *tau32, const void *alpha32, const void *beta32,
                     const void *gamma32, const void *delta32, const void *w1_32, void *wtns, void *a_tau, void *b_tau,
                     void *c_scalars, void *ic_scalars, void *h_tbl, void *dlogs96);

/* ---- host-side group/field helpers (same arithmetic templates as the kernels, compiled for the CPU;
 *      used for the O(1) blinding work of groth16.cpp:209-253 and to fold gathered partial results) --- */
void b200_host_fq_mul(void *r, const void *a, const void *b);
void b200_host_fq_add(void *r, const void *a, const void *b);
void b200_host_fq_sub(void *r, const void *a, const void *b);
void b200_host_fq_neg(void *r, const void *a);
void b200_host_fq_inv(void *r, const void *a);
void b200_host_fr_mul(void *r, const void *a, const void *b);
void b200_host_fr_add(void *r, const void *a, const void *b);
void b200_host_fr_sub(void *r, const void *a, const void *b);
void b200_host_fr_neg(void *r, const void *a);
void b200_host_fr_inv(void *r, const void *a);
void b200_host_fq2_mul(void *r, const void *a, const void *b);
void b200_host_fq2_sqr(void *r, const void *a);
void b200_host_g1_add(void *r_xyzz, const void *a_xyzz, const void *b_xyzz);
void b200_host_g1_madd(void *r_xyzz, const void *a_xyzz, const void *b_affine);
void b200_host_g1_dbl(void *r_xyzz, const void *a_xyzz);
void b200_host_g1_neg(void *r_xyzz, const void *a_xyzz);
void b200_host_g1_to_affine(void *r_affine, const void *a_xyzz);
void b200_host_g1_mul(void *r_xyzz, const void *base_affine, const void *scalar, uint32_t scalar_size);
void b200_host_g2_add(void *r_xyzz, const void *a_xyzz, const void *b_xyzz);
void b200_host_g2_madd(void *r_xyzz, const void *a_xyzz, const void *b_affine);
void b200_host_g2_dbl(void *r_xyzz, const void *a_xyzz);
void b200_host_g2_neg(void *r_xyzz, const void *a_xyzz);
void b200_host_g2_to_affine(void *r_affine, const void *a_xyzz);
void b200_host_g2_mul(void *r_xyzz, const void *base_affine, const void *scalar, uint32_t scalar_size);


/* ---- proof finalisation on the host: blinding + to-affine of src/groth16.cpp:209-253 with explicit r, s
 *      (32-byte little-endian each).  out256 = A (G1 affine 64 B) | B (G2 affine 128 B) | C (G1 affine 64 B),
 *      Montgomery, exactly the fields of Groth16::Proof (groth16.hpp:14-24) */
void b200_groth16_finalize(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                           const void *delta1, const void *delta2, const void *r32, const void *s32, void *out256);
/* The same in two steps, so that the host work which depends only on the verification key and r, s (three G1 and
 * one G2 scalar multiplication - most of it) can run on a host thread WHILE the GPU computes the MSMs:
 * prep640 = r*delta1 | s*delta1 | (rs)*delta1 (G1 XYZZ) | s*delta2 (G2 XYZZ).  finalize = prepare + finalize_prepared. */
void b200_groth16_blind_prepare(const void *delta1, const void *delta2, const void *r32, const void *s32, void *prep640);
void b200_groth16_finalize_prepared(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                                    const void *prep640, const void *r32, const void *s32, void *out256);
/* sum of n gathered per-GPU partial records (n x 768 bytes, layout of b200_prove_msms) into one */
void b200_host_fold_partials(const void *parts768, int n, void *out768);
/* canonical decimal string (<= 78 digits + NUL) of a Montgomery-form Fq element (RawFq::toString) */
void b200_fq_to_decimal(const void *mont32, char *out80);

#ifdef __cplusplus
}
#endif
#endif
