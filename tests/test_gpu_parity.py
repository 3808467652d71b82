"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI
(include/b200snark.h via the ctypes binding), against the CPU oracle on the same inputs, against the
reference's own known-answer vectors, and - at sizes the oracle cannot reach in seconds - through
size-independent properties (known discrete logs, round trips, linearity, shard folding).

Bar: bit-exact.  MSM results are compared as canonical affine points (the XYZZ representative is
free, curve.hpp:11-16); NTT / h data are compared byte for byte.
"""
import ctypes
import os

import pytest

import bn254 as bn
import oracle_lib
import synth_util
import rapidsnark_old_b200 as b200
from rapidsnark_old_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return oracle_lib.best()


def _g1_points(o, n, seed, zero_every=0):
    r = bn.rng(seed)
    g = bn.g1_aff_bytes(bn.G1_GEN)
    pts = []
    for i in range(n):
        if zero_every and i % zero_every == zero_every - 1:
            pts.append(bytes(64))
        else:
            pts.append(o.g1_mul_affine(g, r.randrange(1, bn.R_ORDER)))
    return b"".join(pts)


def _g2_points(o, n, seed, zero_every=0):
    r = bn.rng(seed)
    g = bn.g2_aff_bytes(bn.G2_GEN)
    pts = []
    for i in range(n):
        if zero_every and i % zero_every == zero_every - 1:
            pts.append(bytes(128))
        else:
            pts.append(o.g2_mul_affine(g, r.randrange(1, bn.R_ORDER)))
    return b"".join(pts)


def _scalars(n, seed, kind="full"):
    r = bn.rng(seed)
    out = []
    for i in range(n):
        if kind == "full":
            v = r.getrandbits(256)                     # unreduced, like the reference's MSM bench
        elif kind == "fr":
            v = r.randrange(bn.R_ORDER)
        elif kind == "skew":                           # circom-like: mostly 0/1, some small, few wide
            u = r.random()
            v = 0 if u < 0.35 else 1 if u < 0.7 else r.getrandbits(32) if u < 0.9 else r.randrange(bn.R_ORDER)
        elif kind == "ones":
            v = 1
        elif kind == "max":
            v = (1 << 256) - 1
        else:
            raise ValueError(kind)
        out.append(bn.le32(v))
    return b"".join(out)


# ----------------------------------------------------------------------------- reference KATs on the GPU
def test_multiexp2_golden_kat_gpu(ctx, orc):
    """depends/ffiasm/c/alt_bn128_test.cpp:215-248."""
    pts = [(1626275109576878988287730541908027724405348106427831594181487487855202143055,
            18706364085805828895917702468512381358405767972162700276238017959231481018884),
           (17245156998235704504461341147511350131061011207199931581281143511105381019978,
            3858908536032228066651712470282632925312300188207189106507111128103204506804)]
    sc = [1, 20187316456970436521602619671088988952475789765726813868033071292105413408473]
    bases = b"".join(bn.g1_aff_bytes(P) for P in pts)
    scalars = b"".join(bn.le32(s) for s in sc)
    got = bn.g1_aff_from_bytes(b200.host_g1_to_affine(ctx.msm_g1(bases, scalars, 2)))
    assert got == (9163953212624378696742080269971059027061360176019470242548968584908855004282,
                   20922060990592511838374895951081914567856345629513259026540392951012456141360)


def test_multiexp_algebraic_gpu(ctx, orc):
    """alt_bn128_test.cpp:172-212 at the reference's own size: bases (i+1)G, scalars i+1, n = 40 000."""
    n = 40000
    g = bn.g1_aff_bytes(bn.G1_GEN)
    bases = ctx.fixed_base_g1(g, b"".join(bn.le32(i + 1) for i in range(n)), n)
    scalars = b"".join(bn.le32(i + 1) for i in range(n))
    total = sum((i + 1) ** 2 for i in range(n))
    assert b200.host_g1_to_affine(ctx.msm_g1(bases, scalars, n)) == orc.g1_mul_affine(g, total)


# ----------------------------------------------------------------------------- MSM vs oracle
@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 300, 5000, 70000])
@pytest.mark.parametrize("kind", ["full", "skew"])
def test_msm_g1_vs_oracle(ctx, orc, n, kind):
    bases = _g1_points(orc, min(n, 3000), 100 + n, zero_every=7)
    reps = (n + 2999) // 3000 if n else 0
    bases = (bases * reps)[:64 * n] if n else b""     # repeated points: exercises P == Q inside buckets
    scalars = _scalars(n, 200 + n, kind)
    got = ctx.msm_g1(bases, scalars, n)
    assert orc.g1_to_affine(got) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n))


@pytest.mark.parametrize("n", [0, 1, 2, 17, 300, 4000])
@pytest.mark.parametrize("kind", ["full", "skew"])
def test_msm_g2_vs_oracle(ctx, orc, n, kind):
    bases = _g2_points(orc, min(n, 500), 300 + n, zero_every=5)
    reps = (n + 499) // 500 if n else 0
    bases = (bases * reps)[:128 * n] if n else b""
    scalars = _scalars(n, 400 + n, kind)
    got = ctx.msm_g2(bases, scalars, n)
    assert orc.g2_to_affine(got) == orc.g2_to_affine(orc.g2_msm(bases, scalars, n))


@pytest.mark.parametrize("kind", ["ones", "max"])
def test_msm_degenerate_scalars(ctx, orc, kind):
    n = 2000
    bases = _g1_points(orc, n, 7)
    scalars = _scalars(n, 8, kind)
    assert orc.g1_to_affine(ctx.msm_g1(bases, scalars, n)) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n))


def test_msm_all_zero_and_cancelling(ctx, orc):
    n = 100
    bases = _g1_points(orc, n, 9)
    assert orc.g1_to_affine(ctx.msm_g1(bases, bytes(32 * n), n)) == bytes(64)
    assert orc.g1_to_affine(ctx.msm_g1(bytes(64 * n), _scalars(n, 1), n)) == bytes(64)
    # P and -P with the same scalar cancel
    P = bases[:64]
    negP = P[:32] + bn.to_mont((-bn.from_mont(P[32:64])) % bn.Q)
    sc = bn.le32(123456789) * 2
    assert orc.g1_to_affine(ctx.msm_g1(P + negP, sc, 2)) == bytes(64)


@pytest.mark.parametrize("scalar_size", [1, 8, 16, 31])
def test_msm_short_scalars(ctx, orc, scalar_size):
    n = 500
    r = bn.rng(scalar_size)
    bases = _g1_points(orc, n, 11)
    scalars = bytes(r.getrandbits(8) for _ in range(n * scalar_size))
    got = ctx.msm_g1(bases, scalars, n, scalar_size)
    assert orc.g1_to_affine(got) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n, scalar_size))


@pytest.mark.parametrize("c", [4, 7, 11, 13])
def test_msm_window_override(ctx, orc, c):
    n = 3000
    bases = _g1_points(orc, n, 12)
    scalars = _scalars(n, 13)
    ctx.set_msm_window(c)
    try:
        got = ctx.msm_g1(bases, scalars, n)
    finally:
        ctx.set_msm_window(0)
    assert orc.g1_to_affine(got) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n))


def test_msm_argument_errors(ctx):
    with pytest.raises(b200.B200Error) as e:
        ctx.msm_g1(bytes(64), bytes(64), 1, scalar_size=33)
    assert e.value.code == b200.ERR_ARG


# ----------------------------------------------------------------------------- MSM at full size: known dlogs
@pytest.mark.parametrize("log_n", [16, 20])
def test_msm_g1_known_dlogs_full_size(ctx, orc, log_n):
    n = 1 << log_n
    r = bn.rng(log_n)
    ks = [r.randrange(bn.R_ORDER) for _ in range(n)]
    g = bn.g1_aff_bytes(bn.G1_GEN)
    bases = ctx.fixed_base_g1(g, synth.le32_many(ks), n)
    ss = [r.getrandbits(256) for _ in range(n)]
    total = sum(k * s for k, s in zip(ks, ss)) % bn.R_ORDER
    got = ctx.msm_g1(bases, synth.le32_many(ss), n)
    assert b200.host_g1_to_affine(got) == orc.g1_mul_affine(g, total)


def test_msm_g2_known_dlogs_2_18(ctx, orc):
    n = 1 << 18
    r = bn.rng(5)
    ks = [r.randrange(bn.R_ORDER) for _ in range(n)]
    g = bn.g2_aff_bytes(bn.G2_GEN)
    bases = ctx.fixed_base_g2(g, synth.le32_many(ks), n)
    ss = [r.randrange(bn.R_ORDER) for _ in range(n)]
    total = sum(k * s for k, s in zip(ks, ss)) % bn.R_ORDER
    got = ctx.msm_g2(bases, synth.le32_many(ss), n)
    assert b200.host_g2_to_affine(got) == orc.g2_mul_affine(g, total)


def test_fixed_base_vs_oracle(ctx, orc):
    r = bn.rng(3)
    ks = [0, 1, 2, bn.R_ORDER - 1, (1 << 256) - 1] + [r.getrandbits(256) for _ in range(40)]
    g1, g2 = bn.g1_aff_bytes(bn.G1_GEN), bn.g2_aff_bytes(bn.G2_GEN)
    got1 = ctx.fixed_base_g1(g1, synth.le32_many(ks), len(ks))
    got2 = ctx.fixed_base_g2(g2, synth.le32_many(ks), len(ks))
    for i, k in enumerate(ks):
        assert got1[64 * i:64 * i + 64] == orc.g1_mul_affine(g1, k), i
        assert got2[128 * i:128 * i + 128] == orc.g2_mul_affine(g2, k), i


# ----------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 10, 11, 12, 13, 16])
def test_ntt_vs_oracle(ctx, orc, log_n):
    n = 1 << log_n
    r = bn.rng(log_n)
    data = b"".join(bn.le32(r.randrange(bn.R_ORDER)) for _ in range(n))
    assert ctx.ntt(data) == orc.fr_fft(data)
    assert ctx.ntt(data, inverse=True) == orc.fr_ifft(data)


def test_ntt_reference_roundtrip_kat(ctx):
    """alt_bn128_test.cpp:250-271: ifft(fft(x)) == x for n = 2^10, x_i = i + 1 (Montgomery)."""
    n = 1 << 10
    data = b"".join(bn.to_mont(i + 1, bn.R_ORDER) for i in range(n))
    assert ctx.ntt(ctx.ntt(data), inverse=True) == data


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_large_roundtrip_and_oracle(ctx, orc, log_n):
    """three-pass plan (k > 20) and the bench size: round trip + byte equality with the oracle."""
    import numpy as np
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    raw = rng.integers(0, 1 << 61, size=(n, 4), dtype=np.uint64)   # < 2^253 < r: valid field elements
    data = raw.tobytes()
    fwd = ctx.ntt(data)
    assert ctx.ntt(fwd, inverse=True) == data
    if log_n <= 20 or orc.kind == "reference":       # the OpenMP reference build does 2^22 in seconds
        assert fwd == orc.fr_fft(data)


@pytest.mark.parametrize("log_n", [12, 13, 14, 16, 20, 22])
def test_ntt_tma_variant(ctx, orc, log_n):
    """option "ntt_tma": the passes move their tiles with the TMA engine (cp.async.bulk.tensor, 128-byte swizzle,
    mbarrier) - same bytes as the plain kernel and the oracle, forward, inverse and inside the H pipeline."""
    import numpy as np
    n = 1 << log_n
    rng = np.random.default_rng(1000 + log_n)
    data = rng.integers(0, 1 << 61, size=(n, 4), dtype=np.uint64).tobytes()
    plain_f, plain_i = ctx.ntt(data), ctx.ntt(data, inverse=True)
    ctx.set_option("ntt_tma", 1)
    try:
        tma_f, tma_i = ctx.ntt(data), ctx.ntt(data, inverse=True)
        if log_n <= 14:
            s = synth_util.make(log_n if log_n <= 12 else 12)
            zk = _upload(ctx, s)
            h = zk.h_scalars(s.wtns_bytes())
            zk.free()
            assert h == orc.h_scalars(s.n, s.n_coefs, s.coefs_section(), s.wtns_bytes())
    finally:
        ctx.set_option("ntt_tma", 0)
    assert tma_f == plain_f and tma_i == plain_i
    if log_n <= 16:
        assert tma_f == orc.fr_fft(data)


def test_ntt_domain_too_big(ctx):
    with pytest.raises(b200.B200Error):
        ctx.ntt(bytes(32 * 3))


# ----------------------------------------------------------------------------- H pipeline + five MSMs
def _upload(ctx, s, shard_index=0, shard_count=1):
    p = s.points
    return ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section(), p["A"], p["B1"], p["B2"],
                           p["C"], p["H"], shard_index, shard_count)


@pytest.mark.parametrize("log_n", [4, 6, 10, 12])
def test_h_scalars_and_prove_msms_vs_oracle(ctx, orc, log_n):
    s = synth_util.make(log_n)
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    zk = _upload(ctx, s)
    assert zk.h_scalars(wt) == orc.h_scalars(s.n, s.n_coefs, coefs, wt)
    out = zk.prove_msms(wt)
    p = s.points
    ref = orc.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    assert orc.msms_to_affine(out) == orc.msms_to_affine(ref)
    zk.free()


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_prove_folds_to_unsharded(ctx, orc, shards):
    """multi-GPU partition exercised on one device: shard partials fold to the full result."""
    s = synth_util.make(10)
    wt = s.wtns_bytes()
    parts = []
    for i in range(shards):
        zk = _upload(ctx, s, i, shards)
        parts.append(zk.prove_msms(wt))
        zk.free()
    folded = b200.fold_partials(parts)
    assert orc.msms_to_affine(folded) == synth_util.expected_affine(orc, s)


@pytest.mark.parametrize("shards,chain_cost", [(2, 0.033), (4, 0.15), (8, 0.033)])
def test_uneven_shard_plan_folds_to_unsharded(ctx, orc, shards, chain_cost):
    """explicit shard bounds (b200_zkey_desc.shard_lo_num / hi_num / den, dist.shard_plan: ranks that run a transform
    chain own a smaller point range): the partials still fold to the full result."""
    from rapidsnark_old_b200 import dist as bdist
    s = synth_util.make(10)
    wt = s.wtns_bytes()
    plan = bdist.shard_plan(shards, chain_cost)
    assert plan[0][0] == 0 and plan[-1][1] == bdist.PLAN_DEN and all(plan[i][1] == plan[i + 1][0] for i in range(shards - 1))
    assert plan[0][1] - plan[0][0] < plan[-1][1] - plan[-1][0]          # rank 0 runs a chain: smaller share
    parts = []
    p = s.points
    for i in range(shards):
        zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section(), p["A"], p["B1"], p["B2"], p["C"], p["H"],
                             i, shards, shard_bounds=plan[i] + (bdist.PLAN_DEN,))
        parts.append(zk.prove_msms(wt))
        zk.free()
    assert orc.msms_to_affine(b200.fold_partials(parts)) == synth_util.expected_affine(orc, s)
    with pytest.raises(b200.B200Error):
        ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section(), p["A"], p["B1"], p["B2"], p["C"], p["H"],
                        0, 2, shard_bounds=(5, 3, 8))


def test_two_stage_prove_equals_one_call(ctx, orc):
    """b200_prove_begin(poly_mask = 7) + b200_prove_finish is b200_prove_msms."""
    s = synth_util.make(10)
    wt = s.wtns_bytes()
    zk = _upload(ctx, s)
    one = zk.prove_msms(wt)
    bufs, hstream = zk.prove_begin(wt, False, 7)
    assert all(bufs) and hstream
    two = zk.prove_finish()
    assert orc.msms_to_affine(one) == orc.msms_to_affine(two) == synth_util.expected_affine(orc, s)
    with pytest.raises(b200.B200Error):
        zk.prove_finish()                      # no begin pending
    zk.prove_begin(wt, False, 7)
    with pytest.raises(b200.B200Error):
        zk.prove_begin(wt, False, 7)           # a begin is pending
    assert orc.msms_to_affine(zk.prove_finish()) == orc.msms_to_affine(one)
    zk.free()


@pytest.mark.parametrize("log_n", [6, 12])
def test_fused_prove_is_msms_plus_reference_blinding(ctx, orc, log_n):
    """b200_groth16_prove (one call, blinding interleaved with the collection of the results) = b200_prove_msms followed
    by the reference's blinding (groth16.cpp:209-253, oracle) for the same r, s; also from a device-resident witness."""
    import hashlib
    import torch
    s = synth_util.make(log_n)
    wt = s.wtns_bytes()
    r32 = hashlib.sha256(b"r%d" % log_n).digest()[:31] + b"\0"
    s32 = hashlib.sha256(b"s%d" % log_n).digest()[:31] + b"\0"
    zk = _upload(ctx, s)
    msms, proof = zk.prove(wt, s.vk, r32, s32)
    vk = s.vk
    assert orc.msms_to_affine(msms) == synth_util.expected_affine(orc, s)
    want = orc.blind(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], r32, s32)
    assert proof == want
    d_wt = torch.frombuffer(bytearray(wt), dtype=torch.uint8).cuda()
    msms2, proof2 = zk.prove(d_wt.data_ptr(), s.vk, r32, s32, on_device=True)
    assert proof2 == want and orc.msms_to_affine(msms2) == orc.msms_to_affine(msms)
    assert orc.msms_to_affine(zk.prove_msms(wt)) == orc.msms_to_affine(msms)      # the plain call still works afterwards
    zk.free()
    zk2 = _upload(ctx, s, 0, 2)
    with pytest.raises(b200.B200Error):
        zk2.prove(wt, s.vk, r32, s32)              # one shard of two: no single-call proof
    zk2.free()


def test_zkey_views_prove_concurrently(ctx, orc):
    """b200_zkey_share: three contexts (three host threads) prove against ONE resident zkey at the same time; every
    proof equals the single-context one.  The source may be freed before its views."""
    import hashlib
    import threading
    s = synth_util.make(12)
    wt = s.wtns_bytes()
    r32, s32 = hashlib.sha256(b"vr").digest()[:31] + b"\0", hashlib.sha256(b"vs").digest()[:31] + b"\0"
    zk = _upload(ctx, s)
    _, want = zk.prove(wt, s.vk, r32, s32)
    others = [b200.Context(0) for _ in range(2)]
    views = [c.zkey_share(zk) for c in others]
    got = {}

    def work(i, z):
        got[i] = [z.prove(wt, s.vk, r32, s32)[1] for _ in range(4)]
    th = [threading.Thread(target=work, args=(i, z)) for i, z in enumerate([zk] + views)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(p == want for i in range(3) for p in got[i])
    zk.free()                                   # the tables stay until the last view goes
    assert views[0].prove(wt, s.vk, r32, s32)[1] == want
    with pytest.raises(b200.B200Error):
        others[1].zkey_share(zk)                # freed source
    for z, c in zip(views, others):
        z.free()
        c.close()


@pytest.mark.parametrize("shards", [2, 3, 5])
def test_h_pipeline_spread_over_ranks(orc, shards):
    """The N > 1 flow of dist.prove_msms_distributed on ONE device: every rank is its own context with its own
    shard, runs only the transform chains it owns (dist.poly_mask), the three polynomials are then copied from
    their owners to everybody (what the NCCL broadcasts do), and the folded partials equal the unsharded result."""
    import torch
    from rapidsnark_old_b200 import dist as bdist
    s = synth_util.make(10)
    wt = s.wtns_bytes()
    ctxs = [b200.Context(0) for _ in range(shards)]
    p = s.points
    # ranks that own no chain upload the zkey WITHOUT the coefficient section (they never build a, b, c)
    zks = [c.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section() if bdist.poly_mask(i, shards) else None,
                         p["A"], p["B1"], p["B2"], p["C"], p["H"], i, shards) for i, c in enumerate(ctxs)]
    if shards > 3:
        with pytest.raises(b200.B200Error):
            zks[shards - 1].prove_begin(wt, False, 1)          # no coefficients: cannot run a chain
        with pytest.raises(b200.B200Error):
            zks[shards - 1].h_scalars(wt)
    begun = [zk.prove_begin(wt, False, bdist.poly_mask(i, shards)) for i, zk in enumerate(zks)]
    dev = torch.device("cuda", 0)
    views = [[torch.as_tensor(bdist._DeviceBytes(ptr, s.n * 32), device=dev) for ptr in bufs] for bufs, _ in begun]
    for _, hs in begun:
        torch.cuda.ExternalStream(hs, device=dev).synchronize()
    for poly, owner in enumerate(bdist.poly_owners(shards)):
        for r in range(shards):
            if r != owner:
                views[r][poly].copy_(views[owner][poly])
    torch.cuda.synchronize()
    parts = [zk.prove_finish() for zk in zks]
    assert orc.msms_to_affine(b200.fold_partials(parts)) == synth_util.expected_affine(orc, s)
    for zk, c in zip(zks, ctxs):
        zk.free()
        c.close()


@pytest.mark.parametrize("shards", [2, 4])
def test_in_process_exchange(orc, shards):
    """b200_exchange_polys: the exchange step of ONE process driving N GPUs (the C++ host prover with B200_GPUS=N),
    here N contexts on one device.  Also: the call is refused unless every shard has a begin pending."""
    from rapidsnark_old_b200 import dist as bdist
    s = synth_util.make(10)
    wt = s.wtns_bytes()
    ctxs = [b200.Context(0) for _ in range(shards)]
    zks = [_upload(c, s, i, shards) for i, c in enumerate(ctxs)]
    with pytest.raises(b200.B200Error):
        b200.exchange_polys(zks)
    for i, zk in enumerate(zks):
        zk.prove_begin(wt, False, bdist.poly_mask(i, shards))
    b200.exchange_polys(zks)
    parts = [zk.prove_finish() for zk in zks]
    assert orc.msms_to_affine(b200.fold_partials(parts)) == synth_util.expected_affine(orc, s)
    for zk, c in zip(zks, ctxs):
        zk.free()
        c.close()


def test_zkey_upload_rejects_bad_records(ctx):
    s = synth_util.make(4)
    p = s.points
    bad = bytearray(s.coefs_section())
    bad[4 + 4:4 + 8] = (s.n + 5).to_bytes(4, "little")     # row index out of the domain
    with pytest.raises(b200.B200Error):
        ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, bytes(bad), p["A"], p["B1"], p["B2"], p["C"], p["H"])


@pytest.mark.parametrize("opts", [{"reduce_l": 16}, {"reduce_l": 4, "reduce_l_g2": 2}, {"tree_threads": 32},
                                  {"tree_threads": 128, "warm_max": 2}, {"warm_max": 64}, {"h_streams": 3},
                                  {"g2_minb": 3}, {"log_n": 12, "reduce_l": 8, "tree_threads": 32, "warm_max": 2},
                                  {"lockstep_g1": 1}, {"log_n": 12, "lockstep_g1": 1, "lockstep_g2": 1},
                                  {"log_n": 12, "lockstep_g1": 0, "lockstep_g2": 2}, {"log_n": 12, "g2_minb": 3, "lockstep_g2": 1},
                                  {"log_n": 12, "reduce_l_tail": 8}])
def test_prove_msms_scheduling_options(ctx, orc, opts):
    """Reduce-segment length, tree CTA size, fold threshold, transform streams, G2 occupancy variant, lockstep
    accumulation (one barrier per mixed addition, CTA-uniform trip count): scheduling knobs only - the five points
    never change.  A skewed witness makes the fold paths (warm / hot) do real work and gives the tasks of one CTA
    different lengths."""
    opts = dict(opts)
    s = synth_util.make(opts.pop("log_n", 10))    # 2^12: resident per-window tables, 2^10: plain multi-window path
    wt = bytearray(s.wtns_bytes())
    for i in range(8, s.n_vars, 3):               # two thirds of the wires become 0 / 1: huge |digit| = 1 buckets
        wt[32 * i:32 * i + 32] = (i & 1).to_bytes(32, "little")
    wt = bytes(wt)
    p = s.points
    coefs = s.coefs_section()
    ref = orc.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        zk = _upload(ctx, s)
        assert orc.msms_to_affine(zk.prove_msms(wt)) == orc.msms_to_affine(ref)
        zk.free()
    finally:
        for k in opts:
            ctx.set_option(k, -1 if k == "lockstep_g1" else 0)


@pytest.mark.parametrize("lock_g1,lock_g2", [(1, 1), (1, 2)])
@pytest.mark.parametrize("kind", ["full", "skew"])
def test_msm_lockstep_accumulation(ctx, orc, lock_g1, lock_g2, kind):
    """single MSM calls (k_msm_accumulate with LOCK, 128- and 256-thread CTAs) with tasks of very different lengths
    inside one CTA: finished threads keep arriving at the barrier, the sums are the oracle's."""
    n = 3000
    ctx.set_option("lockstep_g1", lock_g1)
    ctx.set_option("lockstep_g2", lock_g2)
    try:
        b1, sc = _g1_points(orc, 300, 51) * 10, _scalars(n, 52, kind)
        b2 = _g2_points(orc, 300, 53) * 10
        got1, got2 = ctx.msm_g1(b1, sc, n), ctx.msm_g2(b2, sc, n)
    finally:
        ctx.set_option("lockstep_g1", -1)
        ctx.set_option("lockstep_g2", 0)
    assert orc.g1_to_affine(got1) == orc.g1_to_affine(orc.g1_msm(b1, sc, n))
    assert orc.g2_to_affine(got2) == orc.g2_to_affine(orc.g2_msm(b2, sc, n))


@pytest.mark.parametrize("precomp,pc", [(0, 0), (1, 12), (1, 16), (1, 20)])
def test_prove_msms_table_variants(ctx, orc, precomp, pc):
    """resident per-window tables (any window width) and the plain multi-window path give the same points."""
    s = synth_util.make(10)
    ctx.set_option("precomp", precomp)
    ctx.set_option("precomp_c", pc)
    try:
        zk = _upload(ctx, s)
        out = zk.prove_msms(s.wtns_bytes())
        zk.free()
    finally:
        ctx.set_option("precomp", -1)
        ctx.set_option("precomp_c", 0)
    assert orc.msms_to_affine(out) == synth_util.expected_affine(orc, s)


@pytest.mark.parametrize("acc_smem", [0, 1])
@pytest.mark.parametrize("g2", [False, True])
def test_msm_accumulator_variants(ctx, orc, acc_smem, g2):
    n = 3000
    ctx.set_option("acc_smem", acc_smem)
    try:
        if g2:
            bases, scalars = _g2_points(orc, 300, 21) * 10, _scalars(n, 22, "skew")
            got = orc.g2_to_affine(ctx.msm_g2(bases, scalars, n))
            ref = orc.g2_to_affine(orc.g2_msm(bases, scalars, n))
        else:
            bases, scalars = _g1_points(orc, 300, 23) * 10, _scalars(n, 24, "skew")
            got = orc.g1_to_affine(ctx.msm_g1(bases, scalars, n))
            ref = orc.g1_to_affine(orc.g1_msm(bases, scalars, n))
    finally:
        ctx.set_option("acc_smem", -1)
    assert got == ref


@pytest.mark.skipif(not os.environ.get("B200_EXPERIMENTAL"), reason="experimental kernel variant, not measured yet "
                    "(set B200_EXPERIMENTAL=1): G2 accumulation with running sum and points staged in shared memory")
@pytest.mark.parametrize("kind", ["full", "skew"])
def test_msm_g2_staged_accumulation_variant(ctx, orc, kind):
    """option acc_smem = 2 (msm.cuh k_msm_accumulate_staged): cp.async double-buffered points, 168 registers.  The
    addition formula itself (ec_madd_acc_pt) is checked bit-for-bit on the CPU by tests/test_host_cpu.py."""
    n = 4000
    ctx.set_option("acc_smem", 2)
    try:
        bases, scalars = _g2_points(orc, 400, 31) * 10, _scalars(n, 32, kind)
        got = orc.g2_to_affine(ctx.msm_g2(bases, scalars, n))
    finally:
        ctx.set_option("acc_smem", -1)
    assert got == orc.g2_to_affine(orc.g2_msm(bases, scalars, n))


def test_msm_hot_bucket_split(ctx, orc):
    """one bucket far larger than the task cap: sub-tasks + CTA merge (all scalars equal)."""
    n = 20000
    bases = _g1_points(orc, 500, 31) * 40
    scalars = bn.le32(0x1234) * n
    assert orc.g1_to_affine(ctx.msm_g1(bases, scalars, n)) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n))


@pytest.mark.parametrize("batch_log2", [8, 10])
def test_msm_multi_batch_path(ctx, orc, batch_log2):
    """n larger than the sort batch: buckets accumulate across batches (add_existing paths of every kernel)."""
    n = 3000
    bases = _g1_points(orc, 600, 41) * 5
    scalars = _scalars(n, 42, "skew")
    b2 = _g2_points(orc, 300, 43) * 10
    ctx.set_option("max_batch_log2", batch_log2)
    try:
        got1 = ctx.msm_g1(bases, scalars, n)
        got2 = ctx.msm_g2(b2, scalars, n)
    finally:
        ctx.set_option("max_batch_log2", 0)
    assert orc.g1_to_affine(got1) == orc.g1_to_affine(orc.g1_msm(bases, scalars, n))
    assert orc.g2_to_affine(got2) == orc.g2_to_affine(orc.g2_msm(b2, scalars, n))
