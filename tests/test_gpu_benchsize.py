"""GPU parity AT THE BENCHMARK SIZE (run with -m gpu): the CUDA path through the C-ABI against the reference's own
templates compiled from /root/reference (oracle/_ref, prebuilt, travels to the GPU box) on the SAME inputs - the five
pre-blinding points of Prover::prove (src/groth16.cpp:165-207) and the h scalars (:52-163), bit for bit - at 2^20
constraints with the bench's uniform witness and with a circom-like one (70 % of the wires in {0,1}, 20 % below
2^32, 10 % uniform: SURVEY.md 8d config 2), plus 2^16.  The tables are generated on the GPU by the fixed-base
routine (bench.py does the same); the reference arm costs a few CPU-seconds per proof.
"""
import numpy as np
import pytest

import oracle_lib
import rapidsnark_old_b200 as b200
from rapidsnark_old_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ref():
    o = oracle_lib.ref()
    if o is None:
        pytest.skip("oracle/_ref not prebuilt (no /root/reference where build() ran)")
    return o


def _makers(ctx):
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
    return (lambda ks: ctx.fixed_base_g1(g1, synth.le32_many(ks), synth.count32(ks)),
            lambda ks: ctx.fixed_base_g2(g2, synth.le32_many(ks), synth.count32(ks)))


_circuits = {}


def _circuit(ctx, log_n):
    if log_n not in _circuits:
        s = synth.FastSynth(log_n, 2)
        s.build_points(*_makers(ctx))
        _circuits.clear()            # one big circuit at a time (0.5 GB of tables at 2^20)
        _circuits[log_n] = s
    return _circuits[log_n]


def circom_like_witness(n_vars, seed):
    """70 % of the wires in {0, 1}, 20 % below 2^32, 10 % uniform below 2^253; w_0 = 1.  Not a satisfying assignment:
    the functions under test (a, b, c build, transforms, h, MSMs) are defined for any witness."""
    rng = np.random.default_rng(seed)
    w = np.zeros((n_vars, 4), dtype=np.uint64)
    u = rng.random(n_vars)
    small = u < 0.7
    w[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint64)
    mid = (u >= 0.7) & (u < 0.9)
    w[mid, 0] = rng.integers(0, 1 << 32, size=int(mid.sum()), dtype=np.uint64)
    wide = u >= 0.9
    w[wide] = rng.integers(0, 1 << 61, size=(int(wide.sum()), 4), dtype=np.uint64)   # < 2^253 < r
    w[0] = (1, 0, 0, 0)
    return w.tobytes()


@pytest.mark.parametrize("log_n,witness", [(16, "uniform"), (16, "circom"), (20, "uniform"), (20, "circom")])
def test_prove_msms_and_h_vs_reference_at_bench_size(ctx, ref, log_n, witness):
    s = _circuit(ctx, log_n)
    p = s.points
    coefs = s.coefs_section()
    wt = s.wtns_bytes() if witness == "uniform" else circom_like_witness(s.n_vars, log_n)
    zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"])
    try:
        got_h = zk.h_scalars(wt)
        got = zk.prove_msms(wt)
        again = zk.prove_msms(wt)                      # resident zkey, second proof: same points
    finally:
        zk.free()
    assert got_h == ref.h_scalars(s.n, s.n_coefs, coefs, wt), "h scalars differ from the reference"
    want = ref.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    assert ref.msms_to_affine(got) == ref.msms_to_affine(want), "pre-blinding points differ from the reference"
    assert ref.msms_to_affine(again) == ref.msms_to_affine(want)


def test_h_scalars_with_tma_passes_vs_reference_2_20(ctx, ref):
    """the H pipeline at 2^20 with option "ntt_tma" (TMA-moved tiles, strided and contiguous passes, fused twist)."""
    s = _circuit(ctx, 20)
    p = s.points
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    ctx.set_option("ntt_tma", 1)
    try:
        zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"])
        got = zk.h_scalars(wt)
        zk.free()
    finally:
        ctx.set_option("ntt_tma", 0)
    assert got == ref.h_scalars(s.n, s.n_coefs, coefs, wt)


def test_plain_tables_path_vs_reference_2_16(ctx, ref):
    """precomp = 0: the multi-window path (what the one-shot CLI runs) against the reference at 2^16."""
    s = _circuit(ctx, 16)
    p = s.points
    coefs, wt = s.coefs_section(), circom_like_witness(s.n_vars, 99)
    ctx.set_option("precomp", 0)
    try:
        zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"])
        got = zk.prove_msms(wt)
        zk.free()
    finally:
        ctx.set_option("precomp", -1)
    want = ref.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    assert ref.msms_to_affine(got) == ref.msms_to_affine(want)


@pytest.mark.parametrize("shards", [2, 4])
def test_in_process_exchange_concurrent_large(ctx, ref, shards):
    """b200_exchange_polys at a size where the peer copies (8 MB each) and the owners' in-place combine overlap in
    time unless ordered: all shards' finishes are issued from concurrent host threads, as the C++ Prover does."""
    import threading
    from rapidsnark_old_b200 import dist as bdist
    s = _circuit(ctx, 18)
    p = s.points
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    ctxs = [b200.Context(0) for _ in range(shards)]
    zks = [c.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], i, shards)
           for i, c in enumerate(ctxs)]
    try:
        for rep in range(3):
            for i, zk in enumerate(zks):
                zk.prove_begin(wt, False, bdist.poly_mask(i, shards))
            b200.exchange_polys(zks)
            parts = [None] * shards

            def fin(i):
                parts[i] = zks[i].prove_finish()
            th = [threading.Thread(target=fin, args=(i,)) for i in range(shards)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            folded = b200.fold_partials(parts)
            if rep == 0:
                want = ref.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"],
                                      p["H"], wt)
            assert ref.msms_to_affine(folded) == ref.msms_to_affine(want), "rep %d" % rep
    finally:
        for zk, c in zip(zks, ctxs):
            zk.free()
            c.close()
