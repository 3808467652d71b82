"""The synthetic zkey/wtns recipe against the oracle's prove(): the H-table
identity and all five MSM results are checked in the exponent (toxic waste known).  CPU only."""
import pytest

import bn254 as bn
import oracle_lib
import synth_util
from rapidsnark_old_b200 import synth


@pytest.mark.parametrize("impl", ["port", "ref"])
@pytest.mark.parametrize("log_n", [4, 6])
def test_oracle_prove_matches_known_dlogs(impl, log_n):
    o = oracle_lib.port() if impl == "port" else oracle_lib.ref()
    if o is None:
        pytest.skip("oracle/_ref not built")
    s = synth_util.make(log_n)
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    h = o.h_scalars(s.n, s.n_coefs, coefs, wt)
    h_ints = [int.from_bytes(h[i * 32:(i + 1) * 32], "little") for i in range(s.n)]
    assert h_ints == s.h_evals()          # coset evaluations of A*B - C, normal form
    p = s.points
    out = o.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    assert o.msms_to_affine(out) == synth_util.expected_affine(o, s, h_ints)


def test_groth16_identity_in_the_exponent():
    """A*B = alpha*beta + sum_pub w_i K_i + C*delta with the blinded A, B, C of groth16.cpp:222-246."""
    s = synth_util.make(4)
    R = synth.R
    d = s.expected_dlogs()
    r, t = 12345678901234567890, 98765432109876543210
    a = (s.alpha + d["pi_a"] + r * s.delta) % R
    b = (s.beta + d["pi_b"] + t * s.delta) % R
    c = (d["pi_c"] + d["pih"] + t * a + r * b - r * t * s.delta) % R
    pub = sum(s.wtns[i] * s.K[i] for i in range(s.n_public + 1)) % R
    assert a * b % R == (s.alpha * s.beta + pub + c * s.delta) % R


def test_binfile_layout_roundtrip(tmp_path):
    s = synth_util.make(4)
    z = synth.zkey_bytes(s)
    assert z[:4] == b"zkey" and int.from_bytes(z[8:12], "little") == 10
    w = synth.wtns_bytes_file(s)
    assert w[:4] == b"wtns"


@pytest.mark.parametrize("log_n,n_public", [(4, 4), (5, 1), (8, 4), (11, 7)])
def test_fast_synth_equals_python_synth(log_n, n_public):
    """b200_synth_chain (C++ host routine behind synth.FastSynth, what bench.py uses) produces the same witness,
    table scalars, coefficient section and known discrete logs as the Python big-integer generator."""
    a, b = synth.Synth(log_n, 3, n_public), synth.FastSynth(log_n, 3, n_public)
    assert (a.tau, a.alpha, a.beta, a.gamma, a.delta) == (b.tau, b.alpha, b.beta, b.gamma, b.delta)
    assert a.wtns_bytes() == b.wtns_bytes()
    assert synth.le32_many(a.A_tau) == b.A_tau and synth.le32_many(a.B_tau) == b.B_tau
    assert synth.le32_many(a.c_scalars) == b.c_scalars and synth.le32_many(a.ic_scalars) == b.ic_scalars
    assert synth.le32_many(a.h_scalars_tbl) == b.h_scalars_tbl
    assert a.coefs_section() == b.coefs_section() and a.n_coefs == b.n_coefs
    assert (a.dlog_a, a.dlog_b, a.dlog_pub) == (b.dlog_a, b.dlog_b, b.dlog_pub)
    assert synth.count32(b.A_tau) == a.n_vars == synth.count32(a.A_tau)


def _fold_by_single_adds(b200, parts):
    acc = bytearray(parts[0])
    for p in parts[1:]:
        for off, size, add in ((0, 128, b200.host_g1_add), (128, 128, b200.host_g1_add), (256, 128, b200.host_g1_add),
                               (384, 256, b200.host_g2_add), (640, 128, b200.host_g1_add)):
            acc[off:off + size] = add(bytes(acc[off:off + size]), bytes(p[off:off + size]))
    return bytes(acc)


def test_bench_inputs_and_exponent_check_on_cpu():
    """bench.py's input builder (FastSynth + point makers fed packed bytes) and its in-the-exponent proof check,
    end to end on the CPU: the oracle plays the prover."""
    import os
    import sys
    sys.path.insert(0, oracle_lib.ROOT)
    import bench
    import rapidsnark_old_b200 as b200
    o = oracle_lib.best()
    g1m, g2m = synth_util.oracle_point_makers(o)
    unpack = lambda ks: [int.from_bytes(ks[i:i + 32], "little") for i in range(0, len(ks), 32)] \
        if isinstance(ks, (bytes, bytearray)) else ks
    s = bench.build_inputs(6, 2, lambda ks: g1m(unpack(ks)), lambda ks: g2m(unpack(ks)))
    assert isinstance(s, synth.FastSynth)
    p, vk = s.points, s.vk
    msms = o.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section(), p["A"], p["B1"], p["B2"], p["C"],
                        p["H"], s.wtns_bytes())
    r32, s32 = bench.blinding_factors()
    assert len(r32) == len(s32) == 32 and r32[31] == 0 and r32[30] != 0      # 248-bit values, like groth16.cpp:213-217
    proof = b200.groth16_finalize(msms, vk, r32, s32)
    bench.check_known_dlogs(b200, s, msms, proof, r32, s32)
    # the split used by bench.py / the host prover: key-only part on a host thread, the rest after the MSMs
    from concurrent.futures import ThreadPoolExecutor
    from rapidsnark_old_b200 import dist as bdist
    prep = ThreadPoolExecutor(max_workers=1).submit(b200.groth16_blind_prepare, vk, r32, s32)
    folded, proof2 = bdist.finish_proof(msms, vk, r32, s32, prep640=prep.result())
    assert folded == msms and proof2 == proof
    assert proof == o.blind(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], r32, s32)
    assert b200.fold_partials([msms]) == msms
    three = b200.fold_partials([msms, msms, msms])
    assert o.msms_to_affine(three) == o.msms_to_affine(_fold_by_single_adds(b200, [msms, msms, msms]))
    bad = bytearray(proof)
    bad[200] ^= 1
    with pytest.raises(AssertionError):
        bench.check_known_dlogs(b200, s, msms, bytes(bad), r32, s32)


@pytest.mark.parametrize("count", [1, 2, 3, 8])
def test_shard_only_point_tables_are_slices_of_the_full_ones(count):
    """FastSynth.build_shard_points (big multi-GPU runs generate only their own slices) against the full tables, and
    the base-address arithmetic used to hand a slice to b200_zkey_upload."""
    import ctypes
    g1m, g2m = synth_util.oracle_point_makers()
    unpack = lambda ks: [int.from_bytes(ks[i:i + 32], "little") for i in range(0, len(ks), 32)] \
        if isinstance(ks, (bytes, bytearray)) else ks
    mk1, mk2 = (lambda ks: g1m(unpack(ks))), (lambda ks: g2m(unpack(ks)))
    full = synth.FastSynth(5, 2).build_points(mk1, mk2)
    V, P, n = full.n_vars, full.n_public, full.n
    for index in range(count):
        sh = synth.FastSynth(5, 2).build_shard_points(mk1, mk2, index, count)
        assert sh.vk == full.vk
        lo, hi = V * index // count, V * (index + 1) // count
        hlo, hhi = n * index // count, n * (index + 1) // count
        skip = P + 1
        clo, chi = max(lo, skip) - skip, max(hi, skip) - skip
        assert sh.shard_points["A"] == (full.points["A"][64 * lo:64 * hi], lo)
        assert sh.shard_points["B1"] == (full.points["B1"][64 * lo:64 * hi], lo)
        assert sh.shard_points["B2"] == (full.points["B2"][128 * lo:128 * hi], lo)
        assert sh.shard_points["C"] == (full.points["C"][64 * clo:64 * chi], clo)
        assert sh.shard_points["H"] == (full.points["H"][64 * hlo:64 * hhi], hlo)
        for name, size, first, cnt in (("A", 64, lo, hi - lo), ("B2", 128, lo, hi - lo), ("C", 64, clo, chi - clo),
                                       ("H", 64, hlo, hhi - hlo)):
            base = sh.shard_table_address(name)
            got = ctypes.string_at(base + first * size, cnt * size)
            assert got == sh.shard_points[name][0]
