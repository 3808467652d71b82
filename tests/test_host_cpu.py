"""CPU-only tests of everything on the host side of the C-ABI: the library loads and exports every symbol
include/b200snark.h declares, the host arithmetic (the kernels' own templates compiled for the CPU, and
the 4x64 host field used for blinding / Horner) equals the oracle, the product refuses to run without a
GPU, and the CLI / file readers reproduce the reference's error behaviour."""
import ctypes
import os
import re
import subprocess

import pytest

import bn254 as bn
import oracle_lib
import synth_util
import rapidsnark_old_b200 as b200
from rapidsnark_old_b200 import synth

ROOT = oracle_lib.ROOT
PROVER = os.path.join(ROOT, "build", "prover")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200snark.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    L = b200.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert set(b200.EXPORTS) <= declared | {"b200_groth16_finalize", "b200_fq_to_decimal"}


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(b200.B200Error) as e:
        b200.Context(0)
    assert e.value.code == b200.ERR_NO_GPU
    assert "no CPU path" in str(e.value)


def _call(name, size, *args):
    out = ctypes.create_string_buffer(size)
    getattr(b200.lib(), name)(out, *args)
    return out.raw


@pytest.mark.parametrize("name,p", [("fq", bn.Q), ("fr", bn.R_ORDER)])
def test_device_field_algorithm_on_host(name, p):
    """The 8x32-limb even/odd Montgomery product of csrc/field.cuh (host build emulates the PTX carry flag)."""
    r = bn.rng(3)
    crit = [0, 1, 2, p - 1, p - 2, p // 2, (1 << 253) % p, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, bn.MONT_R % p]
    vals = crit + [r.randrange(p) for _ in range(300)]
    rinv = pow(bn.MONT_R, -1, p)
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
        assert int.from_bytes(_call("b200_host_%s_mul" % name, 32, ab, bb), "little") == a * b * rinv % p
        assert int.from_bytes(_call("b200_host_%s_add" % name, 32, ab, bb), "little") == (a + b) % p
        assert int.from_bytes(_call("b200_host_%s_sub" % name, 32, ab, bb), "little") == (a - b) % p
        assert int.from_bytes(_call("b200_host_%s_neg" % name, 32, ab), "little") == (-a) % p
    a = vals[20]
    assert bn.from_mont(_call("b200_host_%s_inv" % name, 32, bn.to_mont(a, p)), p) == pow(a, -1, p)


@pytest.mark.parametrize("karatsuba,lazy,y3seq", [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 1, 1)])
def test_field_product_variants_are_bit_exact(tmp_path, karatsuba, lazy, y3seq):
    """The build-time variants of the device products (field.cuh: Karatsuba 512-bit product, fused a*b - c*d with
    one reduction; fq2.cuh: the lazily reduced Fq2 versions) equal the word-serial Montgomery product on random and
    edge operands, and the accessor form of the mixed addition (ec_madd_acc_pt, experimental staged kernel) equals
    ec_madd including the zero / doubling / cancelling cases - host build of the same source
    (tools/field_variants/host_test.cpp).  y3seq: the Fq2 x*y - z*w with one 512-bit product alive at a time
    (fq2.cuh B200_FQ2_Y3_SEQ, measured slower on B200 and off by default)."""
    exe = tmp_path / "fv"
    src = os.path.join(ROOT, "tools", "field_variants", "host_test.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DB200_KARATSUBA=%d" % karatsuba, "-DB200_LAZY_PAIR=%d" % lazy,
                           "-DB200_FQ2_Y3_SEQ=%d" % y3seq, "-x", "c++", src, "-o", str(exe)])
    r = subprocess.run([str(exe), "30000"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("bad=0") == 5, r.stdout + r.stderr


def test_device_curve_formulas_on_host_vs_oracle():
    o = oracle_lib.best()
    r = bn.rng(5)
    for nm, gb, xs, afs in (("g1", bn.g1_aff_bytes(bn.G1_GEN), 128, 64), ("g2", bn.g2_aff_bytes(bn.G2_GEN), 256, 128)):
        mul, toaff = getattr(o, nm + "_mul"), getattr(o, nm + "_to_affine")
        for i in range(12):
            k1, k2 = r.randrange(bn.R_ORDER), r.randrange(bn.R_ORDER)
            if i == 0: k2 = k1                      # doubling inside add / madd
            if i == 1: k2 = bn.R_ORDER - k1         # P + (-P)
            if i == 2: k2 = 0                       # infinity operand
            P1, P2 = mul(gb, bn.le32(k1)), mul(gb, bn.le32(k2))
            A2 = toaff(P2)
            ref = toaff(getattr(o, nm + "_add")(P1, P2))
            assert toaff(_call("b200_host_%s_add" % nm, xs, P1, P2)) == ref
            assert toaff(_call("b200_host_%s_madd" % nm, xs, P1, A2)) == ref
            assert _call("b200_host_%s_to_affine" % nm, afs, P1) == toaff(P1)
            assert toaff(_call("b200_host_%s_dbl" % nm, xs, P1)) == toaff(getattr(o, nm + "_dbl")(P1))
            assert toaff(_call("b200_host_%s_mul" % nm, xs, gb, bn.le32(k1), ctypes.c_uint32(32))) == toaff(P1)


def test_fq2_kat_on_host():
    """alt_bn128_test.cpp:12-29: (2,2)*(3,3) = (0,12)."""
    e1 = bn.to_mont(2) + bn.to_mont(2)
    e2 = bn.to_mont(3) + bn.to_mont(3)
    assert _call("b200_host_fq2_mul", 64, e1, e2) == bn.to_mont(0) + bn.to_mont(12)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_groth16_finalize_equals_oracle_blind(seed):
    """Host blinding/to-affine (4x64 field) vs the oracle's restatement of groth16.cpp:209-253."""
    o = oracle_lib.best()
    s = synth_util.make(4)
    p = s.points
    msms = o.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, s.coefs_section(), p["A"], p["B1"], p["B2"], p["C"],
                        p["H"], s.wtns_bytes())
    r = bn.rng(seed)
    rb = r.getrandbits(248).to_bytes(32, "little")
    sb = r.getrandbits(248).to_bytes(32, "little")
    vk = s.vk
    ref = o.blind(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], rb, sb)
    out = ctypes.create_string_buffer(256)
    b200.lib().b200_groth16_finalize(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], rb, sb, out)
    assert out.raw == ref
    # and the proof satisfies the Groth16 equation in the exponent
    R = synth.R
    d = s.expected_dlogs()
    ri, si = int.from_bytes(rb, "little"), int.from_bytes(sb, "little")
    a = (s.alpha + d["pi_a"] + ri * s.delta) % R
    assert out.raw[:64] == o.g1_mul_affine(synth.g1_gen_bytes(), a)


def test_fq_to_decimal():
    for v in (0, 1, 10, bn.Q - 1, 12345678901234567890123456789012345678901234567890):
        buf = ctypes.create_string_buffer(80)
        b200.lib().b200_fq_to_decimal(bn.to_mont(v % bn.Q), buf)
        assert buf.value.decode() == str(v % bn.Q)


# ----------------------------------------------------------------------------- CLI / readers
def _run(args):
    return subprocess.run([PROVER] + args, capture_output=True, text=True)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    s = synth_util.make(4)
    zk, wt = d / "c.zkey", d / "w.wtns"
    zk.write_bytes(synth.zkey_bytes(s))
    wt.write_bytes(synth.wtns_bytes_file(s))
    return d, str(zk), str(wt)


def test_cli_usage_message():
    r = _run([])
    assert r.returncode != 0
    assert "Usage: prover <circuit.zkey> <witness.wtns> <proof.json> <public.json>" in r.stderr


def test_cli_rejects_wrong_magic_and_version(files):
    d, zk, wt = files
    r = _run([wt, wt, str(d / "p.json"), str(d / "pub.json")])          # wtns given as zkey
    assert r.returncode != 0 and "Invalid file type. It should be zkey and it us wtns" in r.stderr
    bad = d / "v9.zkey"
    raw = bytearray(open(zk, "rb").read())
    raw[4:8] = (9).to_bytes(4, "little")
    bad.write_bytes(bytes(raw))
    r = _run([str(bad), wt, str(d / "p.json"), str(d / "pub.json")])
    assert r.returncode != 0 and "Invalid version. It should be <=1 and it us 9" in r.stderr


def test_cli_rejects_other_curve_and_protocol(files):
    d, zk, wt = files
    raw = bytearray(open(zk, "rb").read())
    i = raw.index(synth.R.to_bytes(32, "little"))
    raw[i] ^= 2
    bad = d / "curve.zkey"
    bad.write_bytes(bytes(raw))
    r = _run([str(bad), wt, str(d / "p.json"), str(d / "pub.json")])
    assert r.returncode != 0 and "zkey curve not supported" in r.stderr
    raw = bytearray(open(zk, "rb").read())
    raw[12 + 12] = 2                                      # section 1 payload: protocol id
    bad = d / "proto.zkey"
    bad.write_bytes(bytes(raw))
    r = _run([str(bad), wt, str(d / "p.json"), str(d / "pub.json")])
    assert r.returncode != 0 and "zkey file is not groth16" in r.stderr


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_cli_fails_loudly_without_gpu(files):
    d, zk, wt = files
    r = _run([zk, wt, str(d / "p.json"), str(d / "pub.json")])
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr
    assert not os.path.exists(d / "p.json")


# ----------------------------------------------------------------------------- engine surface + JSON helper (host C++)
def _compile_and_run(tmp_path, name, src, extra=()):
    f = tmp_path / (name + ".cpp")
    f.write_text(src)
    exe = tmp_path / name
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "rapidsnark_old_b200", "src"),
                           "-I" + os.path.join(ROOT, "include"), str(f), "-o", str(exe), "-L" + os.path.join(ROOT, "rapidsnark_old_b200"),
                           "-lb200snark", "-Wl,-rpath," + os.path.join(ROOT, "rapidsnark_old_b200")] + list(extra))
    return subprocess.run([str(exe)], capture_output=True, text=True)


def test_engine_surface_of_the_reference(tmp_path):
    """AltBn128::Engine with f1, f2, fr, g1, g2 members and the global F1 / Fr / G1 objects (depends/ffiasm/c/
    alt_bn128.hpp:22-59) as rapidsnark_old_b200/src/engine/alt_bn128.hpp provides them: the calls the reference's
    groth16.cpp makes (fr.mul / add / toMontgomery / toString, g1.mulByScalar / add / sub / copy / toString)."""
    src = r"""
#include "alt_bn128.hpp"
#include <cstdio>
int main() {
    AltBn128::Engine &E = AltBn128::Engine::engine;
    AltBn128::FrElement a, b, c, m;
    E.fr.copy(a, E.fr.one()); E.fr.add(b, a, a); E.fr.mul(c, b, b);
    memset(&m, 0, sizeof m); m.v[0] = 7; E.fr.toMontgomery(m, m); E.fr.mul(m, m, c); E.fr.sub(m, m, a);
    printf("%s %s %s\n", E.fr.toString(c).c_str(), AltBn128::Fr.toString(b).c_str(), E.fr.toString(m).c_str());
    AltBn128::G1PointAffine g; AltBn128::F1.copy(g.x, AltBn128::F1.one()); AltBn128::F1.add(g.y, g.x, g.x);
    AltBn128::G1Point p, q; uint8_t k[32] = {5};
    E.g1.mulByScalar(p, g, k, 32); E.g1.add(q, p, g); E.g1.sub(q, q, p);
    AltBn128::G1PointAffine r; E.g1.copy(r, q);
    printf("%s %s\n%s\n", E.f1.toString(r.x).c_str(), E.f1.toString(r.y).c_str(), E.g1.toString(p).c_str());
    return 0;
}
"""
    r = _compile_and_run(tmp_path, "eng", src)
    assert r.returncode == 0, r.stderr
    five_g = bn.g1_mul(bn.G1_GEN, 5)
    assert r.stdout.splitlines() == ["4 2 27", "1 2", "(%d,%d)" % five_g]


def test_minijson_parse_and_dump(tmp_path):
    """src/minijson.hpp: what FullProver does to the request body (parse, then nlohmann's compact dump with sorted
    keys: src/fullprover.cpp:108-113); malformed text is refused."""
    src = r"""
#include "minijson.hpp"
#include <cstdio>
#include <iostream>
int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        try { std::cout << minijson::dump(minijson::parse(line)) << "\n"; }
        catch (std::runtime_error &e) { std::cout << "ERR\n"; }
    }
    return 0;
}
"""
    f = tmp_path / "mj.cpp"
    f.write_text(src)
    exe = tmp_path / "mj"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "rapidsnark_old_b200", "src"), str(f), "-o", str(exe)])
    cases = [
        ('{"b": [1, 2, "3"], "a": {"y": null, "x": true}}', '{"a":{"x":true,"y":null},"b":[1,2,"3"]}'),
        ('  [ ]  ', '[]'), ('{}', '{}'), ('"a\\nb\\u00e9\\ud83d\\ude00"', None),
        ('{"in": ["21888242871839275222246405745257275088548364400416034343698204186575808495616", -1.5e3]}',
         '{"in":["21888242871839275222246405745257275088548364400416034343698204186575808495616",-1.5e3]}'),
        ('{"a": 1,}', 'ERR'), ('{"a" 1}', 'ERR'), ('[1, 2', 'ERR'), ('{"a": 01}', 'ERR'), ('nul', 'ERR'), ('"\\ud800"', 'ERR'),
        ('{"a": 1} x', 'ERR'), ('', 'ERR'),
    ]
    r = subprocess.run([str(exe)], input="\n".join(c for c, _ in cases) + "\n", capture_output=True, text=True)
    out = r.stdout.splitlines()
    assert len(out) == len(cases), r.stdout
    import json as pyjson
    for (text, want), got in zip(cases, out):
        if want is None:                         # round trip through Python's parser instead of a literal
            assert pyjson.loads(got) == pyjson.loads(text), (text, got)
        else:
            assert got == want, (text, got)


def test_two_lane_bucket_reduction_schedule():
    """The schedule of k_msm_reduce_segments (msm.cuh): lane 0 walks run += B_k, lane 1 adds the value lane 0 held
    after the PREVIOUS step; after L + 1 steps lane 1 holds sum (k + 1) B_k and lane 0 the plain sum - the same
    values as the one-thread recursion  run += B_k; acc += run  (multiexp.cpp:62-96 computes the same sums
    by another recursion).  Modelled on integers: the group law is only ever used as an associative addition."""
    r = bn.rng(11)
    for L in (1, 2, 8, 16, 32, 128):
        B = [r.randrange(1 << 64) if r.randrange(4) else 0 for _ in range(L)]      # a quarter of the buckets empty
        mine = [0, 0]          # lane 0: run, lane 1: acc
        handed = 0
        for k in range(L - 1, -2, -1):
            q0 = B[k] if k >= 0 else 0
            mine = [mine[0] + q0, mine[1] + handed]
            handed = mine[0]                                                       # shuffle from the even lane
        run = acc = 0
        for k in range(L - 1, -1, -1):
            run += B[k]
            acc += run
        assert mine[0] == run == sum(B)
        assert mine[1] == acc == sum((k + 1) * B[k] for k in range(L))


def test_task_plan_cta_aggregation_model():
    """k_msm_plan_count / k_msm_plan_place (msm.cuh) aggregate the task-length histogram and the placement ranks per
    CTA before touching the global arrays.  Model of the two kernels + the scan between them: every task lands in a
    distinct slot, slots are ordered by length descending, every bucket's pieces are all there."""
    r = bn.rng(12)
    CAP, MAXCAP, CTA = 8, 2048, 64
    counts = [r.choice([0, 1, 3, 7, 8, 9, 20, 33]) for _ in range(1000)]
    lenhist = [0] * MAXCAP
    for c0 in range(0, len(counts), CTA):                      # plan_count: one shared histogram per CTA, then flushed
        sh = {}
        for cnt in counts[c0:c0 + CTA]:
            if cnt:
                nfull, rem = divmod(cnt, CAP)
                if nfull:
                    sh[MAXCAP - CAP] = sh.get(MAXCAP - CAP, 0) + nfull
                if rem:
                    sh[MAXCAP - rem] = sh.get(MAXCAP - rem, 0) + 1
        for i, v in sh.items():
            lenhist[i] += v
    cursor, acc = [0] * MAXCAP, 0
    for i in range(MAXCAP):                                    # exclusive scan: longest tasks (smallest index) first
        cursor[i], acc = acc, acc + lenhist[i]
    ntasks = acc
    tasks = [None] * ntasks
    for c0 in range(0, len(counts), CTA):                      # plan_place: local ranks, one global add per bin and CTA
        sh_cnt, ranks = {}, []
        for b in range(c0, min(c0 + CTA, len(counts))):
            nfull, rem = divmod(counts[b], CAP)
            rf = rr = None
            if nfull:
                rf = sh_cnt.get(MAXCAP - CAP, 0)
                sh_cnt[MAXCAP - CAP] = rf + nfull
            if rem:
                rr = sh_cnt.get(MAXCAP - rem, 0)
                sh_cnt[MAXCAP - rem] = rr + 1
            ranks.append((b, nfull, rem, rf, rr))
        base = {}
        for i, v in sh_cnt.items():
            base[i], cursor[i] = cursor[i], cursor[i] + v
        for b, nfull, rem, rf, rr in ranks:
            for j in range(nfull):
                assert tasks[base[MAXCAP - CAP] + rf + j] is None
                tasks[base[MAXCAP - CAP] + rf + j] = (b, j, CAP)
            if rem:
                assert tasks[base[MAXCAP - rem] + rr] is None
                tasks[base[MAXCAP - rem] + rr] = (b, nfull, rem)
    assert all(t is not None for t in tasks)
    lens = [t[2] for t in tasks]
    assert lens == sorted(lens, reverse=True)
    per_bucket = {}
    for b, piece, ln in tasks:
        per_bucket.setdefault(b, []).append(ln)
    assert all(sum(per_bucket.get(b, [])) == counts[b] for b in range(len(counts)))
