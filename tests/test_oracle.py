"""Pins the oracle (CPU, no GPU): the plain-C restatement (orc_*) and the compiled-from-reference
library (ref_*) against (1) the reference's own golden vectors and known-answer tests
(depends/ffiasm/c/alt_bn128_test.cpp), (2) independent Python big-integer arithmetic, (3) each other.
"""
import ctypes
import os
import subprocess

import pytest

import bn254 as bn
import oracle_lib

ROOT = oracle_lib.ROOT
IMPLS = ["port", "ref"]


def get(impl):
    o = oracle_lib.port() if impl == "port" else oracle_lib.ref()
    if o is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return o


# ----------------------------------------------------------------------------- L0 field routines
def _field_lib():
    oracle_lib.port()
    return ctypes.CDLL(oracle_lib.PORT_SO)


def _critical(p):
    """Critical numbers in the spirit of ffiasm/test/fieldasm.js:363-462."""
    vals = {0, 1, 2, p - 1, p - 2, p // 2, p // 2 + 1, (1 << 64) - 1, 1 << 64, (1 << 64) + 1,
            (1 << 128) - 1, 1 << 128, (1 << 192) + 5, (1 << 253) % p, bn.MONT_R % p, (bn.MONT_R * bn.MONT_R) % p}
    r = bn.rng(7)
    vals |= {r.randrange(p) for _ in range(24)}
    return sorted(vals)


@pytest.mark.parametrize("backend", ["adx", "c"])
@pytest.mark.parametrize("name,p", [("Fq", bn.Q), ("Fr", bn.R_ORDER)])
def test_field_ops_against_bigints(name, p, backend):
    """Both backends of the oracle's Montgomery product - the ADX assembly (mulx / adcx / adox, like the reference's
    generated routines) and the plain-C CIOS - against Python big integers on the critical numbers."""
    lib = _field_lib()
    lib.oracle_set_field_backend(1 if backend == "adx" else 0)
    if backend == "adx" and not lib.oracle_field_backend():
        pytest.skip("CPU without ADX / BMI2")
    try:
        _field_ops_check(lib, name, p)
    finally:
        lib.oracle_set_field_backend(1)


def _field_ops_check(lib, name, p):
    vals = _critical(p)
    rinv = pow(bn.MONT_R, -1, p)
    buf = lambda: ctypes.create_string_buffer(32)
    for a in vals:
        ab = a.to_bytes(32, "little")
        o = buf(); getattr(lib, name + "_rawNeg")(o, ab)
        assert int.from_bytes(o.raw, "little") == (-a) % p
        o = buf(); getattr(lib, name + "_rawToMontgomery")(o, ab)
        assert int.from_bytes(o.raw, "little") == a * bn.MONT_R % p
        o = buf(); getattr(lib, name + "_rawFromMontgomery")(o, ab)
        assert int.from_bytes(o.raw, "little") == a * rinv % p
        o = buf(); getattr(lib, name + "_rawMSquare")(o, ab)
        assert int.from_bytes(o.raw, "little") == a * a * rinv % p
        if a:
            o = buf(); getattr(lib, name + "_rawInv")(o, ab)
            # Montgomery in, Montgomery out (fr.cpp.ejs:215-227)
            assert int.from_bytes(o.raw, "little") == pow(a * rinv, -1, p) * bn.MONT_R % p
        for b in vals[::3]:
            bb = b.to_bytes(32, "little")
            o = buf(); getattr(lib, name + "_rawAdd")(o, ab, bb)
            assert int.from_bytes(o.raw, "little") == (a + b) % p
            o = buf(); getattr(lib, name + "_rawSub")(o, ab, bb)
            assert int.from_bytes(o.raw, "little") == (a - b) % p
            o = buf(); getattr(lib, name + "_rawMMul")(o, ab, bb)
            assert int.from_bytes(o.raw, "little") == a * b * rinv % p
            assert getattr(lib, name + "_rawIsEq")(ab, bb) == int(a == b)
        assert getattr(lib, name + "_rawIsZero")(ab) == int(a == 0)


def test_constants_match_reference_sample_asm():
    """q, R2, R3, np of Fq as printed in the reference's pre-generated sample
    (depends/ffiasm/benchmark/fr.asm:7098-7103; that file is for the Fq prime)."""
    lib = _field_lib()
    lib.Fq_rawR2_ptr.restype = ctypes.POINTER(ctypes.c_uint64)
    lib.Fq_rawR3_ptr.restype = ctypes.POINTER(ctypes.c_uint64)
    lib.Fq_rawq_ptr.restype = ctypes.POINTER(ctypes.c_uint64)
    rd = lambda ptr: sum(ptr[i] << (64 * i) for i in range(4))
    assert rd(lib.Fq_rawq_ptr()) == bn.Q
    assert rd(lib.Fq_rawR2_ptr()) == 0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89
    assert rd(lib.Fq_rawR3_ptr()) == 0x20fd6e902d592544ef7f0b0c0ada0afb62f210e6a7283db6b1cd6dafda1530df


# ----------------------------------------------------------------------------- reference KATs
@pytest.mark.parametrize("impl", IMPLS)
def test_f2_simple_mul_kat(impl):
    """alt_bn128_test.cpp:12-29: (2,2)*(3,3) = (0,12)."""
    o = get(impl)
    e1 = bn.to_mont(2) + bn.to_mont(2)
    e2 = bn.to_mont(3) + bn.to_mont(3)
    assert o.fq2_mul(e1, e2) == bn.to_mont(0) + bn.to_mont(12)


@pytest.mark.parametrize("impl", IMPLS)
def test_multiexp2_golden_kat(impl):
    """alt_bn128_test.cpp:215-248: two explicit points, scalars 1 and 2018...473, explicit affine result."""
    o = get(impl)
    pts = [(1626275109576878988287730541908027724405348106427831594181487487855202143055,
            18706364085805828895917702468512381358405767972162700276238017959231481018884),
           (17245156998235704504461341147511350131061011207199931581281143511105381019978,
            3858908536032228066651712470282632925312300188207189106507111128103204506804)]
    sc = [1, 20187316456970436521602619671088988952475789765726813868033071292105413408473]
    bases = b"".join(bn.g1_aff_bytes(P) for P in pts)
    scalars = b"".join(bn.le32(s) for s in sc)
    got = bn.g1_aff_from_bytes(o.g1_to_affine(o.g1_msm(bases, scalars, 2)))
    assert got == (9163953212624378696742080269971059027061360176019470242548968584908855004282,
                   20922060990592511838374895951081914567856345629513259026540392951012456141360)
    assert got == bn.g1_msm(pts, sc)


@pytest.mark.parametrize("impl", IMPLS)
def test_multiexp_algebraic(impl):
    """alt_bn128_test.cpp:172-212: bases (i+1)G, scalars i+1 -> (sum (i+1)^2) G.  4 000 points here
    (the reference uses 40 000; ref_kat runs that size)."""
    o = get(impl)
    n = 4000
    g = bn.g1_aff_bytes(bn.G1_GEN)
    one = o.g1_mul(g, bn.le32(1))
    acc = one
    bases = [o.g1_to_affine(acc)]
    for _ in range(1, n):
        acc = o.g1_madd(acc, g)
        bases.append(o.g1_to_affine(acc))
    scalars = b"".join(bn.le32(i + 1) for i in range(n))
    total = sum((i + 1) ** 2 for i in range(n))
    got = o.g1_to_affine(o.g1_msm(b"".join(bases), scalars, n))
    assert got == o.g1_mul_affine(g, total)
    assert bn.g1_aff_from_bytes(got) == bn.g1_mul(bn.G1_GEN, total)


@pytest.mark.parametrize("impl", IMPLS)
def test_group_order_and_small_multiples(impl):
    """alt_bn128_test.cpp:31-170: P+0, P-P, 3P/4P/5P/65P consistency, r*G1 = 0, r*G2 = 0."""
    o = get(impl)
    g1 = bn.g1_aff_bytes(bn.G1_GEN)
    g2 = bn.g2_aff_bytes(bn.G2_GEN)
    assert o.g1_mul_affine(g1, bn.R_ORDER) == bytes(64)
    assert o.g2_mul_affine(g2, bn.R_ORDER) == bytes(128)
    for k in (1, 2, 3, 4, 5, 8, 65, 2 ** 200 + 12345):
        assert bn.g1_aff_from_bytes(o.g1_mul_affine(g1, k)) == bn.g1_mul(bn.G1_GEN, k)
    for k in (1, 2, 3, 65, 2 ** 130 + 7):
        assert bn.g2_aff_from_bytes(o.g2_mul_affine(g2, k)) == bn.g2_mul(bn.G2_GEN, k)
    one = o.g1_mul(g1, bn.le32(1))
    two = o.g1_dbl(one)
    four = o.g1_dbl(two)
    three = o.g1_madd(two, g1)
    assert o.g1_to_affine(o.g1_add(three, one)) == o.g1_to_affine(four)
    # P + (-P) = 0 through the generic add path
    neg_g = bn.g1_aff_bytes(bn.g1_neg(bn.G1_GEN))
    assert o.g1_to_affine(o.g1_madd(one, neg_g)) == bytes(64)
    # doubling through add(P, P)
    assert o.g1_to_affine(o.g1_add(one, one)) == o.g1_to_affine(two)
    assert o.g1_to_affine(o.g1_madd(one, g1)) == o.g1_to_affine(two)


@pytest.mark.parametrize("impl", IMPLS)
def test_ntt_roundtrip_and_naive(impl):
    """alt_bn128_test.cpp:250-271 (ifft(fft(x)) == x, n = 2^10, x_i = i+1) plus an O(n^2) DFT check."""
    o = get(impl)
    n = 1 << 10
    a = b"".join(bn.to_mont(i + 1, bn.R_ORDER) for i in range(n))
    assert o.fr_ifft(o.fr_fft(a)) == a
    n = 32
    r = bn.rng(3)
    vals = [r.randrange(bn.R_ORDER) for _ in range(n)]
    data = b"".join(bn.to_mont(v, bn.R_ORDER) for v in vals)
    fwd = o.fr_fft(data)
    assert [bn.from_mont(fwd[i * 32:(i + 1) * 32], bn.R_ORDER) for i in range(n)] == bn.ntt_naive(vals)
    inv = o.fr_ifft(data)
    assert [bn.from_mont(inv[i * 32:(i + 1) * 32], bn.R_ORDER) for i in range(n)] == bn.ntt_naive(vals, inverse=True)
    # root(domainPow, idx), fft.hpp:28 / SURVEY Appendix B
    assert bn.from_mont(o.fr_root(21, 1), bn.R_ORDER) == \
        13536764371732269273912573961853310557438878140379554347802702086337840854307
    assert bn.from_mont(o.fr_root(5, 3), bn.R_ORDER) == pow(bn.fr_root_of_unity(5), 3, bn.R_ORDER)


def test_ref_kat_binary_passes():
    """The reference's ffiasm/c/alt_bn128_test.cpp, compiled unmodified against the L0 restatement
    (13 cases incl. the 40 000-point multiExp)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_kat")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_kat not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    assert "13 tests, 0 failed" in out.stdout


# ----------------------------------------------------------------------------- port == reference
def _rand_g1_points(o, n, seed, zero_every=0):
    r = bn.rng(seed)
    g = bn.g1_aff_bytes(bn.G1_GEN)
    pts = []
    for i in range(n):
        if zero_every and i % zero_every == zero_every - 1:
            pts.append(bytes(64))
        else:
            pts.append(o.g1_mul_affine(g, r.randrange(1, bn.R_ORDER)))
    return b"".join(pts)


def _rand_g2_points(o, n, seed):
    r = bn.rng(seed)
    g = bn.g2_aff_bytes(bn.G2_GEN)
    return b"".join(o.g2_mul_affine(g, r.randrange(1, bn.R_ORDER)) for _ in range(n))


def _rand_scalars(n, seed, full_width=True):
    r = bn.rng(seed)
    return b"".join(bn.le32(r.getrandbits(256) if full_width else r.randrange(bn.R_ORDER)) for _ in range(n))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 300, 5000])
def test_msm_port_equals_reference(n):
    p, rf = get("port"), get("ref")
    bases = _rand_g1_points(p, n, 100 + n, zero_every=7)
    scalars = _rand_scalars(n, 200 + n)          # unreduced 256-bit scalars, as the reference's bench
    assert p.g1_to_affine(p.g1_msm(bases, scalars, n)) == rf.g1_to_affine(rf.g1_msm(bases, scalars, n))
    if n <= 300:
        b2 = _rand_g2_points(p, n, 300 + n)
        assert p.g2_to_affine(p.g2_msm(b2, scalars, n)) == rf.g2_to_affine(rf.g2_msm(b2, scalars, n))


def test_msm_small_against_bigints():
    p = get("port")
    n = 9
    r = bn.rng(11)
    ks = [r.randrange(1, bn.R_ORDER) for _ in range(n)]
    pts = [bn.g1_mul(bn.G1_GEN, k) for k in ks]
    sc = [r.getrandbits(256) for _ in range(n)]
    bases = b"".join(bn.g1_aff_bytes(P) for P in pts)
    got = bn.g1_aff_from_bytes(p.g1_to_affine(p.g1_msm(bases, b"".join(bn.le32(s) for s in sc), n)))
    assert got == bn.g1_msm(pts, sc)
    pts2 = [bn.g2_mul(bn.G2_GEN, k) for k in ks[:4]]
    bases2 = b"".join(bn.g2_aff_bytes(P) for P in pts2)
    got2 = bn.g2_aff_from_bytes(p.g2_to_affine(p.g2_msm(bases2, b"".join(bn.le32(s) for s in sc[:4]), 4)))
    assert got2 == bn.g2_msm(pts2, sc[:4])


def test_ntt_port_equals_reference():
    p, rf = get("port"), get("ref")
    for log_n in (1, 2, 7, 12):
        n = 1 << log_n
        r = bn.rng(log_n)
        data = b"".join(bn.le32(r.randrange(bn.R_ORDER)) for _ in range(n))
        assert p.fr_fft(data) == rf.fr_fft(data)
        assert p.fr_ifft(data) == rf.fr_ifft(data)
