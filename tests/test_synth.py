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
