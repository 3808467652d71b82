#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json from the REFERENCE itself (oracle/_ref: the reference's own
curve/multiexp/fft/groth16 templates and its CLI src/main_prover.cpp compiled from /root/reference).
Run in the build container (where /root/reference exists); the JSON is committed and is what the
GPU box compares against - nothing there needs /root/reference.

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bn254 as bn
import oracle_lib
import synth_util
from rapidsnark_old_b200 import synth

FIXED_RS = "1234abcd"


def fixed_rs(seed_hex, ctr, size=31):
    """r / s exactly as oracle/shim/sodium.h derives them under ORACLE_FIXED_RS (ctr = 1 for r, 2 for s)."""
    M = (1 << 64) - 1
    x = (int(seed_hex, 16) + 0x9E3779B97F4A7C15 * ctr) & M
    out = bytearray()
    while len(out) < size:
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out += z.to_bytes(8, "little")
    return bytes(out[:size]) + bytes(32 - size)


def main():
    o = oracle_lib.ref()
    assert o is not None, "oracle/_ref not built (needs /root/reference): make -C oracle ref"
    g = {"version": 1, "source": "oracle/_ref (reference templates + CLI compiled from /root/reference)"}
    r = bn.rng(2024)

    # ---- G1 / G2 MSM vectors (points k*G, unreduced 256-bit scalars, some zero scalars and infinity bases)
    g1, g2 = bn.g1_aff_bytes(bn.G1_GEN), bn.g2_aff_bytes(bn.G2_GEN)
    n1 = 96
    b1 = b"".join(bytes(64) if i % 17 == 5 else o.g1_mul_affine(g1, r.randrange(1, bn.R_ORDER)) for i in range(n1))
    s1 = b"".join(bn.le32(0 if i % 13 == 3 else 1 if i % 13 == 4 else r.getrandbits(256)) for i in range(n1))
    g["msm_g1"] = {"n": n1, "bases": b1.hex(), "scalars": s1.hex(), "affine": o.g1_to_affine(o.g1_msm(b1, s1, n1)).hex()}
    n2 = 40
    b2 = b"".join(bytes(128) if i % 11 == 7 else o.g2_mul_affine(g2, r.randrange(1, bn.R_ORDER)) for i in range(n2))
    s2 = b"".join(bn.le32(r.getrandbits(256)) for _ in range(n2))
    g["msm_g2"] = {"n": n2, "bases": b2.hex(), "scalars": s2.hex(), "affine": o.g2_to_affine(o.g2_msm(b2, s2, n2)).hex()}

    # ---- NTT vectors
    n = 64
    data = b"".join(bn.le32(r.randrange(bn.R_ORDER)) for _ in range(n))
    g["ntt"] = {"n": n, "in": data.hex(), "fft": o.fr_fft(data).hex(), "ifft": o.fr_ifft(data).hex()}

    # ---- a whole proof of the 2^4 synthetic circuit, through the reference CLI with fixed r, s
    s = synth_util.make(4, 1)
    p, vk = s.points, s.vk
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    msms = o.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
    aff = o.msms_to_affine(msms)
    rb, sb = fixed_rs(FIXED_RS, 1), fixed_rs(FIXED_RS, 2)
    proof = o.blind(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], rb, sb)
    with tempfile.TemporaryDirectory() as d:
        zk, wf = os.path.join(d, "c.zkey"), os.path.join(d, "w.wtns")
        open(zk, "wb").write(synth.zkey_bytes(s))
        open(wf, "wb").write(synth.wtns_bytes_file(s))
        env = dict(os.environ, ORACLE_FIXED_RS=FIXED_RS)
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_prover"), zk, wf, os.path.join(d, "proof.json"),
                               os.path.join(d, "public.json")], env=env, cwd=d)
        proof_json = open(os.path.join(d, "proof.json")).read()
        public_json = open(os.path.join(d, "public.json")).read()
    import rapidsnark_old_b200 as b200
    assert b200.proof_json(proof) == proof_json, "ref_blind and the reference CLI disagree?"
    g["circuit_2_4"] = {
        "log_n": 4, "seed": 1, "fixed_rs": FIXED_RS, "r": rb.hex(), "s": sb.hex(),
        "zkey": synth.zkey_bytes(s).hex(), "wtns": synth.wtns_bytes_file(s).hex(),
        "h_scalars": o.h_scalars(s.n, s.n_coefs, coefs, wt).hex(),
        "msms_affine": [a.hex() for a in aff], "proof": proof.hex(),
        "proof_json": proof_json, "public_json": public_json,
    }
    out = os.path.join(HERE, "golden_v1.json")
    json.dump(g, open(out, "w"), indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
