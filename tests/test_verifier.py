"""Real pairing verification (tests/pairing.py) of proofs over the synthetic 2^4 circuit: the reference's own
proof (golden, made by the reference CLI), and - on the GPU box - the CUDA path's proof with fresh random r, s.
Negative controls: a tampered proof and a wrong public input must be rejected."""
import json
import os

import pytest

import bn254 as bn
import oracle_lib
import pairing
import rapidsnark_old_b200 as b200
from test_golden import C, _circuit, _sections

ROOT = oracle_lib.ROOT


def _vk_and_public():
    c = _circuit()
    z = _sections(bytes.fromhex(C["zkey"]))
    ic = [bn.g1_aff_from_bytes(z[3][64 * i:64 * i + 64]) for i in range(c["n_public"] + 1)]
    vk = {"alpha1": bn.g1_aff_from_bytes(c["vk"]["alpha1"]), "beta2": bn.g2_aff_from_bytes(c["vk"]["beta2"]),
          "gamma2": bn.g2_aff_from_bytes(c["vk"]["gamma2"]), "delta2": bn.g2_aff_from_bytes(c["vk"]["delta2"]), "IC": ic}
    public = [int(x) for x in json.loads(C["public_json"])]
    return c, vk, public


def _proof_from_json(text):
    p = json.loads(text)
    return {"A": (int(p["pi_a"][0]), int(p["pi_a"][1])),
            "B": ((int(p["pi_b"][0][0]), int(p["pi_b"][0][1])), (int(p["pi_b"][1][0]), int(p["pi_b"][1][1]))),
            "C": (int(p["pi_c"][0]), int(p["pi_c"][1]))}


def test_pairing_is_bilinear():
    P, Q2 = bn.G1_GEN, bn.G2_GEN
    a, b = 7, 11
    lhs = pairing.final_exp(pairing.miller(bn.g2_mul(Q2, b), bn.g1_mul(P, a)))
    rhs = pairing.final_exp(pairing.miller(Q2, P)) ** (a * b)
    assert lhs == rhs and not (lhs == pairing.F12.one())


def test_reference_cli_proof_verifies_and_tampering_is_rejected():
    c, vk, public = _vk_and_public()
    proof = _proof_from_json(C["proof_json"])
    assert pairing.groth16_verify(vk, proof, public)
    bad = dict(proof, C=bn.g1_add(proof["C"], bn.G1_GEN))
    assert not pairing.groth16_verify(vk, bad, public)
    assert not pairing.groth16_verify(vk, proof, [public[0] + 1] + public[1:])


@pytest.mark.gpu
def test_gpu_proof_with_random_blinding_verifies():
    c, vk, public = _vk_and_public()
    ctx = b200.Context(0)
    zk = ctx.zkey_upload(c["n_vars"], c["n_public"], c["domain"], c["n_coefs"], c["coefs"], c["A"], c["B1"], c["B2"],
                         c["C"], c["H"])
    msms = zk.prove_msms(c["wtns"])
    r32, s32 = os.urandom(31) + b"\0", os.urandom(31) + b"\0"
    proof = _proof_from_json(b200.proof_json(b200.groth16_finalize(msms, c["vk"], r32, s32)))
    zk.free()
    ctx.close()
    assert pairing.groth16_verify(vk, proof, public)
