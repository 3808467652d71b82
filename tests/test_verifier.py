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


def test_config1_reference_cli_2_10_proof_verifies(tmp_path):
    """BASELINE.json configs[0]: 2^10-constraint synthetic zkey + wtns -> the reference's own CPU prover
    (oracle/_ref/ref_prover = src/main_prover.cpp compiled from /root/reference) -> proof.json / public.json ->
    native pairing verification with the verification key exported from the zkey (tools/export_vkey.py)."""
    import subprocess
    import sys
    import synth_util
    from rapidsnark_old_b200 import synth
    ref_prover = os.path.join(ROOT, "oracle", "_ref", "ref_prover")
    if not os.path.exists(ref_prover):
        pytest.skip("oracle/_ref/ref_prover not built")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import export_vkey
    s = synth_util.make(10)
    zk, wt = tmp_path / "c.zkey", tmp_path / "w.wtns"
    zk.write_bytes(synth.zkey_bytes(s))
    wt.write_bytes(synth.wtns_bytes_file(s))
    subprocess.check_call([ref_prover, str(zk), str(wt), str(tmp_path / "proof.json"), str(tmp_path / "public.json")],
                          cwd=tmp_path)
    vkj = export_vkey.export(zk.read_bytes())
    assert vkj["nPublic"] == 4 and len(vkj["IC"]) == 5
    g1 = lambda p: (int(p[0]), int(p[1]))
    g2 = lambda p: ((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1])))
    vk = {"alpha1": g1(vkj["vk_alpha_1"]), "beta2": g2(vkj["vk_beta_2"]), "gamma2": g2(vkj["vk_gamma_2"]),
          "delta2": g2(vkj["vk_delta_2"]), "IC": [g1(p) for p in vkj["IC"]]}
    proof = _proof_from_json((tmp_path / "proof.json").read_text())
    public = [int(x) for x in json.loads((tmp_path / "public.json").read_text())]
    assert public == s.wtns[1:5]
    assert pairing.groth16_verify(vk, proof, public)


def test_verify_cli_tool_accepts_the_reference_proof_and_rejects_a_tampered_one(tmp_path):
    """tools/verify.py = `snarkjs groth16 verify` (verification_key.json public.json proof.json -> OK! / Invalid proof,
    exit 0 / 1), with the key exported from the zkey by tools/export_vkey.py or on the fly (--zkey)."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import export_vkey
    zk = tmp_path / "c.zkey"
    zk.write_bytes(bytes.fromhex(C["zkey"]))
    (tmp_path / "vk.json").write_text(json.dumps(export_vkey.export(zk.read_bytes())))
    (tmp_path / "public.json").write_text(C["public_json"])
    (tmp_path / "proof.json").write_text(C["proof_json"])
    tool = [sys.executable, os.path.join(ROOT, "tools", "verify.py")]
    r = subprocess.run(tool + [str(tmp_path / "vk.json"), str(tmp_path / "public.json"), str(tmp_path / "proof.json")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK!", r.stdout + r.stderr
    r = subprocess.run(tool + ["--zkey", str(zk), str(tmp_path / "public.json"), str(tmp_path / "proof.json")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK!"
    bad = json.loads(C["proof_json"])
    bad["pi_c"][0], bad["pi_c"][1] = bad["pi_a"][0], bad["pi_a"][1]          # a valid curve point, the wrong one
    (tmp_path / "bad.json").write_text(json.dumps(bad))
    r = subprocess.run(tool + [str(tmp_path / "vk.json"), str(tmp_path / "public.json"), str(tmp_path / "bad.json")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.strip() == "Invalid proof"
    pub = json.loads(C["public_json"])
    (tmp_path / "pub2.json").write_text(json.dumps(pub[:-1]))
    r = subprocess.run(tool + [str(tmp_path / "vk.json"), str(tmp_path / "pub2.json"), str(tmp_path / "proof.json")],
                       capture_output=True, text=True)
    assert r.returncode == 2
