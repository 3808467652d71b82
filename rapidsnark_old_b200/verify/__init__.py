"""Groth16 verification for BN254 proofs in the snarkjs / rapidsnark JSON layouts: what
`snarkjs groth16 verify verification_key.json public.json proof.json` does, natively.  Pure Python big-integer pairing
(pairing.py) - verification is O(1) work and not on the prover's hot path.  CLI: tools/verify.py."""
import json

from . import pairing


def _g1(p):
    return None if int(p[2]) == 0 else (int(p[0]), int(p[1]))


def _g2(p):
    if int(p[2][0]) == 0 and int(p[2][1]) == 0:
        return None
    return ((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1])))


def vkey_from_json(vkj):
    if vkj.get("protocol", "groth16") != "groth16":
        raise ValueError("not a groth16 verification key")
    return {"alpha1": _g1(vkj["vk_alpha_1"]), "beta2": _g2(vkj["vk_beta_2"]), "gamma2": _g2(vkj["vk_gamma_2"]),
            "delta2": _g2(vkj["vk_delta_2"]), "IC": [_g1(p) for p in vkj["IC"]]}


def proof_from_json(pj):
    return {"A": _g1(pj["pi_a"]), "B": _g2(pj["pi_b"]), "C": _g1(pj["pi_c"])}


def verify(vkey_json, public_json, proof_json):
    """The three documents as parsed JSON (dict, list, dict) -> True / False.  public.json may be `null` (no public
    inputs: what the reference writes then)."""
    vk = vkey_from_json(vkey_json)
    public = [int(x) for x in (public_json or [])]
    if len(public) + 1 != len(vk["IC"]):
        raise ValueError("public.json has %d signals, the key expects %d" % (len(public), len(vk["IC"]) - 1))
    proof = proof_from_json(proof_json)
    if any(v is None for v in proof.values()):
        return False
    return bool(pairing.groth16_verify(vk, proof, public))


def verify_files(vkey_path, public_path, proof_path):
    return verify(json.load(open(vkey_path)), json.load(open(public_path)), json.load(open(proof_path)))
