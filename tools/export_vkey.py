#!/usr/bin/env python
"""zkey -> verification_key.json in the snarkjs layout (protocol, curve, nPublic, vk_alpha_1, vk_beta_2, vk_gamma_2,
vk_delta_2, IC), so that `snarkjs groth16 verify verification_key.json public.json proof.json` can be run on a machine
that has node (this image does not; tests/pairing.py is the in-repo stand-in).  zkey sections 2 and 3, points are
stored affine in Montgomery form.
    python tools/export_vkey.py circuit.zkey verification_key.json
"""
import json
import sys

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
RINV = pow(1 << 256, -1, Q)


def sections(raw):
    n = int.from_bytes(raw[8:12], "little")
    pos, out = 12, {}
    for _ in range(n):
        typ = int.from_bytes(raw[pos:pos + 4], "little")
        size = int.from_bytes(raw[pos + 4:pos + 12], "little")
        out.setdefault(typ, raw[pos + 12:pos + 12 + size])
        pos += 12 + size
    return out


def fq(b):
    return str(int.from_bytes(b, "little") * RINV % Q)


def g1(b):
    return [fq(b[0:32]), fq(b[32:64]), "1"] if any(b) else ["0", "1", "0"]


def g2(b):
    return [[fq(b[0:32]), fq(b[32:64])], [fq(b[64:96]), fq(b[96:128])], ["1", "0"]] if any(b) else [["0", "0"], ["1", "0"], ["0", "0"]]


def export(zkey_bytes):
    s = sections(zkey_bytes)
    assert zkey_bytes[:4] == b"zkey" and int.from_bytes(s[1][:4], "little") == 1, "not a groth16 zkey"
    h = s[2]
    n_public = int.from_bytes(h[76:80], "little")
    return {"protocol": "groth16", "curve": "bn128", "nPublic": n_public,
            "vk_alpha_1": g1(h[84:148]), "vk_beta_2": g2(h[212:340]), "vk_gamma_2": g2(h[340:468]),
            "vk_delta_2": g2(h[532:660]), "IC": [g1(s[3][64 * i:64 * i + 64]) for i in range(n_public + 1)]}


if __name__ == "__main__":
    vk = export(open(sys.argv[1], "rb").read())
    json.dump(vk, open(sys.argv[2], "w"), indent=1)
