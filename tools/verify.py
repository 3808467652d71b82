#!/usr/bin/env python
"""Groth16 verification from the command line - the stand-in for `snarkjs groth16 verify` (reference README.md:44-53):

    python tools/verify.py verification_key.json public.json proof.json
    python tools/verify.py --zkey circuit.zkey public.json proof.json       (key exported from the zkey on the fly)

Prints "OK!" and exits 0 for a valid proof, "Invalid proof" and exits 1 otherwise (malformed input: exit 2).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main(argv):
    if len(argv) == 5 and argv[1] == "--zkey":
        import export_vkey
        vkj = export_vkey.export(open(argv[2], "rb").read())
        public_path, proof_path = argv[3], argv[4]
    elif len(argv) == 4:
        vkj = json.load(open(argv[1]))
        public_path, proof_path = argv[2], argv[3]
    else:
        print(__doc__, file=sys.stderr)
        return 2
    from rapidsnark_old_b200 import verify
    try:
        ok = verify.verify(vkj, json.load(open(public_path)), json.load(open(proof_path)))
    except (ValueError, KeyError, IndexError, TypeError) as e:
        print("verify: malformed input: %s" % e, file=sys.stderr)
        return 2
    print("OK!" if ok else "Invalid proof")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
