"""Pairing verifier: moved into the package (rapidsnark_old_b200/verify/pairing.py); re-exported for the tests."""
from rapidsnark_old_b200.verify.pairing import *  # noqa: F401,F403
