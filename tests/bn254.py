"""BN254 big-integer helpers: moved into the package (rapidsnark_old_b200/verify/bn254.py); re-exported for the tests."""
from rapidsnark_old_b200.verify.bn254 import *  # noqa: F401,F403
from rapidsnark_old_b200.verify.bn254 import _add, _mul, _Fq, _Fq2  # noqa: F401
