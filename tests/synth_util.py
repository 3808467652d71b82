"""Test helpers: build synthetic circuits with the CPU oracle making the point tables (no GPU needed)."""
import functools

import bn254 as bn
import oracle_lib
from rapidsnark_old_b200 import synth


def oracle_point_makers(o=None):
    o = o or oracle_lib.best()
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()

    def g1_many(scalars):
        return b"".join(o.g1_mul_affine(g1, k) for k in scalars)

    def g2_many(scalars):
        return b"".join(o.g2_mul_affine(g2, k) for k in scalars)

    return g1_many, g2_many


@functools.lru_cache(maxsize=8)
def make(log_n, seed=1):
    s = synth.Synth(log_n, seed)
    s.build_points(*oracle_point_makers())
    return s


def expected_affine(o, s, h_scalars=None):
    """The five pre-blinding MSM results as affine bytes, from the known discrete logs."""
    d = s.expected_dlogs(h_scalars)
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
    return (o.g1_mul_affine(g1, d["pih"]), o.g1_mul_affine(g1, d["pi_a"]), o.g1_mul_affine(g1, d["pib1"]),
            o.g2_mul_affine(g2, d["pi_b"]), o.g1_mul_affine(g1, d["pi_c"]))
