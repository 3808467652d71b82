"""ctypes loaders for the two oracle implementations (test infrastructure only).

  port : oracle/port/liboracle_port.so   plain-C restatement, symbols orc_*
  ref  : oracle/_ref/libref_oracle.so    the reference's own templates, symbols ref_*
Both export the API of oracle/oracle_api.h.  `Oracle` wraps either behind bytes-in/bytes-out calls.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "port", "liboracle_port.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")

u32, u64, vp = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p


def build_port():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])


class Oracle:
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.kind = "reference" if prefix == "ref_" else "port"

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def threads(self):
        return self.fn("threads")()

    # ---- points
    def g1_msm(self, bases, scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(128)
        self.fn("g1_msm")(bases, scalars, u32(scalar_size), u32(n), out)
        return out.raw

    def g2_msm(self, bases, scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(256)
        self.fn("g2_msm")(bases, scalars, u32(scalar_size), u32(n), out)
        return out.raw

    def g1_to_affine(self, xyzz):
        out = ctypes.create_string_buffer(64)
        self.fn("g1_to_affine")(xyzz, out)
        return out.raw

    def g2_to_affine(self, xyzz):
        out = ctypes.create_string_buffer(128)
        self.fn("g2_to_affine")(xyzz, out)
        return out.raw

    def _bin(self, name, size, a, b):
        out = ctypes.create_string_buffer(size)
        self.fn(name)(out, a, b)
        return out.raw

    def g1_add(self, a, b): return self._bin("g1_add", 128, a, b)
    def g1_madd(self, a, b): return self._bin("g1_madd", 128, a, b)
    def g2_add(self, a, b): return self._bin("g2_add", 256, a, b)
    def g2_madd(self, a, b): return self._bin("g2_madd", 256, a, b)
    def fq2_mul(self, a, b): return self._bin("fq2_mul", 64, a, b)

    def g1_dbl(self, a):
        out = ctypes.create_string_buffer(128)
        self.fn("g1_dbl")(out, a)
        return out.raw

    def g2_dbl(self, a):
        out = ctypes.create_string_buffer(256)
        self.fn("g2_dbl")(out, a)
        return out.raw

    def g1_mul(self, base_affine, scalar):
        out = ctypes.create_string_buffer(128)
        self.fn("g1_mul")(out, base_affine, scalar, u32(len(scalar)))
        return out.raw

    def g2_mul(self, base_affine, scalar):
        out = ctypes.create_string_buffer(256)
        self.fn("g2_mul")(out, base_affine, scalar, u32(len(scalar)))
        return out.raw

    def g1_mul_affine(self, base_affine, k):
        return self.g1_to_affine(self.g1_mul(base_affine, int(k).to_bytes(32, "little")))

    def g2_mul_affine(self, base_affine, k):
        return self.g2_to_affine(self.g2_mul(base_affine, int(k).to_bytes(32, "little")))

    # ---- NTT
    def fr_fft(self, data):
        buf = ctypes.create_string_buffer(data, len(data))
        self.fn("fr_fft")(buf, u64(len(data) // 32))
        return buf.raw

    def fr_ifft(self, data):
        buf = ctypes.create_string_buffer(data, len(data))
        self.fn("fr_ifft")(buf, u64(len(data) // 32))
        return buf.raw

    def fr_root(self, domain_pow, idx):
        out = ctypes.create_string_buffer(32)
        self.fn("fr_root")(u32(domain_pow), u64(idx), out)
        return out.raw

    # ---- prover phases
    def h_scalars(self, domain_size, n_coefs, coefs_section, wtns):
        out = ctypes.create_string_buffer(domain_size * 32)
        self.fn("h_scalars")(u32(domain_size), u64(n_coefs), coefs_section, wtns, out)
        return out.raw

    def prove_msms(self, n_vars, n_public, domain_size, n_coefs, coefs_section, pA, pB1, pB2, pC, pH, wtns):
        out = ctypes.create_string_buffer(768)
        self.fn("prove_msms")(u32(n_vars), u32(n_public), u32(domain_size), u64(n_coefs), coefs_section,
                              pA, pB1, pB2, pC, pH, wtns, out)
        return out.raw

    def blind(self, msms768, alpha1, beta1, beta2, delta1, delta2, r32, s32):
        out = ctypes.create_string_buffer(256)
        self.fn("blind")(msms768, alpha1, beta1, beta2, delta1, delta2, r32, s32, out)
        return out.raw

    def msms_to_affine(self, msms768):
        """pih, pi_a, pib1 (G1), pi_b (G2), pi_c (G1) -> canonical affine bytes (64,64,64,128,64)."""
        m = msms768
        return (self.g1_to_affine(m[0:128]), self.g1_to_affine(m[128:256]), self.g1_to_affine(m[256:384]),
                self.g2_to_affine(m[384:640]), self.g1_to_affine(m[640:768]))


_cache = {}


def port():
    if "port" not in _cache:
        if not os.path.exists(PORT_SO):
            build_port()
        _cache["port"] = Oracle(PORT_SO, "orc_")
    return _cache["port"]


def ref():
    """The compiled-from-reference oracle, or None when it has not been built (no /root/reference)."""
    if "ref" not in _cache:
        _cache["ref"] = Oracle(REF_SO, "ref_") if os.path.exists(REF_SO) else None
    return _cache["ref"]


def best():
    """Prefer the real reference when its prebuilt library is present."""
    return ref() or port()
