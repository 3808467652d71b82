"""rapidsnark_old_b200 - B200-native Groth16 (BN254) hot path behind a C-ABI.

This module is only the ctypes binding of ``include/b200snark.h`` (``libb200snark.so``, built from
``csrc/`` by ``__graft_entry__.build()``).  The product is the CUDA library; there is no Python or CPU
fallback: if the shared object is missing, or no CUDA device is present, the calls raise.

Interface mirrored (reference iden3/rapidsnark-old):
  Context.msm_g1 / msm_g2   <- Curve::multiMulByScalar      depends/ffiasm/c/curve.hpp:118-121
  Context.ntt               <- FFT<Fr>::fft / ifft          depends/ffiasm/c/fft.hpp:24-25
  Context.zkey_upload       <- Groth16::makeProver          src/groth16.cpp:9-46
  ZKey.prove_msms           <- Prover::prove, pre-blinding  src/groth16.cpp:48-207
"""
import ctypes
import os
import weakref

# more hardware work queues than the default 8, so that the library's streams (main, H, one per in-flight MSM) and
# torch's / NCCL's never share one and serialise; only effective before the process's first CUDA call
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libb200snark.so")   # B200_LIB: A/B builds of the same ABI

OK, ERR_ARG, ERR_CUDA, ERR_NO_GPU, ERR_RANGE = 0, 1, 2, 3, 4

_u32, _u64, _vp, _int = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int

EXPORTS = [
    "b200_init", "b200_free", "b200_last_error", "b200_launch_count", "b200_last_phase_ms", "b200_phase_name",
    "b200_last_timeline",
    "b200_msm_g1", "b200_msm_g2", "b200_msm_g1_dev", "b200_msm_g2_dev", "b200_set_msm_window", "b200_set_option",
    "b200_ntt_fr", "b200_ntt_fr_dev",
    "b200_zkey_upload", "b200_zkey_free", "b200_zkey_share", "b200_h_scalars", "b200_prove_msms", "b200_prove_msms_dev", "b200_stream",
    "b200_prove_begin", "b200_prove_finish", "b200_exchange_polys", "b200_groth16_prove",
    "b200_groth16_finalize", "b200_groth16_blind_prepare", "b200_groth16_finalize_prepared", "b200_host_fold_partials",
    "b200_fq_to_decimal",
    "b200_fixed_base_g1", "b200_fixed_base_g2", "b200_synth_chain",
    "b200_host_fq_mul", "b200_host_fq_add", "b200_host_fq_sub", "b200_host_fq_neg", "b200_host_fq_inv",
    "b200_host_fr_mul", "b200_host_fr_add", "b200_host_fr_sub", "b200_host_fr_neg", "b200_host_fr_inv",
    "b200_host_fq2_mul", "b200_host_fq2_sqr",
    "b200_host_g1_add", "b200_host_g1_madd", "b200_host_g1_dbl", "b200_host_g1_neg", "b200_host_g1_to_affine",
    "b200_host_g1_mul",
    "b200_host_g2_add", "b200_host_g2_madd", "b200_host_g2_dbl", "b200_host_g2_neg", "b200_host_g2_to_affine",
    "b200_host_g2_mul",
]


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200snark error %d: %s" % (code, msg))
        self.code = code


class VKey(ctypes.Structure):
    _fields_ = [("alpha1", _vp), ("beta1", _vp), ("beta2", _vp), ("delta1", _vp), ("delta2", _vp)]


class ZKeyDesc(ctypes.Structure):
    _fields_ = [("n_vars", _u32), ("n_public", _u32), ("domain_size", _u32), ("n_coefs", _u64),
                ("coefs", _vp), ("points_a", _vp), ("points_b1", _vp), ("points_b2", _vp),
                ("points_c", _vp), ("points_h", _vp), ("shard_index", _u32), ("shard_count", _u32),
                ("shard_lo_num", _u32), ("shard_hi_num", _u32), ("shard_den", _u32)]


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built - there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C rapidsnark_old_b200/csrc)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_error.argtypes = [_vp]
        L.b200_phase_name.restype = ctypes.c_char_p
        L.b200_launch_count.restype = _u64
        L.b200_launch_count.argtypes = [_vp]
        L.b200_init.argtypes = [_int, ctypes.POINTER(_vp)]
        L.b200_free.argtypes = [_vp]
        for name in ("b200_msm_g1", "b200_msm_g2", "b200_msm_g1_dev", "b200_msm_g2_dev"):
            getattr(L, name).argtypes = [_vp, _vp, _vp, _u32, _u64, _vp]
        L.b200_set_msm_window.argtypes = [_vp, _int]
        L.b200_set_option.argtypes = [_vp, ctypes.c_char_p, _int]
        L.b200_ntt_fr.argtypes = [_vp, _vp, _u64, _int]
        L.b200_ntt_fr_dev.argtypes = [_vp, _vp, _u64, _int]
        L.b200_zkey_upload.argtypes = [_vp, ctypes.POINTER(ZKeyDesc), ctypes.POINTER(_vp)]
        L.b200_zkey_free.argtypes = [_vp]
        L.b200_zkey_share.argtypes = [_vp, _vp, ctypes.POINTER(_vp)]
        L.b200_h_scalars.argtypes = [_vp, _vp, _vp, _vp]
        L.b200_prove_msms.argtypes = [_vp, _vp, _vp, _vp]
        L.b200_prove_msms_dev.argtypes = [_vp, _vp, _vp, _vp]
        L.b200_prove_begin.argtypes = [_vp, _vp, _vp, _int, _u32, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]
        L.b200_prove_finish.argtypes = [_vp, _vp, _vp]
        L.b200_groth16_prove.argtypes = [_vp, _vp, _vp, _int, ctypes.POINTER(VKey), _vp, _vp, _vp, _vp]
        L.b200_exchange_polys.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _int]
        L.b200_stream.restype = _vp
        L.b200_stream.argtypes = [_vp]
        L.b200_groth16_finalize.argtypes = [_vp] * 9
        L.b200_groth16_blind_prepare.argtypes = [_vp] * 5
        L.b200_groth16_finalize_prepared.argtypes = [_vp] * 8
        L.b200_host_fold_partials.argtypes = [_vp, _int, _vp]
        L.b200_fq_to_decimal.argtypes = [_vp, _vp]
        L.b200_fixed_base_g1.argtypes = [_vp, _vp, _vp, _u64, _vp]
        L.b200_fixed_base_g2.argtypes = [_vp, _vp, _vp, _u64, _vp]
        L.b200_synth_chain.argtypes = [_u32, _u32] + [_vp] * 13
        L.b200_last_phase_ms.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), _int]
        L.b200_last_timeline.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), _int]
        _lib = L
    return _lib


def _ptr(buf):
    """bytes / bytearray / ctypes buffer / numpy array / int address -> void*"""
    if isinstance(buf, int):
        return _vp(buf)
    if isinstance(buf, bytes):
        return ctypes.cast(ctypes.c_char_p(buf), _vp)
    if hasattr(buf, "ctypes"):
        return _vp(buf.ctypes.data)
    if isinstance(buf, bytearray):
        return ctypes.cast((ctypes.c_char * len(buf)).from_buffer(buf), _vp)
    return ctypes.cast(buf, _vp)


# ----------------------------------------------------------------------------- host helpers (CPU)
def _host_call(name, out_size, *args):
    out = ctypes.create_string_buffer(out_size)
    getattr(lib(), name)(out, *args)
    return out.raw


def host_g1_add(a, b): return _host_call("b200_host_g1_add", 128, a, b)
def host_g2_add(a, b): return _host_call("b200_host_g2_add", 256, a, b)
def host_g1_to_affine(a): return _host_call("b200_host_g1_to_affine", 64, a)
def host_g2_to_affine(a): return _host_call("b200_host_g2_to_affine", 128, a)
def host_g1_mul(base_affine, scalar): return _host_call("b200_host_g1_mul", 128, base_affine, scalar, _u32(len(scalar)))
def host_g2_mul(base_affine, scalar): return _host_call("b200_host_g2_mul", 256, base_affine, scalar, _u32(len(scalar)))


def groth16_finalize(msms768, vk, r32, s32):
    """Blinding + to-affine on the host (groth16.cpp:209-253): -> A(64) | B(128) | C(64) affine Montgomery."""
    out = ctypes.create_string_buffer(256)
    lib().b200_groth16_finalize(_ptr(msms768), _ptr(vk["alpha1"]), _ptr(vk["beta1"]), _ptr(vk["beta2"]),
                                _ptr(vk["delta1"]), _ptr(vk["delta2"]), _ptr(r32), _ptr(s32), out)
    return out.raw


def fq_to_decimal(mont32):
    buf = ctypes.create_string_buffer(80)
    lib().b200_fq_to_decimal(_ptr(mont32), buf)
    return buf.value.decode()


def proof_json(proof256):
    """proof.json text exactly as the reference's `proofFile << proof->toJson()` (groth16.cpp:268-301)."""
    d = [fq_to_decimal(proof256[i * 32:(i + 1) * 32]) for i in range(8)]
    return ('{"pi_a":["%s","%s","1"],"pi_b":[["%s","%s"],["%s","%s"],["1","0"]],"pi_c":["%s","%s","1"],'
            '"protocol":"groth16"}' % (d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]))


def exchange_polys(zkeys):
    """In-process multi-GPU exchange (b200_exchange_polys): zkeys[g] = shard g of len(zkeys), each after
    prove_begin(..., poly_mask = polynomials i with i % n == g)."""
    n = len(zkeys)
    ctxs = (_vp * n)(*[z.ctx.handle for z in zkeys])
    zks = (_vp * n)(*[z.handle for z in zkeys])
    rc = lib().b200_exchange_polys(ctxs, zks, n)
    if rc != OK:
        raise B200Error(rc, "b200_exchange_polys failed")


def fold_partials(parts768):
    """Sum per-GPU partial results of prove_msms (pih, pi_a, pib1 | pi_b | pi_c) into one 768-byte record."""
    out = ctypes.create_string_buffer(768)
    lib().b200_host_fold_partials(b"".join(bytes(p) for p in parts768), len(parts768), out)
    return out.raw


def groth16_blind_prepare(vk, r32, s32):
    """The part of the blinding that needs only the key and r, s (r*delta1, s*delta1, rs*delta1, s*delta2): run it on
    a host thread while the GPU works (ctypes releases the GIL), then groth16_finalize_prepared."""
    out = ctypes.create_string_buffer(640)
    lib().b200_groth16_blind_prepare(_ptr(vk["delta1"]), _ptr(vk["delta2"]), _ptr(r32), _ptr(s32), out)
    return out.raw


def groth16_finalize_prepared(msms768, vk, prep640, r32, s32):
    out = ctypes.create_string_buffer(256)
    lib().b200_groth16_finalize_prepared(_ptr(msms768), _ptr(vk["alpha1"]), _ptr(vk["beta1"]), _ptr(vk["beta2"]),
                                         _ptr(prep640), _ptr(r32), _ptr(s32), out)
    return out.raw


# ----------------------------------------------------------------------------- device context
class ZKey:
    def __init__(self, ctx, handle, desc_keepalive):
        self.ctx, self.handle, self._keep = ctx, handle, desc_keepalive
        self.domain_size = desc_keepalive[0].domain_size
        ctx._zkeys.add(self)      # b200_zkey_free needs the context alive: Context.close() frees its zkeys first

    def h_scalars(self, wtns):
        out = ctypes.create_string_buffer(self.domain_size * 32)
        self.ctx._check(lib().b200_h_scalars(self.ctx.handle, self.handle, _ptr(wtns), out))
        return out.raw

    def prove_msms(self, wtns):
        out = ctypes.create_string_buffer(768)
        self.ctx._check(lib().b200_prove_msms(self.ctx.handle, self.handle, _ptr(wtns), out))
        return out.raw

    def prove_msms_dev(self, d_wtns):
        out = ctypes.create_string_buffer(768)
        self.ctx._check(lib().b200_prove_msms_dev(self.ctx.handle, self.handle, _vp(d_wtns), out))
        return out.raw

    def prove(self, wtns, vk, r32, s32, on_device=False):
        """The whole proof in one call (b200_groth16_prove): -> (msms768, proof256 = A | B | C affine Montgomery).
        vk: dict with alpha1, beta1, beta2, delta1, delta2 (bytes)."""
        key = VKey(_ptr(vk["alpha1"]), _ptr(vk["beta1"]), _ptr(vk["beta2"]), _ptr(vk["delta1"]), _ptr(vk["delta2"]))
        proof, msms = ctypes.create_string_buffer(256), ctypes.create_string_buffer(768)
        w = _vp(wtns) if on_device else _ptr(wtns)
        self.ctx._check(lib().b200_groth16_prove(self.ctx.handle, self.handle, w, 1 if on_device else 0, ctypes.byref(key),
                                                 _ptr(r32), _ptr(s32), proof, msms))
        return msms.raw, proof.raw

    def prove_begin(self, wtns, on_device=False, poly_mask=7):
        """Stage 1 of the two-stage prove (multi-GPU): -> ([d_a, d_b, d_c] device addresses, H-stream handle).
        Returns without synchronising; exchange the buffers on that stream, then call prove_finish()."""
        bufs = (_vp * 3)()
        hs = _vp()
        w = _vp(wtns) if on_device else _ptr(wtns)
        self.ctx._check(lib().b200_prove_begin(self.ctx.handle, self.handle, w, 1 if on_device else 0, _u32(poly_mask),
                                               bufs, ctypes.byref(hs)))
        return [int(b or 0) for b in bufs], int(hs.value or 0)

    def prove_finish(self):
        out = ctypes.create_string_buffer(768)
        self.ctx._check(lib().b200_prove_finish(self.ctx.handle, self.handle, out))
        return out.raw

    def free(self):
        if self.handle:
            lib().b200_zkey_free(self.handle)
            self.handle = None
            self.ctx._zkeys.discard(self)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU, one stream (b200_ctx).  Raises B200Error(ERR_NO_GPU) when there is no CUDA device."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib().b200_init(device, ctypes.byref(h))
        if rc != OK:
            raise B200Error(rc, lib().b200_last_error(None).decode())
        self.handle = h
        self._zkeys = weakref.WeakSet()

    def _check(self, rc):
        if rc != OK:
            raise B200Error(rc, lib().b200_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            for zk in list(self._zkeys):      # a zkey dereferences its context when freed (device, stream)
                zk.free()
            lib().b200_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- MSM
    def msm_g1(self, bases, scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(128)
        self._check(lib().b200_msm_g1(self.handle, _ptr(bases), _ptr(scalars), scalar_size, n, out))
        return out.raw

    def msm_g2(self, bases, scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(256)
        self._check(lib().b200_msm_g2(self.handle, _ptr(bases), _ptr(scalars), scalar_size, n, out))
        return out.raw

    def msm_g1_dev(self, d_bases, d_scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(128)
        self._check(lib().b200_msm_g1_dev(self.handle, _vp(d_bases), _vp(d_scalars), scalar_size, n, out))
        return out.raw

    def msm_g2_dev(self, d_bases, d_scalars, n, scalar_size=32):
        out = ctypes.create_string_buffer(256)
        self._check(lib().b200_msm_g2_dev(self.handle, _vp(d_bases), _vp(d_scalars), scalar_size, n, out))
        return out.raw

    def set_msm_window(self, c_bits):
        lib().b200_set_msm_window(self.handle, c_bits)

    def set_option(self, name, value):
        self._check(lib().b200_set_option(self.handle, name.encode(), int(value)))

    # ---- NTT (natural order in/out, Montgomery data)
    def ntt(self, data, inverse=False):
        buf = ctypes.create_string_buffer(bytes(data), len(data))
        self._check(lib().b200_ntt_fr(self.handle, buf, len(data) // 32, 1 if inverse else 0))
        return buf.raw

    def ntt_dev(self, d_ptr, n, inverse=False):
        self._check(lib().b200_ntt_fr_dev(self.handle, _vp(d_ptr), n, 1 if inverse else 0))

    # ---- zkey residency
    def zkey_upload(self, n_vars, n_public, domain_size, n_coefs, coefs_section, points_a, points_b1, points_b2,
                    points_c, points_h, shard_index=0, shard_count=1, shard_bounds=None):
        """shard_bounds = (lo_num, hi_num, den): uneven split, this shard owns [len*lo_num/den, len*hi_num/den)."""
        lo_num, hi_num, den = shard_bounds if shard_bounds else (0, 0, 0)
        keep = [coefs_section, points_a, points_b1, points_b2, points_c, points_h]
        d = ZKeyDesc(n_vars, n_public, domain_size, n_coefs, _ptr(coefs_section), _ptr(points_a), _ptr(points_b1),
                     _ptr(points_b2), _ptr(points_c), _ptr(points_h), shard_index, shard_count, lo_num, hi_num, den)
        h = _vp()
        self._check(lib().b200_zkey_upload(self.handle, ctypes.byref(d), ctypes.byref(h)))
        return ZKey(self, h, (d, keep))

    def zkey_share(self, zk):
        """A view of `zk`'s resident tables for THIS context (same device): concurrent proofs against one zkey."""
        h = _vp()
        self._check(lib().b200_zkey_share(self.handle, zk.handle, ctypes.byref(h)))
        return ZKey(self, h, zk._keep)

    # ---- synthetic tables
    def fixed_base_g1(self, base_affine, scalars32, n):
        out = ctypes.create_string_buffer(64 * max(n, 1))
        self._check(lib().b200_fixed_base_g1(self.handle, _ptr(base_affine), _ptr(scalars32), n, out))
        return out.raw[:64 * n]

    def fixed_base_g2(self, base_affine, scalars32, n):
        out = ctypes.create_string_buffer(128 * max(n, 1))
        self._check(lib().b200_fixed_base_g2(self.handle, _ptr(base_affine), _ptr(scalars32), n, out))
        return out.raw[:128 * n]

    def stream(self):
        """cudaStream_t (as int) all kernels of this context are launched on."""
        return int(lib().b200_stream(self.handle) or 0)

    # ---- instrumentation
    def launch_count(self):
        return int(lib().b200_launch_count(self.handle))

    def timeline(self):
        """[(phase name, start ms, end ms)] of the last call's timed segments (set_option("timeline", 1) first)."""
        arr = (ctypes.c_float * 768)()
        k = lib().b200_last_timeline(self.handle, arr, 768)
        return [(lib().b200_phase_name(int(arr[3 * i])).decode(), float(arr[3 * i + 1]), float(arr[3 * i + 2])) for i in range(k)]

    def phase_ms(self):
        arr = (ctypes.c_float * 16)()
        k = lib().b200_last_phase_ms(self.handle, arr, 16)
        return {lib().b200_phase_name(i).decode(): float(arr[i]) for i in range(k)}
