"""Synthetic Groth16 zkey / wtns generator with known toxic waste.

The reference ships no zkey/wtns fixtures, so every input is generated: a chain circuit
``(w_j + 3*w_{j+1}) * w_{j+1} = w_{j+2}`` over ``n_vars = n - 6`` wires with ``n_public = 4`` public
signals, the snarkjs-style binding rows, and all five point tables as *known* multiples of the
generators.  Because the discrete logs are known, every MSM result of the prover can be checked as one
scalar multiplication (``expected_*``) and a whole proof can be checked in the exponent, at any size.

Layout written by ``write_zkey`` / ``write_wtns`` is exactly what the reference reads
(src/binfile_utils.cpp:34-60, src/zkey_utils.cpp:17-52, src/wtns_utils.cpp:12-25, sections 1-9).

Scalar arithmetic is Python big integers; the point tables are produced by two callables
``g1_mul_many(list_of_ints) -> bytes`` / ``g2_mul_many`` supplied by the caller (the GPU fixed-base
kernel ``Context.fixed_base_g1/g2`` in bench.py, the CPU oracle in the no-GPU tests).
"""
import random
import struct

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT = 1 << 256
G1 = (1, 2)
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
       11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930,
       4082367875863433681332203403145435568316851327593401208105741076214120093531))


def fq_mont(x):
    return (x * MONT % Q).to_bytes(32, "little")


def g1_gen_bytes():
    return fq_mont(G1[0]) + fq_mont(G1[1])


def g2_gen_bytes():
    return b"".join(fq_mont(v) for v in (G2[0][0], G2[0][1], G2[1][0], G2[1][1]))


def root_of_unity(log_n):
    w = pow(5, (R - 1) >> 28, R)
    for _ in range(28 - log_n):
        w = w * w % R
    return w


def batch_inverse(vals):
    n = len(vals)
    pref = [1] * (n + 1)
    for i, v in enumerate(vals):
        pref[i + 1] = pref[i] * v % R
    inv = pow(pref[n], -1, R)
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = pref[i] * inv % R
        inv = inv * vals[i] % R
    return out


def le32_many(vals):
    if isinstance(vals, (bytes, bytearray, memoryview)):      # already packed (FastSynth)
        return bytes(vals)
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def count32(vals):
    """Number of 32-byte scalars in a list of integers or a packed byte string."""
    return len(vals) // 32 if isinstance(vals, (bytes, bytearray, memoryview)) else len(vals)


class Synth:
    """All the pieces of one synthetic circuit instance (see module docstring)."""

    def __init__(self, log_n, seed=1, n_public=4, witness="uniform"):
        rnd = random.Random(seed)
        n = 1 << log_n
        assert n >= 16
        self.log_n, self.n, self.n_public = log_n, n, n_public
        V = self.n_vars = n - 6
        P = n_public
        n_cons = self.n_cons = V - 2
        self.tau, self.alpha, self.beta, self.gamma, self.delta = (rnd.randrange(2, R) for _ in range(5))

        # ---- witness (normal form).  "uniform": chain values are pseudo-random full-width field elements.
        # "circomlike": most of the chain collapses to tiny values (w1 = 0 makes every wire 0, so instead
        # the chain is re-seeded: blocks of zeros / ones / small values are not expressible in this
        # circuit), so only "uniform" is a satisfiable witness; skewed scalars are exercised on the MSM
        # level by the tests and the microbench.
        w = [0] * V
        w[0] = 1
        w[1] = rnd.randrange(2, R)
        for j in range(n_cons):
            w[j + 2] = (w[j] + 3 * w[j + 1]) * w[j + 1] % R
        self.wtns = w

        # ---- section 4 records (m, c, s, value)
        recs = []
        for j in range(n_cons):
            recs.append((0, j, j, 1))
            recs.append((0, j, j + 1, 3))
            recs.append((1, j, j + 1, 1))
        for i in range(P + 1):
            recs.append((0, n_cons + i, i, 1))
        self.records = recs
        self.n_coefs = len(recs)

        # ---- Lagrange basis of the plain domain at tau
        tau, omega = self.tau, root_of_unity(log_n)
        zt = (pow(tau, n, R) - 1) % R
        wj, dens, ws = 1, [], []
        for j in range(n):
            ws.append(wj)
            dens.append(n * (tau - wj) % R)
            wj = wj * omega % R
        dinv = batch_inverse(dens)
        L = [zt * ws[j] % R * dinv[j] % R for j in range(n)]

        A = [0] * V
        B = [0] * V
        C = [0] * V
        for (m, c, s, v) in recs:
            if m == 0:
                A[s] = (A[s] + v * L[c]) % R
            else:
                B[s] = (B[s] + v * L[c]) % R
        for j in range(n_cons):
            C[j + 2] = (C[j + 2] + L[j]) % R
        self.A_tau, self.B_tau, self.C_tau = A, B, C
        dinv_delta, ginv = pow(self.delta, -1, R), pow(self.gamma, -1, R)
        K = [(self.beta * A[i] + self.alpha * B[i] + C[i]) % R for i in range(V)]
        self.K = K
        self.ic_scalars = [K[i] * ginv % R for i in range(P + 1)]
        self.c_scalars = [K[i] * dinv_delta % R for i in range(P + 1, V)]

        # ---- H table: Z(tau) * Lc_i(tau) / (-2 delta), Lagrange basis of the coset g*omega^i
        g = root_of_unity(log_n + 1)
        xs, x = [], g
        for i in range(n):
            xs.append(x)
            x = x * omega % R
        dens = [(-n) * (tau - xs[i]) % R for i in range(n)]
        dinv = batch_inverse(dens)
        tn1 = (pow(tau, n, R) + 1) % R
        hfac = zt * pow((-2 * self.delta) % R, -1, R) % R
        self.h_scalars_tbl = [tn1 * xs[i] % R * dinv[i] % R * hfac % R for i in range(n)]
        self._coset = (xs, tn1, dinv)
        self.points = None

    # ---- coefficient section bytes: u32 count + 44-byte records, value stored as v*R^2 (Appendix A)
    def coefs_section(self):
        out = [struct.pack("<I", self.n_coefs)]
        r2 = MONT * MONT % R
        cache = {}
        for (m, c, s, v) in self.records:
            if v not in cache:
                cache[v] = (v * r2 % R).to_bytes(32, "little")
            out.append(struct.pack("<III", m, c, s) + cache[v])
        return b"".join(out)

    def wtns_bytes(self):
        return le32_many(self.wtns)

    def build_points(self, g1_mul_many, g2_mul_many):
        """Fill the five tables + vk points from the known scalars."""
        V, P = self.n_vars, self.n_public
        self.points = {
            "A": g1_mul_many(self.A_tau),
            "B1": g1_mul_many(self.B_tau),
            "B2": g2_mul_many(self.B_tau),
            "C": g1_mul_many(self.c_scalars),
            "H": g1_mul_many(self.h_scalars_tbl),
            "IC": g1_mul_many(self.ic_scalars),
        }
        vk1 = g1_mul_many([self.alpha, self.beta, self.delta])
        vk2 = g2_mul_many([self.beta, self.gamma, self.delta])
        self.vk = {"alpha1": vk1[0:64], "beta1": vk1[64:128], "delta1": vk1[128:192],
                   "beta2": vk2[0:128], "gamma2": vk2[128:256], "delta2": vk2[256:384]}
        assert len(self.points["A"]) == 64 * V and len(self.points["C"]) == 64 * (V - P - 1)
        return self

    # known discrete logs used by the in-the-exponent proof check of bench.py
    @property
    def dlog_a(self):
        return sum(w * a for w, a in zip(self.wtns, self.A_tau)) % R

    @property
    def dlog_b(self):
        return sum(w * b for w, b in zip(self.wtns, self.B_tau)) % R

    @property
    def dlog_pub(self):
        return sum(self.wtns[i] * self.K[i] for i in range(self.n_public + 1)) % R

    # ---- expected discrete logs of the five pre-blinding MSM results (groth16.cpp:165-207)
    def expected_dlogs(self, h_scalars=None):
        w, V, P = self.wtns, self.n_vars, self.n_public
        ea = sum(w[i] * self.A_tau[i] for i in range(V)) % R
        eb = sum(w[i] * self.B_tau[i] for i in range(V)) % R
        ec = sum(w[i] * self.c_scalars[i - P - 1] for i in range(P + 1, V)) % R
        if h_scalars is None:
            h_scalars = self.h_evals()
        eh = sum(h * t for h, t in zip(h_scalars, self.h_scalars_tbl)) % R
        return {"pih": eh, "pi_a": ea, "pib1": eb, "pi_b": eb, "pi_c": ec}

    def h_evals(self):
        """(A*B - C)(x_i) on the coset, computed directly from the polynomials' values at tau-free data:
        A(X) = sum_c a_c L_c(X) with a_c = <row c of A, w>; evaluated at x_i through the barycentric form."""
        n, w = self.n, self.wtns
        a = [0] * n
        b = [0] * n
        for (m, c, s, v) in self.records:
            if m == 0:
                a[c] = (a[c] + v * w[s]) % R
            else:
                b[c] = (b[c] + v * w[s]) % R
        cvals = [a[i] * b[i] % R for i in range(n)]
        omega, g = root_of_unity(self.log_n), root_of_unity(self.log_n + 1)
        # barycentric evaluation at x = g*omega^i of the degree<n interpolant of vals on {omega^j}:
        #   f(x) = (x^n - 1)/n * sum_j vals_j * omega^j / (x - omega^j),  x^n - 1 = -2 on the coset
        # O(n^2): only used for small n in tests
        assert n <= 4096, "h_evals is quadratic; use it for small domains only"
        ws = [pow(omega, j, R) for j in range(n)]
        out = []
        ninv = pow(n, -1, R)
        for i in range(n):
            x = g * ws[i] % R
            inv = batch_inverse([(x - ws[j]) % R for j in range(n)])
            fa = sum(a[j] * ws[j] % R * inv[j] for j in range(n)) % R
            fb = sum(b[j] * ws[j] % R * inv[j] for j in range(n)) % R
            fc = sum(cvals[j] * ws[j] % R * inv[j] for j in range(n)) % R
            k = (-2) * ninv % R
            out.append(((fa * k) * (fb * k) - fc * k) % R)
        return out


class FastSynth:
    """The same circuit instance as Synth(log_n, seed, n_public) with the big vectors computed by the library's
    host routine b200_synth_chain (C++, seconds at 2^24 instead of minutes) and kept as packed little-endian bytes
    instead of lists of Python integers.  Interface used by bench.py / tools: n, n_vars, n_public, n_coefs, toxic
    waste, wtns_bytes(), coefs_section(), build_points(), points, vk, dlog_a / dlog_b / dlog_pub."""

    def __init__(self, log_n, seed=1, n_public=4):
        import ctypes
        from . import lib
        rnd = random.Random(seed)
        n = 1 << log_n
        assert n >= 16
        self.log_n, self.n, self.n_public = log_n, n, n_public
        V = self.n_vars = n - 6
        P = n_public
        self.n_cons = V - 2
        self.tau, self.alpha, self.beta, self.gamma, self.delta = (rnd.randrange(2, R) for _ in range(5))
        w1 = rnd.randrange(2, R)
        self.n_coefs = 3 * self.n_cons + P + 1
        b = lambda v: int(v).to_bytes(32, "little")
        mk = lambda k: ctypes.create_string_buffer(32 * k)
        wt, a, bb, cs, ic, h, dl = mk(V), mk(V), mk(V), mk(V - P - 1), mk(P + 1), mk(n), mk(3)
        rc = lib().b200_synth_chain(log_n, P, b(self.tau), b(self.alpha), b(self.beta), b(self.gamma), b(self.delta),
                                    b(w1), wt, a, bb, cs, ic, h, dl)
        if rc != 0:
            raise ValueError("b200_synth_chain failed (%d)" % rc)
        self._wtns = wt.raw
        self.A_tau, self.B_tau, self.c_scalars, self.ic_scalars, self.h_scalars_tbl = a.raw, bb.raw, cs.raw, ic.raw, h.raw
        self.dlog_a, self.dlog_b, self.dlog_pub = (int.from_bytes(dl.raw[32 * i:32 * i + 32], "little") for i in range(3))
        self.points = None

    def wtns_bytes(self):
        return self._wtns

    def coefs_section(self):
        return self.coefs_array().tobytes()

    def coefs_array(self):
        """zkey section 4 (u32 count, then the packed 44-byte records) as ONE numpy byte buffer filled in place and
        cached: at 2^26 the section is 8.9 GB and must not be copied around."""
        import numpy as np
        if getattr(self, "_coefs_buf", None) is not None:
            return self._coefs_buf
        nc, P = self.n_cons, self.n_public
        rec = np.dtype([("m", "<u4"), ("c", "<u4"), ("s", "<u4"), ("v", "V32")])
        buf = np.zeros(4 + self.n_coefs * 44, dtype=np.uint8)
        buf[:4] = np.frombuffer(struct.pack("<I", self.n_coefs & 0xffffffff), dtype=np.uint8)
        out = buf[4:].view(rec)
        r2 = MONT * MONT % R
        one, three = (np.frombuffer((v * r2 % R).to_bytes(32, "little"), dtype="V32")[0] for v in (1, 3))
        j = np.arange(nc, dtype=np.uint32)
        body = out[:3 * nc].reshape(nc, 3)
        body["c"] = j[:, None]
        body["m"][:, 2] = 1
        body["s"][:, 0] = j
        body["s"][:, 1] = j + 1
        body["s"][:, 2] = j + 1
        body["v"][:, 0] = one
        body["v"][:, 1] = three
        body["v"][:, 2] = one
        tail = out[3 * nc:]
        i = np.arange(P + 1, dtype=np.uint32)
        tail["c"] = nc + i
        tail["s"] = i
        tail["v"] = one
        self._coefs_buf = buf
        return buf

    build_points = Synth.build_points

    # ---- one shard's slices only (multi-GPU runs of big circuits: 2^26 tables are 24 GB, times eight ranks)
    def build_shard_points(self, g1_mul_many, g2_mul_many, index, count, bounds=None):
        """Point tables restricted to what b200_zkey_upload reads for shard `index` of `count` (same partition rule:
        [len * index / count, len * (index + 1) / count) of the witness-indexed tables and of H - or, with
        bounds = (lo_num, hi_num, den), [len * lo_num / den, len * hi_num / den)).  Fills
        self.shard_points = {name: (bytes, first_point_index)} and self.vk; see shard_table_address()."""
        V, P, n = self.n_vars, self.n_public, self.n
        if bounds:
            lo, hi = V * bounds[0] // bounds[2], V * bounds[1] // bounds[2]
            hlo, hhi = n * bounds[0] // bounds[2], n * bounds[1] // bounds[2]
        else:
            lo, hi = V * index // count, V * (index + 1) // count
            hlo, hhi = n * index // count, n * (index + 1) // count
        skip = P + 1
        clo, chi = max(lo, skip) - skip, max(hi, skip) - skip          # C table: signals skip .. V-1
        cut = lambda b, a, z: b[32 * a:32 * z]
        self.shard_points = {
            "A": (g1_mul_many(cut(self.A_tau, lo, hi)), lo),
            "B1": (g1_mul_many(cut(self.B_tau, lo, hi)), lo),
            "B2": (g2_mul_many(cut(self.B_tau, lo, hi)), lo),
            "C": (g1_mul_many(cut(self.c_scalars, clo, chi)), clo),
            "H": (g1_mul_many(cut(self.h_scalars_tbl, hlo, hhi)), hlo),
        }
        vk1 = g1_mul_many([self.alpha, self.beta, self.delta])
        vk2 = g2_mul_many([self.beta, self.gamma, self.delta])
        self.vk = {"alpha1": vk1[0:64], "beta1": vk1[64:128], "delta1": vk1[128:192],
                   "beta2": vk2[0:128], "gamma2": vk2[128:256], "delta2": vk2[256:384]}
        return self

    def shard_table_address(self, name):
        """Address to hand to b200_zkey_upload as the table's base: the library only reads this shard's range, which
        starts first_point_index points after the base, so base = address(slice) - first_point_index * point_size.
        (The buffer is kept alive in self._shard_keep.)"""
        import ctypes
        data, first = self.shard_points[name]
        size = 128 if name == "B2" else 64
        buf = ctypes.create_string_buffer(data, max(len(data), 1))
        self.__dict__.setdefault("_shard_keep", []).append(buf)
        return ctypes.addressof(buf) - first * size


# ----------------------------------------------------------------------------- iden3 binfile writers
def _binfile(magic, version, sections):
    out = [magic, struct.pack("<II", version, len(sections))]
    for (typ, payload) in sections:
        out.append(struct.pack("<IQ", typ, len(payload)))
        out.append(payload)
    return b"".join(out)


def zkey_bytes(s):
    """Sections 1-9 as the reference reads them; 10 (contributions) left empty."""
    assert s.points is not None, "call build_points first"
    hdr = struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
    hdr += struct.pack("<III", s.n_vars, s.n_public, s.n)
    hdr += s.vk["alpha1"] + s.vk["beta1"] + s.vk["beta2"] + s.vk["gamma2"] + s.vk["delta1"] + s.vk["delta2"]
    secs = [(1, struct.pack("<I", 1)), (2, hdr), (3, s.points["IC"]), (4, s.coefs_section()),
            (5, s.points["A"]), (6, s.points["B1"]), (7, s.points["B2"]), (8, s.points["C"]),
            (9, s.points["H"]), (10, b"")]
    return _binfile(b"zkey", 1, secs)


def wtns_bytes_file(s):
    hdr = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", s.n_vars)
    return _binfile(b"wtns", 2, [(1, hdr), (2, s.wtns_bytes())])
