"""Pure-Python big-integer BN254 helpers for the tests (tiny cases only).

Independent of both the oracle and the CUDA path: constants from SURVEY.md Appendix B
(reference: depends/ffiasm/c/alt_bn128.hpp:37-48, tasksfile.js:10-11), textbook affine
short-Weierstrass arithmetic.  Used to pin the oracle against the reference's golden
vectors and to build small inputs.
"""
import random

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT_R = 1 << 256

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
G2_B = (19485874751759354771024239261021720505790618469301721065564631296452457478373,
        266929791119991161246907387137283842545076965332900288569378510910307636690)


# ----------------------------------------------------------------------------- byte encodings
def to_mont(x, p=Q):
    return (x * MONT_R % p).to_bytes(32, "little")


def from_mont(b, p=Q):
    return int.from_bytes(b, "little") * pow(MONT_R, -1, p) % p


def le32(x):
    return int(x).to_bytes(32, "little")


def g1_aff_bytes(P):
    if P is None:
        return bytes(64)
    return to_mont(P[0]) + to_mont(P[1])


def g1_aff_from_bytes(b):
    if b == bytes(64):
        return None
    return (from_mont(b[:32]), from_mont(b[32:64]))


def g2_aff_bytes(P):
    if P is None:
        return bytes(128)
    (xa, xb), (ya, yb) = P
    return to_mont(xa) + to_mont(xb) + to_mont(ya) + to_mont(yb)


def g2_aff_from_bytes(b):
    if b == bytes(128):
        return None
    v = [from_mont(b[i * 32:(i + 1) * 32]) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


# ----------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


class _Fq:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % Q)
    sub = staticmethod(lambda a, b: (a - b) % Q)
    mul = staticmethod(lambda a, b: a * b % Q)
    neg = staticmethod(lambda a: (-a) % Q)
    inv = staticmethod(lambda a: pow(a, -1, Q))


class _Fq2:
    zero, one = (0, 0), (1, 0)
    add, sub, mul, neg, inv = map(staticmethod, (f2_add, f2_sub, f2_mul, f2_neg, f2_inv))


# ----------------------------------------------------------------------------- affine group law (None = infinity)
def _add(F, P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        x1x1 = F.mul(x1, x1)
        lam = F.mul(F.add(F.add(x1x1, x1x1), x1x1), F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def _mul(F, P, k):
    acc = None
    while k:
        if k & 1:
            acc = _add(F, acc, P)
        P = _add(F, P, P)
        k >>= 1
    return acc


def g1_add(P, S): return _add(_Fq, P, S)
def g1_mul(P, k): return _mul(_Fq, P, k)
def g1_neg(P): return None if P is None else (P[0], (-P[1]) % Q)
def g2_add(P, S): return _add(_Fq2, P, S)
def g2_mul(P, k): return _mul(_Fq2, P, k)
def g2_neg(P): return None if P is None else (P[0], f2_neg(P[1]))


def g1_on_curve(P):
    return P is None or (P[1] * P[1] - P[0] ** 3 - 3) % Q == 0


def g2_on_curve(P):
    if P is None:
        return True
    x, y = P
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), G2_B)) == (0, 0)


def g1_msm(points, scalars):
    acc = None
    for P, k in zip(points, scalars):
        acc = g1_add(acc, g1_mul(P, k))
    return acc


def g2_msm(points, scalars):
    acc = None
    for P, k in zip(points, scalars):
        acc = g2_add(acc, g2_mul(P, k))
    return acc


# ----------------------------------------------------------------------------- Fr roots of unity
def fr_root_of_unity(log_n):
    """Primitive 2^log_n-th root used by the reference (fft.cpp:52-83): 5^((r-1)/2^28) squared down."""
    w = pow(5, (R_ORDER - 1) >> 28, R_ORDER)
    for _ in range(28 - log_n):
        w = w * w % R_ORDER
    return w


def ntt_naive(a, inverse=False):
    """O(n^2) DFT over Fr with the reference's root, natural in -> natural out."""
    n = len(a)
    log_n = n.bit_length() - 1
    w = fr_root_of_unity(log_n)
    if inverse:
        w = pow(w, -1, R_ORDER)
    out = []
    for k in range(n):
        wk = pow(w, k, R_ORDER)
        acc, x = 0, 1
        for j in range(n):
            acc = (acc + a[j] * x) % R_ORDER
            x = x * wk % R_ORDER
        out.append(acc)
    if inverse:
        ninv = pow(n, -1, R_ORDER)
        out = [v * ninv % R_ORDER for v in out]
    return out


def rng(seed):
    return random.Random(seed)
