"""Pure-Python BN254 optimal-ate pairing and a Groth16 verifier (host-side tool: a proof is three points).

Replaces `snarkjs groth16 verify` (README.md:44-53 of the reference; tools/verify.py is the CLI): textbook construction,
Fq12 = Fq[w]/(w^12 - 18 w^6 + 82) with Fq2 = Fq[u]/(u^2+1) embedded through u = w^6 - 9, G2 points mapped to
E(Fq12) by the sextic untwist (x w^2, y w^3), Miller loop over 6t+2 = 29793968203157093288 plus the two
Frobenius lines, final exponentiation by (q^12-1)/r.  Independent of the oracle and of the CUDA path.
"""
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ATE = 29793968203157093288
LOG_ATE = 63


class F12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % Q for x in c]

    @staticmethod
    def one():
        return F12([1] + [0] * 11)

    @staticmethod
    def zero():
        return F12([0] * 12)

    def __add__(self, o): return F12([a + b for a, b in zip(self.c, o.c)])
    def __sub__(self, o): return F12([a - b for a, b in zip(self.c, o.c)])
    def __neg__(self): return F12([-a for a in self.c])
    def __eq__(self, o): return self.c == o.c

    def __mul__(self, o):
        if isinstance(o, int):
            return F12([a * o for a in self.c])
        t = [0] * 23
        a, b = self.c, o.c
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        for k in range(22, 11, -1):          # w^12 = 18 w^6 - 82
            v = t[k]
            if v:
                t[k - 6] += 18 * v
                t[k - 12] -= 82 * v
        return F12(t[:12])

    def __pow__(self, e):
        r, b = F12.one(), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def inv(self):
        # Fermat in the field of q^12 elements would be huge; use the extended Euclid on polynomials over Fq
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [82, 0, 0, 0, 0, 0, -18 % Q, 0, 0, 0, 0, 0, 1]

        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        def pdiv(a, b):
            da, db = deg(a), deg(b)
            t = list(a)
            o = [0] * len(a)
            binv = pow(b[db], -1, Q)
            for i in range(da - db, -1, -1):
                o[i] = t[db + i] * binv % Q
                for c in range(db + 1):
                    t[c + i] = (t[c + i] - o[i] * b[c]) % Q
            return o[:deg(o) + 1]
        while deg(low):
            r = pdiv(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                    new[i + j] = (new[i + j] - low[i] * r[j]) % Q
            lm, low, hm, high = nm, new, lm, low
        li = pow(low[0], -1, Q)
        return F12([x * li for x in lm[:12]])

    def is_zero(self):
        return not any(self.c)


W = F12([0, 1] + [0] * 10)
W2, W3 = W * W, W * W * W


def fq2_to_f12(a, b):
    """a + b u with u = w^6 - 9"""
    return F12([a - 9 * b, 0, 0, 0, 0, 0, b, 0, 0, 0, 0, 0])


def untwist(P2):
    (xa, xb), (ya, yb) = P2
    return (fq2_to_f12(xa, xb) * W2, fq2_to_f12(ya, yb) * W3)


def cast_g1(P1):
    return (F12([P1[0]] + [0] * 11), F12([P1[1]] + [0] * 11))


def _double(P):
    x, y = P
    lam = (x * x * 3) * (y * 2).inv()
    nx = lam * lam - x * 2
    return (nx, lam * (x - nx) - y)


def _add(P, S):
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        return _double(P) if y1 == y2 else None
    lam = (y2 - y1) * (x2 - x1).inv()
    nx = lam * lam - x1 - x2
    return (nx, lam * (x1 - nx) - y1)


def _line(P1, P2, T):
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if not (x1 == x2):
        m = (y2 - y1) * (x2 - x1).inv()
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1 * 3) * (y1 * 2).inv()
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller(Q2, P1):
    """Miller loop value (before the final exponentiation) for Q2 in G2 (affine Fq2 pair), P1 in G1."""
    if Q2 is None or P1 is None:
        return F12.one()
    Qt, P = untwist(Q2), cast_g1(P1)
    Rp, f = Qt, F12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _line(Rp, Rp, P)
        Rp = _double(Rp)
        if ATE & (1 << i):
            f = f * _line(Rp, Qt, P)
            Rp = _add(Rp, Qt)
    Q1 = (Qt[0] ** Q, Qt[1] ** Q)
    nQ2 = (Q1[0] ** Q, -(Q1[1] ** Q))
    f = f * _line(Rp, Q1, P)
    Rp = _add(Rp, Q1)
    f = f * _line(Rp, nQ2, P)
    return f


def final_exp(f):
    return f ** ((Q ** 12 - 1) // R)


def pairing_product_is_one(pairs):
    """prod e(P1_i, Q2_i) == 1 with a single final exponentiation."""
    f = F12.one()
    for P1, Q2 in pairs:
        f = f * miller(Q2, P1)
    return final_exp(f) == F12.one()


def g1_add(P, S):
    if P is None: return S
    if S is None: return P
    if P[0] == S[0]:
        if (P[1] + S[1]) % Q == 0: return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, Q) % Q
    else:
        lam = (S[1] - P[1]) * pow(S[0] - P[0], -1, Q) % Q
    x = (lam * lam - P[0] - S[0]) % Q
    return (x, (lam * (P[0] - x) - P[1]) % Q)


def g1_mul(P, k):
    acc = None
    while k:
        if k & 1: acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


def groth16_verify(vk, proof, public):
    """vk: alpha1 (G1), beta2, gamma2, delta2 (G2), IC (list of G1); proof: A (G1), B (G2), C (G1); public: ints.
    e(A,B) = e(alpha1,beta2) e(vk_x,gamma2) e(C,delta2)   <=>   e(-A,B) e(alpha1,beta2) e(vk_x,gamma2) e(C,delta2) = 1"""
    vk_x = vk["IC"][0]
    for w, ic in zip(public, vk["IC"][1:]):
        vk_x = g1_add(vk_x, g1_mul(ic, w % R))
    A = proof["A"]
    negA = (A[0], (-A[1]) % Q)
    return pairing_product_is_one([(negA, proof["B"]), (vk["alpha1"], vk["beta2"]), (vk_x, vk["gamma2"]),
                                   (proof["C"], vk["delta2"])])
