"""BN254 optimal-ate pairing in Python ints (oracle; test infrastructure only -- slow by design).

Restates what the reference's `verify` subcommand reaches through halo2-axiom's
`VerifierSHPLONK` -> `pairing::MultiMillerLoop` over halo2curves `bn256` [UPSTREAM, un-vendored;
SURVEY.md §3.4, §8(f) rank 1]: the final KZG check e(lhs, [1]_2) = e(W', [tau]_2).  The
construction is the textbook one (Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), G2 points mapped through the
sextic twist into Fq12, affine Miller loop with loop count 6t+2, plain final exponentiation); it is
pinned by bilinearity / non-degeneracy properties in tests/test_oracle_pairing.py, and it makes
oracle/verifier.py a verifier that needs no trapdoor.
"""
from .field import P_MOD, R_MOD

ATE_LOOP_COUNT = 29793968203157093288          # 6t + 2, t = 4965661367192848881
LOG_ATE_LOOP_COUNT = 63
FQ12_MOD_COEFFS = (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0)    # w^12 = 18 w^6 - 82

# G2 generator (x = x0 + x1 u, y = y0 + y1 u over Fq2 = Fq[u]/(u^2 + 1)); curve y^2 = x^3 + 3/(9 + u)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def _deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b):
    dega, degb = _deg(a), _deg(b)
    temp = list(a)
    o = [0] * len(a)
    inv_lead = pow(b[degb], -1, P_MOD)
    for i in range(dega - degb, -1, -1):
        q = temp[degb + i] * inv_lead % P_MOD
        o[i] = (o[i] + q) % P_MOD
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - b[c] * q) % P_MOD
    return o[:_deg(o) + 1]


class FQ12:
    """Element of Fq[w] / (w^12 - 18 w^6 + 82), coefficients low degree first."""
    __slots__ = ("c",)

    def __init__(self, coeffs):
        assert len(coeffs) == 12
        self.c = [x % P_MOD for x in coeffs]

    @staticmethod
    def one():
        return FQ12([1] + [0] * 11)

    @staticmethod
    def zero():
        return FQ12([0] * 12)

    @staticmethod
    def scalar(x):
        return FQ12([x] + [0] * 11)

    def __eq__(self, o):
        return self.c == o.c

    def __add__(self, o):
        return FQ12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return FQ12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return FQ12([-a for a in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return FQ12([a * o for a in self.c])
        b = [0] * 23
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for exp in range(22, 11, -1):          # reduce with w^12 = 18 w^6 - 82
            top = b[exp] % P_MOD
            if top:
                b[exp - 12] -= top * 82
                b[exp - 6] += top * 18
            b[exp] = 0
        return FQ12(b[:12])

    def inv(self):
        """Extended Euclid on polynomials over Fq."""
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [x % P_MOD for x in FQ12_MOD_COEFFS] + [1]
        while _deg(low):
            r = _poly_rounded_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % P_MOD for x in nm]
            new = [x % P_MOD for x in new]
            lm, low, hm, high = nm, new, lm, low
        li = pow(low[0], -1, P_MOD)
        return FQ12([x * li for x in lm[:12]])

    def __truediv__(self, o):
        return self * o.inv()

    def __pow__(self, e):
        acc, base = FQ12.one(), self
        while e:
            if e & 1:
                acc = acc * base
            base = base * base
            e >>= 1
        return acc

    def is_zero(self):
        return not any(self.c)


# ---- Fq2 and G2 (affine, identity = None) -----------------------------------------------------------------
def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P_MOD, (a[0] * b[1] + a[1] * b[0]) % P_MOD)


def fq2_add(a, b):
    return ((a[0] + b[0]) % P_MOD, (a[1] + b[1]) % P_MOD)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % P_MOD, (a[1] - b[1]) % P_MOD)


def fq2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P_MOD)
    return (a[0] * d % P_MOD, -a[1] * d % P_MOD)


B2 = fq2_mul((3, 0), fq2_inv((9, 1)))


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return fq2_sub(fq2_mul(y, y), fq2_add(fq2_mul(fq2_mul(x, x), x), B2)) == (0, 0)


def g2_double(pt):
    if pt is None:
        return None
    x, y = pt
    if y == (0, 0):
        return None
    m = fq2_mul(fq2_mul((3, 0), fq2_mul(x, x)), fq2_inv(fq2_mul((2, 0), y)))
    nx = fq2_sub(fq2_mul(m, m), fq2_mul((2, 0), x))
    ny = fq2_sub(fq2_mul(m, fq2_sub(x, nx)), y)
    return (nx, ny)


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        return g2_double(a) if a[1] == b[1] else None
    m = fq2_mul(fq2_sub(b[1], a[1]), fq2_inv(fq2_sub(b[0], a[0])))
    nx = fq2_sub(fq2_sub(fq2_mul(m, m), a[0]), b[0])
    ny = fq2_sub(fq2_mul(m, fq2_sub(a[0], nx)), a[1])
    return (nx, ny)


def g2_mul(pt, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, pt)
        pt = g2_double(pt)
        k >>= 1
    return acc


# ---- pairing --------------------------------------------------------------------------------------------------
_W = FQ12([0, 1] + [0] * 10)
_W2 = _W * _W
_W3 = _W2 * _W


def twist(pt):
    """G2 point over Fq2 -> the isomorphic curve y^2 = x^3 + 3 over Fq12."""
    if pt is None:
        return None
    (x0, x1), (y0, y1) = pt
    nx = FQ12([x0 - 9 * x1, 0, 0, 0, 0, 0, x1, 0, 0, 0, 0, 0])       # u = w^6 - 9
    ny = FQ12([y0 - 9 * y1, 0, 0, 0, 0, 0, y1, 0, 0, 0, 0, 0])
    return (nx * _W2, ny * _W3)


def _cast_g1(pt):
    return (FQ12.scalar(pt[0]), FQ12.scalar(pt[1]))


def _double12(pt):
    x, y = pt
    m = (x * x * 3) / (y * 2)
    nx = m * m - x * 2
    ny = m * (x - nx) - y
    return (nx, ny)


def _add12(a, b):
    if a[0] == b[0]:
        assert a[1] == b[1]
        return _double12(a)
    m = (b[1] - a[1]) / (b[0] - a[0])
    nx = m * m - a[0] - b[0]
    ny = m * (a[0] - nx) - a[1]
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not x1 == x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1 * 3) / (y1 * 2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q2, p1):
    """Unreduced optimal-ate Miller function f_{6t+2,Q}(P) * line corrections; Q in G2, P in G1 (both not identity)."""
    Q, P = twist(q2), _cast_g1(p1)
    R, f = Q, FQ12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(R, R, P)
        R = _double12(R)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(R, Q, P)
            R = _add12(R, Q)
    Q1 = (Q[0] ** P_MOD, Q[1] ** P_MOD)
    nQ2 = (Q1[0] ** P_MOD, -(Q1[1] ** P_MOD))
    f = f * _linefunc(R, Q1, P)
    R = _add12(R, Q1)
    f = f * _linefunc(R, nQ2, P)
    return f


def final_exponentiate(f):
    return f ** ((P_MOD ** 12 - 1) // R_MOD)


def pairing(q2, p1):
    if q2 is None or p1 is None:
        return FQ12.one()
    return final_exponentiate(miller_loop(q2, p1))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs (G1 point, G2 point): one shared final exponentiation."""
    f = FQ12.one()
    for p1, q2 in pairs:
        if p1 is None or q2 is None:
            continue
        f = f * miller_loop(q2, p1)
    return final_exponentiate(f) == FQ12.one()
