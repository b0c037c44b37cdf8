"""BN254 G1 (y^2 = x^3 + 3 over Fq) in Python ints (oracle; test-only).

Restates halo2curves `bn256::{G1Affine, G1}` [UPSTREAM, un-vendored] and the
definition of halo2's `best_multiexp` (= sum_i s_i * P_i; the value is unique,
so any correct algorithm is bit-exact on the affine result).
Affine points are (x, y) tuples; the identity is None (encoded as (0,0) in the
64-byte in-memory layout, as halo2curves does).
"""
from .field import P_MOD, R_MOD, to_mont, from_mont_fast

B_COEFF = 3
G1_GEN = (1, 2)


def is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P_MOD == 0


# --- Jacobian arithmetic (X, Y, Z), identity Z = 0 ---------------------------
J_INF = (1, 1, 0)


def to_jac(pt):
    return J_INF if pt is None else (pt[0], pt[1], 1)


def to_affine(j):
    X, Y, Z = j
    if Z == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def jac_double(j):
    X, Y, Z = j
    if Z == 0 or Y == 0:
        return J_INF
    p = P_MOD
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def jac_add(a, b):
    if a[2] == 0:
        return b
    if b[2] == 0:
        return a
    p = P_MOD
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return jac_double(a)
        return J_INF
    H = (U2 - U1) % p
    Rr = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (Rr * Rr - HHH - 2 * V) % p
    Y3 = (Rr * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def jac_neg(a):
    return (a[0], (-a[1]) % P_MOD, a[2])


def jac_mul(j, k):
    k %= R_MOD
    acc = J_INF
    while k:
        if k & 1:
            acc = jac_add(acc, j)
        j = jac_double(j)
        k >>= 1
    return acc


def g1_mul(pt, k):
    return to_affine(jac_mul(to_jac(pt), k))


def g1_add(a, b):
    return to_affine(jac_add(to_jac(a), to_jac(b)))


def msm_naive(scalars, points):
    """Definition of the MSM: sum_i scalars[i] * points[i] (affine result)."""
    acc = J_INF
    for s, pt in zip(scalars, points):
        if s % R_MOD and pt is not None:
            acc = jac_add(acc, jac_mul(to_jac(pt), s))
    return to_affine(acc)


def msm_pippenger(scalars, points, c=8):
    """Serial Pippenger (halo2 `best_multiexp` shape: per-window buckets,
    running-sum reduction).  Used to cross-check msm_naive and at mid sizes."""
    nwin = (254 + c - 1) // c
    total = J_INF
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = jac_double(total)
        buckets = [J_INF] * ((1 << c) - 1)
        for s, pt in zip(scalars, points):
            d = ((s % R_MOD) >> (w * c)) & ((1 << c) - 1)
            if d and pt is not None:
                buckets[d - 1] = jac_add(buckets[d - 1], to_jac(pt))
        run = J_INF
        acc = J_INF
        for b in reversed(buckets):
            run = jac_add(run, b)
            acc = jac_add(acc, run)
        total = jac_add(total, acc)
    return to_affine(total)


# --- byte layout: halo2curves G1Affine in memory = x||y, 4xu64 LE Montgomery --
def g1_to_mont_bytes(pt):
    if pt is None:
        return b"\0" * 64
    return to_mont(pt[0], P_MOD).to_bytes(32, "little") + to_mont(pt[1], P_MOD).to_bytes(32, "little")


def g1_from_mont_bytes(b):
    x = from_mont_fast(int.from_bytes(b[:32], "little"), P_MOD)
    y = from_mont_fast(int.from_bytes(b[32:64], "little"), P_MOD)
    if x == 0 and y == 0:
        return None
    return (x, y)


def g1_compress(pt):
    """halo2curves compressed encoding [UPSTREAM-RECALL]: x little-endian,
    bit 6 of byte 31 = y sign (y & 1), bit 7 = identity flag."""
    if pt is None:
        b = bytearray(32)
        b[31] |= 0x80
        return bytes(b)
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << 6
    return bytes(b)


# --- toy SRS (tests only) ---------------------------------------------------
def srs_from_tau(tau, n):
    """g[i] = tau^i * G1 and g_lagrange[i] = l_i(tau) * G1 over the size-n
    domain (halo2 `ParamsKZG::setup` [UPSTREAM-RECALL])."""
    from .field import omega, inv
    k = n.bit_length() - 1
    assert 1 << k == n
    G = to_jac(G1_GEN)
    g = []
    t = 1
    for _ in range(n):
        g.append(to_affine(jac_mul(G, t)))
        t = t * tau % R_MOD
    w = omega(k)
    tn = (pow(tau, n, R_MOD) - 1) % R_MOD
    ninv = inv(n)
    gl = []
    wi = 1
    for _ in range(n):
        # l_i(tau) = w^i (tau^n - 1) / (n (tau - w^i))
        li = wi * tn % R_MOD * ninv % R_MOD * inv((tau - wi) % R_MOD) % R_MOD
        gl.append(to_affine(jac_mul(G, li)))
        wi = wi * w % R_MOD
    return g, gl
