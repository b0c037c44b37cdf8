"""Radix-2 NTT over BN254 Fr (oracle; test-only).

Restates halo2-axiom `arithmetic::best_fft` and `EvaluationDomain`
{lagrange_to_coeff, coeff_to_extended, extended_to_coeff} [UPSTREAM-RECALL,
un-vendored; SURVEY.md App. C.3/C.4]: natural-order input, natural-order
output, out[j] = sum_i a[i] * w^(i*j); the inverse uses w^-1 and scales by
n^-1.  The extended coset is zeta * H_ext with zeta = Fr::ZETA (coefficient i
is multiplied by zeta^(i mod 3) before the forward transform).
"""
from .field import R_MOD, omega, inv, FR_ZETA


def bitrev(i, k):
    r = 0
    for _ in range(k):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def ntt(a, k=None, inverse=False):
    """In: list of ints (natural order).  Out: new list (natural order)."""
    n = len(a)
    if k is None:
        k = n.bit_length() - 1
    assert n == 1 << k
    a = list(a)
    for i in range(n):
        j = bitrev(i, k)
        if i < j:
            a[i], a[j] = a[j], a[i]
    w = omega(k)
    if inverse:
        w = inv(w)
    half = n >> 1
    tw = [1] * max(half, 1)
    for i in range(1, half):
        tw[i] = tw[i - 1] * w % R_MOD
    m = 1
    while m < n:
        step = half // m
        for s in range(0, n, 2 * m):
            for j in range(m):
                t = a[s + j + m] * tw[j * step] % R_MOD
                u = a[s + j]
                a[s + j] = (u + t) % R_MOD
                a[s + j + m] = (u - t) % R_MOD
        m <<= 1
    if inverse:
        ninv = inv(n)
        a = [x * ninv % R_MOD for x in a]
    return a


def dft_naive(a, k, inverse=False):
    n = 1 << k
    w = omega(k)
    if inverse:
        w = inv(w)
    out = []
    for j in range(n):
        wj = pow(w, j, R_MOD)
        acc = 0
        x = 1
        for i in range(n):
            acc = (acc + a[i] * x) % R_MOD
            x = x * wj % R_MOD
        out.append(acc)
    if inverse:
        ninv = inv(n)
        out = [x * ninv % R_MOD for x in out]
    return out


def coset_powers(n, g):
    out = [1] * n
    for i in range(1, n):
        out[i] = out[i - 1] * g % R_MOD
    return out


def coeff_to_extended(coeffs, k_ext, g=FR_ZETA):
    """Coefficients (len <= 2^k_ext) -> evaluations on g * H_ext."""
    n = 1 << k_ext
    a = list(coeffs) + [0] * (n - len(coeffs))
    x = 1
    for i in range(n):
        a[i] = a[i] * x % R_MOD
        x = x * g % R_MOD
    return ntt(a, k_ext)


def extended_to_coeff(evals, k_ext, g=FR_ZETA):
    n = 1 << k_ext
    a = ntt(evals, k_ext, inverse=True)
    gi = inv(g)
    x = 1
    for i in range(n):
        a[i] = a[i] * x % R_MOD
        x = x * gi % R_MOD
    return a


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc
