"""The C restatement (oracle/c) is pinned against the pure-Python oracle, which is
pinned by the reference's bfv.in / bfv.json known answers."""
import random

import numpy as np

from oracle import cbind, curve, field, ntt
from oracle.poly import Poly


def test_c_field_mul_matches_python():
    rng = random.Random(11)
    for which, mod in ((0, field.R_MOD), (1, field.P_MOD)):
        rinv = pow(1 << 256, -1, mod)
        for _ in range(200):
            a, b = rng.randrange(mod), rng.randrange(mod)
            assert cbind.field_mul(which, a, b) == a * b * rinv % mod
        assert cbind.field_mul(which, mod - 1, mod - 1) == (mod - 1) ** 2 * rinv % mod


def _mont(vals, mod=field.R_MOD):
    return cbind.ints_to_u64x4([field.to_mont(v, mod) for v in vals])


def _unmont(arr, mod=field.R_MOD):
    return [field.from_mont_fast(v, mod) for v in cbind.u64x4_to_ints(arr)]


def test_c_ntt_matches_python():
    rng = random.Random(12)
    for k in (1, 3, 6, 10):
        n = 1 << k
        cols = [[rng.randrange(field.R_MOD) for _ in range(n)] for _ in range(2)]
        for inverse in (False, True):
            data = _mont([v for c in cols for v in c])
            cbind.ntt(data, k, 2, inverse=inverse)
            got = _unmont(data)
            for b, c in enumerate(cols):
                assert got[b * n:(b + 1) * n] == ntt.ntt(c, k, inverse=inverse)
        data = _mont(cols[0])
        cbind.ntt(data, k, 1, inverse=False, coset=True)
        assert _unmont(data) == ntt.coeff_to_extended(cols[0], k)
        cbind.ntt(data, k, 1, inverse=True, coset=True)
        assert _unmont(data) == cols[0]


def test_c_srs_and_msm_match_python():
    tau = 0x1F2E3D4C5B6A7988
    k = 4
    g_py, gl_py = curve.srs_from_tau(tau, 1 << k)
    g, gl = cbind.srs(k, tau)
    for arr, ref in ((g, g_py), (gl, gl_py)):
        for i in range(1 << k):
            assert curve.g1_from_mont_bytes(arr[i].tobytes()) == ref[i]
    rng = random.Random(13)
    scal = [[rng.randrange(field.R_MOD) for _ in range(1 << k)] for _ in range(3)]
    scal[2] = [0, 1, field.R_MOD - 1, 2, 255, 256, 0, 0, 1, 1, 1, 3, 0, 7, 0, 1 << 200]
    out = cbind.msm(_mont([v for c in scal for v in c]), gl, 1 << k, 3)
    for b in range(3):
        assert curve.g1_from_mont_bytes(out[b].tobytes()) == curve.msm_naive(scal[b], gl_py)


def test_c_msm_threaded_mid_size_matches_serial_pippenger():
    k = 8
    g, _ = cbind.srs(k, 0xABCDEF0123456789ABCDEF, want_gl=False)
    pts = [curve.g1_from_mont_bytes(g[i].tobytes()) for i in range(1 << k)]
    rng = random.Random(14)
    sc = [rng.randrange(field.R_MOD) for _ in range(1 << k)]
    out = cbind.msm(_mont(sc), g, 1 << k, 1)
    assert curve.g1_from_mont_bytes(out[0].tobytes()) == curve.msm_pippenger(sc, pts, c=6)


def test_c_stage1_matches_poly_rs_restatement(bfv_input):
    Q = 536870909
    pk0 = [int(x) for x in bfv_input["pk0"]]
    u = [int(x) for x in bfv_input["u"]]
    cyclo = [int(x) for x in bfv_input["cyclo"]]
    prod = cbind.poly_mul(pk0, u)
    P = Poly(pk0, 29).mul(Poly(u, 29))
    assert prod == P.coefficients
    red = cbind.poly_reduce(prod, Q)
    assert red == P.reduce_by_modulus(Q).coefficients
    q, r = cbind.divide_by_cyclo(red, cyclo, Q)
    q_py, r_py = Poly(red, 29).divide_by_cyclo(Poly(cyclo, 29), Q)
    assert q == q_py.coefficients and r == r_py.coefficients
    z = cbind.divide_by_cyclo([0] * 2047, [0] * 1025, Q)
    assert z == ([0] * 1025, [0] * 2049)
