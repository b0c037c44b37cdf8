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


# ---- stage (1) in C (orc_bfv_witness): the CPU arm's witness generator ---------------------------------------
def _canon(arr):
    return cbind.u64x4_to_ints(cbind.from_mont_array(np.ascontiguousarray(arr)))


def test_c_witness_equals_python_oracle_on_bfv_in(bfv_input, golden_gamma, oracle_tables, digests):
    """All 1,288,314 advice cells and 286,756 lookup cells of bfv.in, cell by cell, against oracle/bfv.py (itself
    pinned by c0 / c1 of bfv.in and by configs/bfv.json) and against the committed digests."""
    import hashlib
    a0, a1, a2, lk = cbind.bfv_witness(bfv_input, 1024, 536870909, 7, 19, golden_gamma)
    tab = oracle_tables
    assert _canon(a0) == tab["phase0"].ctx.advice
    assert _canon(a1) == tab["ctx_gate"].advice
    assert _canon(a2) == tab["ctx_rlc"].advice
    assert _canon(lk) == [v for col in tab["lookup"] for v in col]
    for key, arr in (("phase0_advice_sha256", a0), ("phase1_gate_advice_sha256", a1), ("phase1_rlc_advice_sha256", a2),
                     ("lookup_cells_sha256", lk)):
        assert hashlib.sha256(cbind.from_mont_array(np.ascontiguousarray(arr)).tobytes()).hexdigest() == digests[key]


def test_c_witness_keygen_input_and_small_parameters(bfv_empty_input):
    """The all-zero keygen input (zero shortcut of divide_by_cyclo, src/poly.rs:118-123) and a small (N, Q) against
    the Python oracle; a coefficient above Q is the reference's assert at src/poly.rs:28."""
    import pytest
    from oracle import bfv as obfv
    from oracle.poly import OracleError
    a0, a1, a2, lk = cbind.bfv_witness(bfv_empty_input, 1024, 536870909, 7, 19, 7)
    assert (a0.shape[0], a1.shape[0], a2.shape[0], lk.shape[0]) == (23558, 1231992, 32764, 286756)
    rng = random.Random(4)
    N, Q, T, B = 16, 65521, 5, 9
    par = obfv.BfvParams(N=N, Q=Q, T=T, B=B)
    pk0, pk1 = [rng.randrange(Q) for _ in range(N)], [rng.randrange(Q) for _ in range(N)]
    u = [rng.choice([0, 1, Q - 1]) for _ in range(N)]
    e0, e1 = [rng.randrange(-B, B + 1) % Q for _ in range(N)], [rng.randrange(-B, B + 1) % Q for _ in range(N)]
    m = [rng.randrange(-(T // 2), T // 2 + 1) % Q for _ in range(N)]

    def ring_mul(a, b):
        p = [0] * (2 * N - 1)
        for i in range(N):
            for j in range(N):
                p[i + j] += a[i] * b[j]
        return [(p[N - 1 + i] - (p[i - 1] if i else 0)) % Q for i in range(N)]

    c0 = [(x + (Q // T) * mm + e) % Q for x, mm, e in zip(ring_mul(pk0, u), m, e0)]
    c1 = [(x + e) % Q for x, e in zip(ring_mul(pk1, u), e1)]
    inp = {"pk0": pk0, "pk1": pk1, "m": m, "u": u, "e0": e0, "e1": e1, "c0": c0, "c1": c1, "cyclo": [1] + [0] * (N - 1) + [1]}
    sinp = {k: [str(x) for x in v] for k, v in inp.items()}
    tab = obfv.build_tables(sinp, 12345, par, k=9, unusable_rows=20)
    a0, a1, a2, lk = cbind.bfv_witness(inp, N, Q, T, B, 12345)
    assert _canon(a0) == tab["phase0"].ctx.advice
    assert _canon(a1) == tab["ctx_gate"].advice
    assert _canon(a2) == tab["ctx_rlc"].advice
    assert _canon(lk) == [v for col in tab["lookup"] for v in col]
    bad = dict(inp)
    bad["e0"] = [Q + 1] + e0[1:]
    with pytest.raises(OracleError):
        cbind.bfv_witness(bad, N, Q, T, B, 12345)


def test_bench_cpu_arm_inputs_are_valid_encryptions():
    """bench.py's CPU-arm input generator: c0, c1 satisfy the circuit's equalities (no is_equal cell is 0)."""
    import bench
    inp = bench.cpu_synth_input(np.random.default_rng(3))
    N, Q = 1024, 536870909
    a0, a1, a2, lk = cbind.bfv_witness(inp, N, Q, 7, 19, 99)
    # the last call is c1.constrain_equality (12 cells per coefficient, is_zero flag at cell 4 of each block)
    tail = _canon(a1[-12 * N:])
    assert all(tail[12 * i + 4] == 1 for i in range(N))
    assert bench.host_threads() >= 1
