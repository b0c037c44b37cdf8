"""The Fiat-Shamir transcript, on the CPU: published Poseidon vectors pin the oracle's parameter generation and
permutation; the library's host code (optimised permutation, sponge, BLAKE2b variant) is held against the oracle.

Published vectors (Poseidon reference implementation, `poseidonperm_x5_254_3` / circomlib's poseidon constants and
test suite): for BN254 Fr, x^5, R_F = 8,
  t = 3, R_P = 57: first round constants 0x0ee9a592...cd8e6e, 0x00f14452...56e864; permute(0, 1, 2)[0] =
                   0x115cc0f5...417189a (= circomlib poseidon([1, 2]))
  t = 2, R_P = 56: first round constant 0x09c46e9e...abd7a7; permute(0, 1)[0] = 0x29176100...2820133
                   (= circomlib poseidon([1]))
The transcript's own instance (t = 5, R_P = 60, the parameters of snark-verifier's PoseidonTranscript) comes out of the
same generator; no published vector for it is available offline.
"""
import ctypes
import importlib.util
import os
import random

import pytest

from oracle import transcript
from oracle.field import R_MOD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

T3_RC0 = 0x0EE9A592BA9A9518D05986D656F40C2114C4993C11BB29938D21D47304CD8E6E
T3_RC1 = 0x00F1445235F2148C5986587169FC1BCD887B08D4D00868DF5696FFF40956E864
T3_OUT = [0x115CC0F5E7D690413DF64C6B9662E9CF2A3617F2743245519E19607A4417189A,
          0x0FCA49B798923AB0239DE1C9E7A4A9A2210312B6A2F616D18B5A87F9B628AE29,
          0x0E7AE82E40091E63CBD4F16A6D16310B3729D4B6E138FCF54110E2867045A30C]
T2_RC0 = 0x09C46E9EC68E9BD4FE1FAABA294CBA38A71AA177534CDD1B6C7DC0DBD0ABD7A7
T2_OUT0 = 0x29176100EAA962BDC1FE6C654D6A3C130E96A4D1168B33848B897DC502820133


def _generator():
    spec = importlib.util.spec_from_file_location("gen_poseidon", os.path.join(ROOT, "zk-fhe_b200", "csrc", "gen_poseidon.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_poseidon_matches_published_vectors():
    p3 = transcript.poseidon_params(3, 8, 57)
    assert p3[0][0] == T3_RC0 and p3[0][1] == T3_RC1
    assert transcript.poseidon_permute([0, 1, 2], p3, 8, 57) == T3_OUT
    p2 = transcript.poseidon_params(2, 8, 56)
    assert p2[0][0] == T2_RC0
    assert transcript.poseidon_permute([0, 1], p2, 8, 56)[0] == T2_OUT0


def test_table_generator_matches_published_vectors_and_the_oracle():
    """zk-fhe_b200/csrc/gen_poseidon.py (writes the product's tables) is a second copy of the procedure."""
    g = _generator()
    try:
        g.T, g.R_P = 3, 57
        rc, mds = g.generate()
        assert rc[0] == T3_RC0 and rc[1] == T3_RC1
        assert g.permute_plain([0, 1, 2], rc, mds) == T3_OUT
        assert g.permute_optimised([0, 1, 2], mds, g.optimise(rc, mds)) == T3_OUT      # the sparse form, on a published vector
    finally:
        g.T, g.R_P = 5, 60
    rc, mds = g.generate()
    assert (rc, mds) == tuple(transcript.poseidon_params())
    rnd = random.Random(3)
    st = [rnd.randrange(R_MOD) for _ in range(5)]
    assert g.permute_optimised(list(st), mds, g.optimise(rc, mds)) == transcript.poseidon_permute(list(st))


def test_mds_matrix_is_mds():
    """Every square submatrix of a Cauchy matrix is invertible; check the 1x1 and 2x2 minors and the determinant."""
    _, m = transcript.poseidon_params()
    t = len(m)
    assert all(v % R_MOD for row in m for v in row)
    for i in range(t):
        for j in range(i + 1, t):
            for a in range(t):
                for b in range(a + 1, t):
                    assert (m[i][a] * m[j][b] - m[i][b] * m[j][a]) % R_MOD
    g = _generator()
    inv = g.inv_matrix(m)
    assert g.mat_mul(m, inv) == [[int(i == j) for j in range(t)] for i in range(t)]


# ---- the library's host code (no GPU needed: plain host arithmetic behind the C ABI) ---------------------------
@pytest.fixture(scope="module")
def lib():
    import zk_fhe_b200
    return zk_fhe_b200.load_library()


def _mont(v):
    return (v << 256) % R_MOD


def _unmont(v):
    return v * pow(1 << 256, -1, R_MOD) % R_MOD


def _permute(lib, state, plain, may_be_unavailable=False):
    buf = bytearray(b"".join(_mont(v).to_bytes(32, "little") for v in state))
    arr = (ctypes.c_char * len(buf)).from_buffer(buf)
    rc = lib.zkfhe_poseidon_permute(ctypes.addressof(arr), plain)
    if may_be_unavailable and rc != 0:
        return None
    assert rc == 0
    return [_unmont(int.from_bytes(buf[32 * i:32 * i + 32], "little")) for i in range(5)]


def test_host_poseidon_permutation_equals_oracle(lib):
    rnd = random.Random(9)
    cases = [[0] * 5, [1 << 64, 0, 0, 0, 0], [R_MOD - 1] * 5, [R_MOD - 1, 0, 1, 2, R_MOD - 2]]
    cases += [[rnd.randrange(R_MOD) for _ in range(5)] for _ in range(60)]
    for st in cases:
        want = transcript.poseidon_permute(list(st))
        assert _permute(lib, st, 0) == want          # optimised form (what the transcript runs)
        assert _permute(lib, st, 1) == want          # textbook form
        assert _permute(lib, st, 2) == want          # scalar optimised form (the fallback without AVX-512 IFMA)
        got = _permute(lib, st, 3, may_be_unavailable=True)       # the AVX-512 IFMA form, where this CPU has it
        assert got is None or got == want
    bad = bytearray(b"\xff" * 160)
    arr = (ctypes.c_char * 160).from_buffer(bad)
    assert lib.zkfhe_poseidon_permute(ctypes.addressof(arr), 0) != 0       # non-canonical input is refused


def test_ifma_permutation_equals_scalar_on_many_states(lib):
    """The vector form keeps loosely reduced values and a lagged row product (csrc/poseidon_ifma.cpp): hold it
    against the scalar form on a few thousand states, edge limbs included.  Vacuous without AVX-512 IFMA."""
    rnd = random.Random(2026)
    edge = [0, 1, R_MOD - 1, R_MOD - 2, (1 << 52) - 1, 1 << 52, (1 << 104) - 1, (1 << 208) - 1, (1 << 253), R_MOD >> 1]

    def raw(state, form):
        buf = bytearray(b"".join(v.to_bytes(32, "little") for v in state))       # any canonical value is a valid Montgomery form
        arr = (ctypes.c_char * len(buf)).from_buffer(buf)
        return bytes(buf) if lib.zkfhe_poseidon_permute(ctypes.addressof(arr), form) == 0 else None

    for i in range(3000):
        st = [rnd.choice(edge) if rnd.random() < 0.2 else rnd.randrange(R_MOD) for _ in range(5)]
        want = raw(st, 2)
        assert want is not None
        for form in (0, 3):
            got = raw(st, form)
            assert got is None or got == want, (i, form)


def _replay(lib, kind, items):
    script = bytearray()
    n_sq = 0
    for it in items:
        if it[0] == "s":
            script += b"\x01" + int(it[1]).to_bytes(32, "little")
        elif it[0] == "p":
            x, y = it[1]
            script += b"\x02" + x.to_bytes(32, "little") + y.to_bytes(32, "little")
        else:
            script += b"\x03"
            n_sq += 1
    out = bytearray(32 * n_sq)
    n = ctypes.c_size_t()
    sa = (ctypes.c_char * len(script)).from_buffer(script)
    oa = (ctypes.c_char * max(len(out), 1)).from_buffer(out if out else bytearray(1))
    assert lib.zkfhe_transcript_replay(kind, ctypes.addressof(sa), len(script), ctypes.addressof(oa), len(out), ctypes.byref(n)) == 0
    assert n.value == n_sq
    return [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(n_sq)]


@pytest.mark.parametrize("kind", [0, 1])
def test_host_transcript_equals_oracle(lib, kind):
    from oracle import curve
    rnd = random.Random(21 + kind)
    pts = [curve.g1_mul(curve.G1_GEN, rnd.randrange(1, R_MOD)) for _ in range(6)]
    items = [("s", 12345), ("q",)]                                 # squeeze on a short buffer (padding path)
    items += [("s", rnd.randrange(R_MOD)) for _ in range(3)] + [("q",)]       # exactly rate-1 items + the pad = one block
    items += [("s", rnd.randrange(R_MOD)) for _ in range(4)] + [("q",)]       # exact multiple of the rate: extra block
    items += [("p", p) for p in pts] + [("p", (0, 0)), ("q",), ("q",)]        # points, the identity, back-to-back squeezes
    items += [("s", R_MOD - 1), ("p", pts[0]), ("s", 0), ("q",)]
    got = _replay(lib, kind, items)
    t = transcript.make(kind)
    want = []
    for it in items:
        if it[0] == "s":
            t.common_scalar(it[1])
        elif it[0] == "p":
            t.common_point(None if it[1] == (0, 0) else it[1])
        else:
            want.append(t.squeeze())
    assert got == want


def test_poseidon_point_encoding_reduces_coordinates_mod_r(lib):
    """Coordinates are Fq elements; the transcript absorbs them mod r (Fq modulus < 2r: one subtraction)."""
    from oracle.field import P_MOD as Q_MOD
    x = R_MOD + 5
    assert x < Q_MOD
    got = _replay(lib, 1, [("p", (x, R_MOD - 1)), ("q",)])
    t = transcript.PoseidonTranscript()
    t.common_scalar(5)
    t.common_scalar(R_MOD - 1)
    assert got == [t.squeeze()]


RFC7539_A1_1 = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                             "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


def test_reference_test_srs_trapdoor(lib):
    """`--insecure-test-srs` / bench.py use the trapdoor of the reference's own fallback SRS (`ParamsKZG::setup` from
    `ChaCha20Rng::from_seed([0; 32])`): ChaCha20's zero-key block is RFC 7539 appendix A.1 test vector #1, and tau is
    that block as a little-endian 512-bit integer mod r -- in the oracle and in the library."""
    assert transcript.chacha20_block([0] * 8) == RFC7539_A1_1
    tau = int.from_bytes(RFC7539_A1_1, "little") % R_MOD
    assert transcript.reference_test_tau() == tau
    out, ks = bytearray(32), bytearray(64)
    a = (ctypes.c_char * 32).from_buffer(out)
    b = (ctypes.c_char * 64).from_buffer(ks)
    assert lib.zkfhe_reference_test_tau(ctypes.addressof(a), ctypes.addressof(b)) == 0
    assert bytes(ks) == RFC7539_A1_1
    assert _unmont(int.from_bytes(out, "little")) == tau
    import zk_fhe_b200
    assert zk_fhe_b200.reference_test_tau() == tau
