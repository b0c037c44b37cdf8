"""Host-side pieces of the `verify` path that need no GPU: zkfhe_pairing_check and zkfhe_srs_g2 of
libzkfhe_b200.so against the oracle's pairing / G2 arithmetic."""
import ctypes

from oracle import curve, pairing
from oracle.field import P_MOD, R_MOD

import zk_fhe_b200
from zk_fhe_b200.capi import _addr, fr_mont_bytes


def _fq(x):
    return ((x << 256) % P_MOD).to_bytes(32, "little")


def _g1(p):
    return bytes(64) if p is None else _fq(p[0]) + _fq(p[1])


def _g2(q):
    return bytes(128) if q is None else _fq(q[0][0]) + _fq(q[0][1]) + _fq(q[1][0]) + _fq(q[1][1])


def _check(pairs):
    lib = zk_fhe_b200.load_library()
    g1 = bytearray(b"".join(_g1(p) for p, _ in pairs))
    g2 = bytearray(b"".join(_g2(q) for _, q in pairs))
    out = ctypes.c_int(-1)
    rc = lib.zkfhe_pairing_check(_addr(g1), _addr(g2), len(pairs), ctypes.byref(out))
    assert rc == 0
    return bool(out.value)


def test_pairing_check_agrees_with_oracle_on_products():
    G = curve.G1_GEN
    neg_g = (G[0], (-G[1]) % P_MOD)
    for a in (1, 2, 0xABCDEF123456789, R_MOD - 5):
        good = [(curve.g1_mul(G, a), pairing.G2_GEN), (neg_g, pairing.g2_mul(pairing.G2_GEN, a))]
        bad = [(curve.g1_mul(G, a + 1), pairing.G2_GEN), (neg_g, pairing.g2_mul(pairing.G2_GEN, a))]
        assert _check(good) and pairing.pairing_product_is_one(good)
        assert not _check(bad) and not pairing.pairing_product_is_one(bad)
    assert _check([(None, pairing.G2_GEN), (neg_g, None)])           # identities contribute 1
    assert not _check([(G, pairing.G2_GEN)])                           # non-degenerate
    assert _check([])


def test_pairing_check_kzg_identity():
    tau, z = 0x5EED5EED5EED, 0x1234
    coeffs = [11, 22, 33, 44]
    ev = lambda at: sum(c * pow(at, i, R_MOD) for i, c in enumerate(coeffs)) % R_MOD
    c_pt = curve.g1_mul(curve.G1_GEN, ev(tau))
    w_pt = curve.g1_mul(curve.G1_GEN, (ev(tau) - ev(z)) * pow(tau - z, -1, R_MOD) % R_MOD)
    lhs = curve.g1_add(curve.g1_add(c_pt, curve.g1_mul(curve.G1_GEN, (-ev(z)) % R_MOD)), curve.g1_mul(w_pt, z))
    neg_w = (w_pt[0], (-w_pt[1]) % P_MOD)
    assert _check([(lhs, pairing.G2_GEN), (neg_w, pairing.g2_mul(pairing.G2_GEN, tau))])
    assert not _check([(lhs, pairing.G2_GEN), (neg_w, pairing.g2_mul(pairing.G2_GEN, tau + 1))])


def test_pairing_check_rejects_unreduced_coordinates():
    lib = zk_fhe_b200.load_library()
    g1 = bytearray(b"\xff" * 64)
    g2 = bytearray(_g2(pairing.G2_GEN))
    out = ctypes.c_int(-1)
    assert lib.zkfhe_pairing_check(_addr(g1), _addr(g2), 1, ctypes.byref(out)) == -2


def test_srs_g2_matches_oracle():
    lib = zk_fhe_b200.load_library()
    for tau in (1, 2, 0x5EED5EED5EED5EED5EED5EED5EED, R_MOD - 1):
        out = bytearray(128)
        t = fr_mont_bytes(tau)
        assert lib.zkfhe_srs_g2(_addr(t), _addr(out)) == 0
        assert bytes(out) == _g2(pairing.g2_mul(pairing.G2_GEN, tau))


def test_pairing_values_fast_path_equals_reference_construction_equals_oracle():
    """zkfhe_pairing: the production path (Q on the twist over Fq2, sparse lines, Frobenius maps, BN final-exponentiation
    chain) returns bit for bit what the plain construction returns, and both equal the Python oracle's GT element."""
    lib = zk_fhe_b200.load_library()
    r_inv = pow(1 << 256, -1, P_MOD)
    for a, b in ((1, 1), (0xABCDEF123456789, 0x1234567), (R_MOD - 2, 3)):
        p1, q2 = curve.g1_mul(curve.G1_GEN, a), pairing.g2_mul(pairing.G2_GEN, b)
        want = pairing.pairing(q2, p1).c
        pb, qb = bytearray(_g1(p1)), bytearray(_g2(q2))
        for reference_construction in (0, 1):
            out = bytearray(384)
            assert lib.zkfhe_pairing(_addr(pb), _addr(qb), reference_construction, _addr(out)) == 0
            got = [int.from_bytes(out[32 * i:32 * i + 32], "little") * r_inv % P_MOD for i in range(12)]
            assert got == want, (a, b, reference_construction)
    one = bytearray(384)
    zero_pt = bytearray(64)
    qb = bytearray(_g2(pairing.G2_GEN))
    assert lib.zkfhe_pairing(_addr(zero_pt), _addr(qb), 0, _addr(one)) == 0          # e(identity, Q) = 1
    assert int.from_bytes(one[:32], "little") * r_inv % P_MOD == 1 and not any(one[32:])


def test_proof_point_encoding_round_trips_and_rejects_non_points():
    """The 32-byte point encoding of the proof (halo2curves' bn256 G1Affine as SURVEY App. C.2 recalls it), host code
    of the library against the oracle's curve arithmetic: multiples of the generator and their negatives round-trip,
    the identity is the flag alone, and bytes that are not an encoding are refused."""
    import ctypes
    import random

    import zk_fhe_b200
    from oracle import curve, field
    lib = zk_fhe_b200.load_library()
    P = field.P_MOD

    def compress(pt):
        raw = (b"\0" * 64) if pt is None else pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little")
        src, out = (ctypes.c_char * 64).from_buffer_copy(raw), (ctypes.c_char * 32)()
        assert lib.zkfhe_point_compress(ctypes.addressof(src), ctypes.addressof(out)) == 0
        return bytes(out)

    def decompress(enc):
        src, out = (ctypes.c_char * 32).from_buffer_copy(enc), (ctypes.c_char * 64)()
        if lib.zkfhe_point_decompress(ctypes.addressof(src), ctypes.addressof(out)) != 0:
            return "invalid"
        x, y = int.from_bytes(bytes(out)[:32], "little"), int.from_bytes(bytes(out)[32:], "little")
        return None if x == 0 and y == 0 else (x, y)

    rnd = random.Random(5)
    for k in [1, 2, 3, 7, field.R_MOD - 1] + [rnd.randrange(1, field.R_MOD) for _ in range(40)]:
        pt = curve.g1_mul(curve.G1_GEN, k)
        for q in (pt, (pt[0], P - pt[1])):
            enc = compress(q)
            assert enc == curve.g1_compress(q)                          # the oracle's restatement of the encoding
            assert len(enc) == 32 and (enc[31] & 0x80) == 0 and ((enc[31] >> 6) & 1) == (q[1] & 1)
            assert int.from_bytes(enc[:31] + bytes([enc[31] & 0x3F]), "little") == q[0]
            assert decompress(enc) == q
    ident = compress(None)
    assert ident == curve.g1_compress(None)
    assert ident == b"\0" * 31 + b"\x80" and decompress(ident) is None
    assert decompress(b"\0" * 31 + b"\xc0") == "invalid"                       # identity flag with a sign bit
    assert decompress((1).to_bytes(31, "little") + b"\x80") == "invalid"       # identity flag with a non-zero x
    assert decompress(P.to_bytes(32, "little")) == "invalid"                  # x = p is not canonical
    non_residue = next(x for x in range(2, 50) if pow((x ** 3 + 3) % P, (P - 1) // 2, P) != 1)
    assert decompress(non_residue.to_bytes(32, "little")) == "invalid"        # x^3 + 3 is not a square
    assert decompress(b"\0" * 32) == "invalid"                                # x = 0 without the flag: 3 is not a square mod p
