"""The oracle's BN254 pairing (oracle/pairing.py), pinned by the defining properties of a pairing:
bilinearity, non-degeneracy, order r -- and by the KZG identity it exists to check."""
from oracle import curve, pairing
from oracle.field import P_MOD, R_MOD


def test_g2_generator_and_group_law():
    assert pairing.g2_is_on_curve(pairing.G2_GEN)
    assert pairing.g2_mul(pairing.G2_GEN, R_MOD - 1) == (pairing.G2_GEN[0], pairing.fq2_sub((0, 0), pairing.G2_GEN[1]))
    a, b = 0x1234567, 0x89ABCDE
    assert pairing.g2_mul(pairing.g2_mul(pairing.G2_GEN, a), b) == pairing.g2_mul(pairing.G2_GEN, a * b)
    assert pairing.g2_is_on_curve(pairing.g2_mul(pairing.G2_GEN, a))


def test_pairing_is_bilinear_and_non_degenerate():
    e1 = pairing.pairing(pairing.G2_GEN, curve.G1_GEN)
    assert not e1 == pairing.FQ12.one()
    assert e1 ** R_MOD == pairing.FQ12.one()
    a, b = 1234567891011, 98765432101
    assert pairing.pairing(pairing.g2_mul(pairing.G2_GEN, b), curve.g1_mul(curve.G1_GEN, a)) == e1 ** (a * b % R_MOD)
    assert pairing.pairing(pairing.G2_GEN, curve.g1_mul(curve.G1_GEN, a)) * \
        pairing.pairing(pairing.G2_GEN, curve.g1_mul(curve.G1_GEN, b)) == e1 ** ((a + b) % R_MOD)


def test_kzg_opening_identity_under_the_pairing():
    """Commit p(X) = 3 + 5X + 7X^2 under tau, open at z: e(C - p(z) G, [1]_2) == e(W, [tau - z]_2)."""
    tau, z = 0xDEADBEEFCAFE, 0x4242
    coeffs = [3, 5, 7]
    pz = sum(c * pow(z, i, R_MOD) for i, c in enumerate(coeffs)) % R_MOD
    c_pt = curve.g1_mul(curve.G1_GEN, sum(c * pow(tau, i, R_MOD) for i, c in enumerate(coeffs)) % R_MOD)
    # quotient (p(X) - p(z)) / (X - z) = 7X + (5 + 7z)
    w_pt = curve.g1_mul(curve.G1_GEN, (7 * tau + 5 + 7 * z) % R_MOD)
    lhs = curve.g1_add(curve.g1_add(c_pt, curve.g1_mul(curve.G1_GEN, (-pz) % R_MOD)), curve.g1_mul(w_pt, z))
    neg_w = (w_pt[0], (-w_pt[1]) % P_MOD)
    s_g2 = pairing.g2_mul(pairing.G2_GEN, tau)
    assert pairing.pairing_product_is_one([(lhs, pairing.G2_GEN), (neg_w, s_g2)])
    assert not pairing.pairing_product_is_one([(lhs, pairing.G2_GEN), (neg_w, pairing.g2_mul(pairing.G2_GEN, tau + 1))])
