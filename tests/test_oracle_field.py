"""Oracle constants vs SURVEY.md Appendix A (recomputed, then compared)."""
import random

from oracle import curve, field, ntt


def test_fr_constants_match_survey_appendix_a():
    assert field.R_MOD.bit_length() == 254 and field.P_MOD.bit_length() == 254
    assert field.FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert field.FR_DELTA == 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
    assert field.FR_ZETA == 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
    assert pow(field.FR_ZETA, 3, field.R_MOD) == 1 and field.FR_ZETA != 1
    assert field.omega(13) == 0x10E3D295C1599FF535A1BB49F23D81AA03BD0ED25881F9ED12B179AF67F67AE1
    assert field.omega(16) == 0x09D2CC4B5782FBE923E49ACE3F647643A5F5D8FB89091C3ABABD582133584B29
    assert field.omega(19) == 0x0CF1526AAAFAC6BACBB67D11A4077806B123F767E4B0883D14CC0193568FC082
    assert pow(field.omega(13), 1 << 13, field.R_MOD) == 1
    assert pow(field.omega(13), 1 << 12, field.R_MOD) == field.R_MOD - 1


def test_montgomery_constants_match_survey_appendix_a():
    assert field.FR_R == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert field.FR_R2 == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert field.FR_R3 == 0x0CF8594B7FCC657C893CC664A19FCFED2A489CBE1CFBB6B85E94D8E1B4BF0040
    assert field.FR_INV64 == 0xC2E1F593EFFFFFFF
    assert field.FQ_R == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert field.FQ_R2 == 0x06D89F71CAB8351F47AB1EFF0A417FF6B5E71911D44501FBF32CFC5B538AFA89
    assert field.FQ_R3 == 0x20FD6E902D592544EF7F0B0C0ADA0AFB62F210E6A7283DB6B1CD6DAFDA1530DF
    assert field.FQ_INV64 == 0x87D20782E4866389


def test_mont_roundtrip_and_batch_inv():
    rng = random.Random(1)
    xs = [rng.randrange(field.R_MOD) for _ in range(50)] + [0, 1, field.R_MOD - 1]
    for x in xs:
        assert field.from_mont(field.to_mont(x)) == x
        assert field.mont_le32_to_fe(field.fe_to_mont_le32(x)) == x
    inv = field.batch_inv(xs)
    for x, i in zip(xs, inv):
        assert (x * i) % field.R_MOD == (1 if x else 0)
    assert field.unpack_fr_mont(field.pack_fr_mont(xs)) == xs


def test_ntt_golden_and_roundtrip(digests):
    for key, v in digests["ntt"].items():
        k = int(key[1:])
        a = [int(x, 16) for x in v["in"]]
        out = [int(x, 16) for x in v["out"]]
        assert ntt.ntt(a, k) == out
        assert ntt.ntt(out, k, inverse=True) == a
    rng = random.Random(2)
    a = [rng.randrange(field.R_MOD) for _ in range(32)]
    assert ntt.ntt(a, 5) == ntt.dft_naive(a, 5)
    # coset evaluation agrees with Horner at zeta * w_ext^j
    ext = ntt.coeff_to_extended(a, 7)
    w = field.omega(7)
    for j in (0, 1, 5, 127):
        x = field.FR_ZETA * pow(w, j, field.R_MOD) % field.R_MOD
        assert ext[j] == ntt.poly_eval(a, x)
    assert ntt.extended_to_coeff(ext, 7)[:32] == a


def test_curve_basics_and_msm_golden(digests):
    G = curve.G1_GEN
    assert curve.is_on_curve(G)
    assert curve.g1_mul(G, field.R_MOD) is None
    assert curve.g1_add(G, G) == curve.g1_mul(G, 2)
    assert curve.g1_add(curve.g1_mul(G, 5), curve.g1_mul(G, field.R_MOD - 5)) is None
    m = digests["msm"]
    pts = [(int(x, 16), int(y, 16)) for x, y in m["points"]]
    sc = [int(s, 16) for s in m["scalars"]]
    want = (int(m["result"][0], 16), int(m["result"][1], 16))
    assert all(curve.is_on_curve(p) for p in pts)
    assert curve.msm_naive(sc, pts) == want
    assert curve.msm_pippenger(sc, pts, c=5) == want
    for p in pts[:3]:
        assert curve.g1_from_mont_bytes(curve.g1_to_mont_bytes(p)) == p
    assert curve.g1_from_mont_bytes(curve.g1_to_mont_bytes(None)) is None


def test_toy_srs_lagrange_consistency():
    tau = 0xDEADBEEF
    g, gl = curve.srs_from_tau(tau, 8)
    # committing the same polynomial in both bases gives the same point
    rng = random.Random(3)
    coeffs = [rng.randrange(field.R_MOD) for _ in range(8)]
    evals = ntt.ntt(coeffs, 3)
    assert curve.msm_naive(coeffs, g) == curve.msm_naive(evals, gl)
