"""BN254 scalar field Fr and base field Fq as Python ints (oracle; test-only).

Restates the constants of halo2curves `bn256::{Fr,Fq}` [UPSTREAM, un-vendored;
call sites: /root/reference/src/poly_chip.rs:90,135,158,199 read F::MODULUS].
Every constant below is *recomputed* from the two moduli and cross-checked
against SURVEY.md Appendix A in tests/test_oracle_field.py.
"""

# --- moduli -----------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq

FR_S = 28                      # 2-adicity of r-1
FR_GENERATOR = 7               # multiplicative generator used by halo2curves
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)   # order 2^28
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)                     # generator of the odd-order part
FR_ZETA = pow(FR_GENERATOR, (R_MOD - 1) // 3, R_MOD) ** 2 % R_MOD  # cube root of unity (halo2curves Fr::ZETA)

MONT_BITS = 256
MONT_R = 1 << MONT_BITS


def mont_consts(mod):
    """(R mod m, R^2 mod m, R^3 mod m, -m^-1 mod 2^64)."""
    r1 = MONT_R % mod
    r2 = r1 * r1 % mod
    r3 = r2 * r1 % mod
    inv64 = (-pow(mod, -1, 1 << 64)) % (1 << 64)
    return r1, r2, r3, inv64


FR_R, FR_R2, FR_R3, FR_INV64 = mont_consts(R_MOD)
FQ_R, FQ_R2, FQ_R3, FQ_INV64 = mont_consts(P_MOD)


# --- helpers ----------------------------------------------------------------
def to_mont(x, mod=R_MOD):
    return (x << MONT_BITS) % mod


def from_mont(x, mod=R_MOD):
    return x * pow(MONT_R, -1, mod) % mod


_RINV = {R_MOD: pow(MONT_R, -1, R_MOD), P_MOD: pow(MONT_R, -1, P_MOD)}


def from_mont_fast(x, mod=R_MOD):
    return x * _RINV[mod] % mod


def omega(k):
    """Generator of the 2^k-th roots of unity in Fr (halo2 EvaluationDomain)."""
    assert 0 <= k <= FR_S
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R_MOD)


def inv(x, mod=R_MOD):
    return pow(x, -1, mod)


def batch_inv(xs, mod=R_MOD):
    """Montgomery's trick; zeros map to zero (halo2 batch_invert semantics)."""
    acc = 1
    pref = []
    for x in xs:
        pref.append(acc)
        if x:
            acc = acc * x % mod
    acc = pow(acc, -1, mod)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        x = xs[i]
        if x:
            out[i] = acc * pref[i] % mod
            acc = acc * x % mod
    return out


# --- byte layouts -----------------------------------------------------------
def fe_to_le32(x):
    """Canonical 32-byte little-endian (transcript / proof encoding)."""
    return int(x).to_bytes(32, "little")


def fe_to_mont_le32(x, mod=R_MOD):
    """In-memory layout of halo2curves field elements: 4xu64 LE limbs of x*R."""
    return to_mont(x, mod).to_bytes(32, "little")


def mont_le32_to_fe(b, mod=R_MOD):
    return from_mont_fast(int.from_bytes(b, "little"), mod)


def pack_fr_mont(values, mod=R_MOD):
    """List of ints -> bytes (n*32) in Montgomery LE limb form."""
    return b"".join(fe_to_mont_le32(v, mod) for v in values)


def unpack_fr_mont(buf, mod=R_MOD):
    rinv = _RINV[mod]
    return [int.from_bytes(buf[i:i + 32], "little") * rinv % mod for i in range(0, len(buf), 32)]
