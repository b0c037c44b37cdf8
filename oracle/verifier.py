"""Independent verifier for proofs made by zk-fhe_b200's prover (oracle; test-only).

Restates the verifier side of the halo2-style protocol documented at the top of
zk-fhe_b200/csrc/prover.cu [the reference reaches halo2 `verify_proof` through
halo2-scaffold's `verify` subcommand, README.md:48-54; those crates are un-vendored, so the
reference's own proofs / vk cannot be produced here: PARITY UNPINNED at this level].

What this checks is soundness-relevant and independent of the CUDA code: transcript replay,
every gate / permutation / lookup identity at the challenge point x against the quotient,
and the SHPLONK multi-open.  The final KZG check is done with the test SRS trapdoor
(tests know tau):  [L(tau)] == (tau - u) * [W'(tau)]  in G1, which is the statement the
pairing e(F + u W', G2) == e(W', tau G2) proves.
"""
from . import cbind, curve, field, transcript
from .field import R_MOD, FR_DELTA, omega

BLINDING_FACTORS = 6
PERM_CHUNK = 2
ROT_LAST = 1000
SET_ROTS = [[0], [0, 1, 2, 3], [0, 1, 2], [0, -1], [0, 1], [0, 1, ROT_LAST]]
SET_0, SET_0123, SET_012, SET_0m1, SET_01, SET_01L = range(6)


class VerifyError(AssertionError):
    pass


class Vk:
    """Everything the verifier knows: layout numbers + fixed commitments (affine ints or None)."""

    def __init__(self, info, unusable_rows, fixed_commitments):
        self.k = info["k"]
        self.n = 1 << self.k
        self.n_gate0, self.n_gate1, self.n_rlc, self.n_lookup = (info[x] for x in ("n_gate0", "n_gate1", "n_rlc", "n_lookup"))
        self.n_advice, self.n_perm, self.n_fixed, self.n_chunks = (info[x] for x in ("n_advice", "n_perm", "n_fixed", "n_chunks"))
        self.usable = info["usable_rows"]
        self.lookup_bits = info["lookup_bits"]
        self.instances = info["instances"]
        self.unusable_rows = unusable_rows
        self.fixed = fixed_commitments
        n_gate = self.n_gate0 + self.n_gate1
        n_sel = n_gate + self.n_rlc
        self.fx_qgate, self.fx_qrlc, self.fx_const, self.fx_table = 0, n_gate, n_sel, n_sel + 1
        self.fx_l0, self.fx_sigma = n_sel + 2, n_sel + 5
        assert self.n_fixed == self.fx_sigma + self.n_perm and len(fixed_commitments) == self.n_fixed
        t = transcript.Blake2bTranscript()
        for s in (self.k, self.n_gate0, self.n_gate1, self.n_rlc, self.n_lookup, unusable_rows, self.lookup_bits,
                  self.instances, BLINDING_FACTORS, PERM_CHUNK):
            t.common_scalar(s)
        for c in fixed_commitments:
            t.common_point(c)
        self.digest = t.squeeze()


def _msm(points, scalars):
    """sum scalars[i] * points[i] over affine int points (None = identity), via the C oracle."""
    import numpy as np
    pts = [p for p in points]
    n = 1
    while n < len(pts):
        n <<= 1
    buf = b"".join(curve.g1_to_mont_bytes(p) for p in pts) + b"\0" * (64 * (n - len(pts)))
    bases = np.frombuffer(buf, dtype=np.uint64).reshape(n, 8).copy()
    sc = cbind.ints_to_u64x4([field.to_mont(s % R_MOD) for s in scalars] + [0] * (n - len(pts)))
    out = cbind.msm(sc, bases, n, 1)
    return curve.g1_from_mont_bytes(out[0].tobytes())


def verify(vk, instances, proof, tau, transcript_kind=0, s_g2=None):
    """Raises VerifyError unless `proof` is valid for `instances` under `vk`.  tau: SRS trapdoor."""
    n, k, usable = vk.n, vk.k, vk.usable
    r = R_MOD
    n_gate = vk.n_gate0 + vk.n_gate1
    pos = 0

    def rd_point():
        nonlocal pos
        # 32 bytes as halo2curves writes a bn256 G1Affine: x little-endian, bit 6 of the last byte
        # = parity of y, bit 7 = identity
        raw = bytearray(proof[pos:pos + 32])
        if len(raw) != 32:
            raise VerifyError("proof truncated")
        pos += 32
        inf, odd = bool(raw[31] & 0x80), bool(raw[31] & 0x40)
        raw[31] &= 0x3F
        x = int.from_bytes(raw, "little")
        if inf:
            if x != 0 or odd:
                raise VerifyError("commitment is not a curve point")
            pt = None
        else:
            if x >= field.P_MOD:
                raise VerifyError("commitment is not a curve point")
            rhs = (x * x * x + 3) % field.P_MOD
            y = pow(rhs, (field.P_MOD + 1) // 4, field.P_MOD)
            if y * y % field.P_MOD != rhs:
                raise VerifyError("commitment is not a curve point")
            if bool(y & 1) != odd:
                y = field.P_MOD - y
            pt = (x, y)
            if not curve.is_on_curve(pt):
                raise VerifyError("commitment is not a curve point")
        tr.common_point(pt)
        return pt

    def rd_scalar():
        nonlocal pos
        v = int.from_bytes(proof[pos:pos + 32], "little")
        pos += 32
        if v >= r:
            raise VerifyError("non-canonical scalar")
        tr.common_scalar(v)
        return v

    if len(instances) != vk.instances:
        raise VerifyError("wrong number of instances")
    tr = transcript.make(transcript_kind)
    tr.common_scalar(vk.digest)
    for v in instances:
        tr.common_scalar(v)
    advice_cm = [rd_point() for _ in range(vk.n_gate0)]
    gamma_rlc = tr.squeeze()
    advice_cm += [rd_point() for _ in range(vk.n_advice - vk.n_gate0)]
    tr.squeeze()                                                    # theta (single-column lookups)
    lookup_cm = [rd_point() for _ in range(2 * vk.n_lookup)]        # A'_l, S'_l interleaved
    beta = tr.squeeze()
    gamma = tr.squeeze()
    zp_cm = [rd_point() for _ in range(vk.n_chunks)]
    zl_cm = [rd_point() for _ in range(vk.n_lookup)]
    r_cm = rd_point()
    y = tr.squeeze()
    h_cm = [rd_point() for _ in range(3)]
    x = tr.squeeze()

    # ---- opening table (same order as the prover) ----------------------------------------
    table = []      # (commitment, set)
    for c in range(vk.n_advice):
        table.append((advice_cm[c], SET_0123 if c < n_gate else SET_012 if c < n_gate + vk.n_rlc else SET_0))
    fixed_idx = [f for f in range(vk.n_fixed) if not (vk.fx_l0 <= f < vk.fx_sigma)]
    for f in fixed_idx:
        table.append((vk.fixed[f], SET_0))
    for l in range(vk.n_lookup):
        table += [(lookup_cm[2 * l], SET_0m1), (lookup_cm[2 * l + 1], SET_0), (zl_cm[l], SET_01)]
    for j in range(vk.n_chunks):
        table.append((zp_cm[j], SET_01L if j + 1 < vk.n_chunks else SET_01))
    table.append((r_cm, SET_0))
    xn = pow(x, n, r)
    h_comb_cm = curve.g1_add(curve.g1_add(h_cm[0], curve.g1_mul(h_cm[1], xn)), curve.g1_mul(h_cm[2], xn * xn % r))
    table.append((h_comb_cm, SET_0))
    evals = []
    for i, (_, s) in enumerate(table):
        if i + 1 == len(table):
            evals.append(None)                                      # h_comb(x): recomputed below
        else:
            evals.append([rd_scalar() for _ in SET_ROTS[s]])

    # ---- named evaluations ------------------------------------------------------------------------
    ev = iter(evals)
    adv = [next(ev) for _ in range(vk.n_advice)]
    fixed = dict(zip(fixed_idx, (e[0] for e in (next(ev) for _ in fixed_idx))))
    lk = [(next(ev), next(ev), next(ev)) for _ in range(vk.n_lookup)]     # A' [x, w^-1 x], S' [x], Z [x, wx]
    zp = [next(ev) for _ in range(vk.n_chunks)]
    r_eval = next(ev)[0]

    w = omega(k)
    # Lagrange basis at x: l_i(x) = w^i (x^n - 1) / (n (x - w^i))
    zh = (xn - 1) % r
    if zh == 0:
        raise VerifyError("challenge x lies in the domain")
    rows = list(range(vk.instances)) + [0, usable] + list(range(usable + 1, n))
    wi = {i: pow(w, i, r) for i in set(rows)}
    invs = dict(zip(wi, field.batch_inv([(x - wi[i]) % r for i in wi])))
    ninv = pow(n, -1, r)

    def lag(i):
        return wi[i] * zh % r * ninv % r * invs[i] % r

    l0, l_last = lag(0), lag(usable)
    l_blind = sum(lag(i) for i in range(usable + 1, n)) % r
    l_act = (1 - l_last - l_blind) % r
    inst_eval = sum(v * lag(i) for i, v in enumerate(instances)) % r

    # ---- all identities at x, folded with powers of y (Horner, expression 0 first) ------------------------
    exprs = []
    for c in range(n_gate):
        a0, a1, a2, a3 = adv[c]
        exprs.append(fixed[vk.fx_qgate + c] * (a0 + a1 * a2 - a3) % r)
    for j in range(vk.n_rlc):
        a0, a1, a2 = adv[n_gate + j]
        exprs.append(fixed[vk.fx_qrlc + j] * (a0 * gamma_rlc + a1 - a2) % r)
    m = vk.n_chunks

    def perm_val(c):
        return adv[c][0] if c < vk.n_advice else fixed[vk.fx_const] if c == vk.n_advice else inst_eval

    exprs.append(l0 * (1 - zp[0][0]) % r)
    exprs.append(l_last * (zp[m - 1][0] ** 2 - zp[m - 1][0]) % r)
    for j in range(1, m):
        exprs.append(l0 * (zp[j][0] - zp[j - 1][2]) % r)
    for j in range(m):
        left, right = zp[j][1], zp[j][0]
        for c in range(j * PERM_CHUNK, min((j + 1) * PERM_CHUNK, vk.n_perm)):
            v = perm_val(c)
            left = left * (v + beta * fixed[vk.fx_sigma + c] + gamma) % r
            right = right * (v + beta * pow(FR_DELTA, c, r) * x + gamma) % r
        exprs.append(l_act * (left - right) % r)
    lookup_adv_base = n_gate + vk.n_rlc
    for l in range(vk.n_lookup):
        (ap, ap_m1), (sp,), (z, z_w) = lk[l]
        a, s = adv[lookup_adv_base + l][0], fixed[vk.fx_table]
        exprs.append(l0 * (1 - z) % r)
        exprs.append(l_last * (z * z - z) % r)
        exprs.append(l_act * (z_w * (ap + beta) * (sp + gamma) - z * (a + beta) * (s + gamma)) % r)
        exprs.append(l0 * (ap - sp) % r)
        exprs.append(l_act * (ap - sp) * (ap - ap_m1) % r)
    acc = 0
    for e in exprs:
        acc = (acc * y + e) % r
    h_eval = acc * pow(zh, -1, r) % r
    evals[-1] = [h_eval]

    # ---- SHPLONK ------------------------------------------------------------------------------------------
    yq = tr.squeeze()
    v = tr.squeeze()
    W = rd_point()
    u = tr.squeeze()
    Wp = rd_point()
    if pos != len(proof):
        raise VerifyError("trailing bytes in proof")
    pts = {-1: x * pow(w, -1, r) % r, 0: x, 1: x * w % r, 2: x * w * w % r, 3: x * pow(w, 3, r) % r,
           ROT_LAST: x * pow(w, usable, r) % r}
    zt = 1
    for t in pts.values():
        zt = zt * (u - t) % r
    scal_pts, scal = [], []                    # F = sum scal_i * scal_pts_i
    g_coeff = 0
    vpow = 1
    for s in range(6):
        members = [(cm, e) for (cm, st), e in zip(table, evals) if st == s]
        if members:
            T = [pts[rot] for rot in SET_ROTS[s]]
            # r_s(u): Lagrange interpolation of the combined evaluations at u
            comb = [0] * len(T)
            yp = 1
            for _, e in members:
                for i in range(len(T)):
                    comb[i] = (comb[i] + yp * e[i]) % r
                yp = yp * yq % r
            r_u = 0
            for i, ti in enumerate(T):
                num = den = 1
                for q, tq in enumerate(T):
                    if q != i:
                        num = num * (u - tq) % r
                        den = den * (ti - tq) % r
                r_u = (r_u + comb[i] * num % r * pow(den, -1, r)) % r
            z_s = 1
            for t in T:
                z_s = z_s * (u - t) % r
            zc = zt * pow(z_s, -1, r) % r
            yp = 1
            for cm, _ in members:
                scal_pts.append(cm)
                scal.append(vpow * zc % r * yp % r)
                yp = yp * yq % r
            g_coeff = (g_coeff - vpow * zc % r * r_u) % r
        vpow = vpow * v % r
    scal_pts += [curve.G1_GEN, W]
    scal += [g_coeff, (-zt) % r]
    F = _msm(scal_pts, scal)
    lhs = curve.g1_add(F, curve.g1_mul(Wp, u))
    if s_g2 is not None:
        # the verifier's own check, no trapdoor: e(F + u W', [1]_2) * e(-W', [tau]_2) == 1
        from . import pairing
        neg_wp = None if Wp is None else (Wp[0], (-Wp[1]) % curve.P_MOD)
        if not pairing.pairing_product_is_one([(lhs, pairing.G2_GEN), (neg_wp, s_g2)]):
            raise VerifyError("KZG opening check failed (pairing)")
        return True
    rhs = curve.g1_mul(Wp, tau)
    if lhs != rhs:
        raise VerifyError("KZG opening check failed")
    return True
