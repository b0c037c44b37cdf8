#!/usr/bin/env python3
"""Generates poseidon_consts.h: round constants and MDS matrix of the Poseidon permutation
used by the Fiat-Shamir transcript (t = 5, rate 4, R_F = 8, R_P = 60, alpha = 5 over BN254 Fr
-- the parameter set of snark-verifier's PoseidonTranscript).

Constants follow the published Poseidon parameter generation (Grain LFSR in self-shrinking
mode, rejection sampling of field elements, Cauchy MDS matrix M[i][j] = 1/(x_i + y_j)) -- the
procedure of the Poseidon reference implementation (generate_parameters_grain.sage), which the
PSE `poseidon` crate (the un-vendored dependency of snark-verifier) restates in its `grain.rs`
[UPSTREAM-RECALL].  The generator is PINNED by published vectors: run with (t, R_P) = (3, 57) and
(2, 56) it reproduces the reference implementation's / circomlib's first round constants and the
known answers permute(0,1,2) = 0x115cc0f5...417189a and permute(0,1) = 0x29176100...02820133
(tests/test_oracle_transcript.py, on oracle/transcript.py's independent copy of the procedure and
on this file).  For (5, 60) no published vector exists offline; the tables are what the same,
KAT-checked procedure yields for the parameter set snark-verifier instantiates.

The header carries the permutation in its OPTIMISED form (Poseidon paper App. B; the shape of the
PSE crate's `OptimizedConstants`): partial rounds add one scalar constant and multiply by a sparse
matrix (9 products instead of 25).  `optimise()` derives those tables from (RC, MDS) and checks
the optimised permutation against the plain one on random states before anything is written.
"""
import os

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
T, R_F, R_P, N_BITS = 5, 8, 60, 254


class Grain:
    def __init__(self, n_bits, t, r_f, r_p):
        bits = []
        for val, width in ((1, 2), (0, 4), (n_bits, 12), (t, 12), (r_f, 10), (r_p, 10)):
            bits += [(val >> (width - 1 - i)) & 1 for i in range(width)]
        bits += [1] * 30
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._update()

    def _update(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self):
        while True:
            b = self._update()
            while b == 0:
                self._update()
                b = self._update()
            return self._update()

    def bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v

    def field_element(self):
        while True:
            v = self.bits(N_BITS)
            if v < R_MOD:
                return v


def generate():
    g = Grain(N_BITS, T, R_F, R_P)
    rc = [g.field_element() for _ in range((R_F + R_P) * T)]
    while True:
        xs_ys = [g.bits(N_BITS) % R_MOD for _ in range(2 * T)]
        if len(set(xs_ys)) != 2 * T:
            continue
        xs, ys = xs_ys[:T], xs_ys[T:]
        if any((x + y) % R_MOD == 0 for x in xs for y in ys):
            continue
        mds = [[pow((x + y) % R_MOD, -1, R_MOD) for y in ys] for x in xs]
        return rc, mds


def inv_matrix(m):
    """Inverse of a small matrix over Fr (Gauss-Jordan)."""
    k = len(m)
    a = [list(row) + [int(i == j) for j in range(k)] for i, row in enumerate(m)]
    for c in range(k):
        piv = next(r for r in range(c, k) if a[r][c] % R_MOD)
        a[c], a[piv] = a[piv], a[c]
        iv = pow(a[c][c], -1, R_MOD)
        a[c] = [v * iv % R_MOD for v in a[c]]
        for r in range(k):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % R_MOD for x, y in zip(a[r], a[c])]
    return [row[k:] for row in a]


def mat_mul(a, b):
    return [[sum(a[i][k] * b[k][j] for k in range(len(b))) % R_MOD for j in range(len(b[0]))] for i in range(len(a))]


def mat_vec(m, v):
    return [sum(m[i][j] * v[j] for j in range(len(v))) % R_MOD for i in range(len(m))]


def permute_plain(s, rc, mds):
    half = R_F // 2
    for r in range(R_F + R_P):
        s = [(s[i] + rc[r * T + i]) % R_MOD for i in range(T)]
        if r < half or r >= half + R_P:
            s = [pow(v, 5, R_MOD) for v in s]
        else:
            s[0] = pow(s[0], 5, R_MOD)
        s = mat_vec(mds, s)
    return s


def optimise(rc, mds):
    """(RC, MDS) -> tables of the equivalent optimised permutation:
         first R_F/2 full rounds: s += C[r]; s = s^5; s = M s
         partial round i:         s0 += K[i]; s0 = s0^5; s = SPARSE_i s  (i < R_P - 1)  |  s = M_LAST s  (i = R_P - 1)
         last R_F/2 full rounds:  s += C'[r]; s = s^5; s = M s           (C' differs from RC in its first round only)
       SPARSE_i = [[m00, v], [w, I]]: new s0 = m00 s0 + <v, s[1:]>, new s_j = w_j s0 + s_j."""
    half = R_F // 2
    rows = [list(rc[r * T:(r + 1) * T]) for r in range(R_F + R_P)]
    # 1. constants: elements 1.. of a partial round's constant commute with the partial S-box; pushed through M they
    #    join the next round's constant
    K = []
    for r in range(half, half + R_P):
        K.append(rows[r][0])
        rest = [0] + rows[r][1:]
        push = mat_vec(mds, rest)
        rows[r + 1] = [(a + b) % R_MOD for a, b in zip(rows[r + 1], push)]
    # 2. matrices: M_i = M' M'' with M' = diag(1, Mhat) commuting with the partial S-box (and fixing e0, so the scalar
    #    constants are untouched); M' joins the next round's matrix
    sparse = []
    cur = [list(r) for r in mds]
    for i in range(R_P - 1):
        m00, v = cur[0][0], cur[0][1:]
        w = [cur[j][0] for j in range(1, T)]
        mhat = [cur[j][1:] for j in range(1, T)]
        what = mat_vec(inv_matrix(mhat), w)
        sparse.append((m00, v, what))
        mprime = [[int(j == 0) for j in range(T)]] + [[0] + mhat[j] for j in range(T - 1)]
        cur = mat_mul(mds, mprime)
    m_last = cur
    first = [rows[r] for r in range(half)]
    second = [rows[r] for r in range(half + R_P, R_F + R_P)]
    return first, K, sparse, m_last, second


def permute_optimised(s, mds, opt):
    first, K, sparse, m_last, second = opt
    for c in first:
        s = mat_vec(mds, [pow((a + b) % R_MOD, 5, R_MOD) for a, b in zip(s, c)])
    for i, k in enumerate(K):
        s0 = pow((s[0] + k) % R_MOD, 5, R_MOD)
        if i < len(sparse):
            m00, v, w = sparse[i]
            s = [(m00 * s0 + sum(a * b for a, b in zip(v, s[1:]))) % R_MOD] + [(w[j] * s0 + s[j + 1]) % R_MOD for j in range(T - 1)]
        else:
            s = mat_vec(m_last, [s0] + s[1:])
    for c in second:
        s = mat_vec(mds, [pow((a + b) % R_MOD, 5, R_MOD) for a, b in zip(s, c)])
    return s


def limbs64(x):
    return ", ".join("0x%016xULL" % ((x >> (64 * i)) & (2**64 - 1)) for i in range(4))


def main():
    import random
    rc, mds = generate()
    opt = optimise(rc, mds)
    rnd = random.Random(5)
    for _ in range(8):
        st = [rnd.randrange(R_MOD) for _ in range(T)]
        assert permute_optimised(list(st), mds, opt) == permute_plain(list(st), rc, mds), "optimised Poseidon != plain Poseidon"
    first, K, sparse, m_last, second = opt
    mont = lambda v: (v << 256) % R_MOD
    out = ["// GENERATED by gen_poseidon.py -- do not edit.", "#pragma once", "#include <cstdint>",
           f"#define POSEIDON_T {T}", f"#define POSEIDON_RF {R_F}", f"#define POSEIDON_RP {R_P}",
           "// Montgomery form (R = 2^256), 4 x u64 little-endian limbs",
           "// plain form: round constants and MDS matrix (the optimised tables below are derived from these)",
           f"static const uint64_t POSEIDON_RC[{len(rc)}][4] = {{"]
    out += [f"    {{{limbs64(mont(c))}}}," for c in rc]
    out += ["};", f"static const uint64_t POSEIDON_MDS[{T * T}][4] = {{"]
    out += [f"    {{{limbs64(mont(mds[i][j]))}}}," for i in range(T) for j in range(T)]
    out += ["};", "// optimised form: constants of the full rounds (first half, second half)",
            f"static const uint64_t POSEIDON_C_FIRST[{len(first) * T}][4] = {{"]
    out += [f"    {{{limbs64(mont(c))}}}," for row in first for c in row]
    out += ["};", f"static const uint64_t POSEIDON_C_SECOND[{len(second) * T}][4] = {{"]
    out += [f"    {{{limbs64(mont(c))}}}," for row in second for c in row]
    out += ["};", "// one scalar constant per partial round (added to state[0])",
            f"static const uint64_t POSEIDON_K[{len(K)}][4] = {{"]
    out += [f"    {{{limbs64(mont(c))}}}," for c in K]
    out += ["};", "// sparse matrices of partial rounds 0 .. R_P-2: m00, v[T-1] (first row), w[T-1] (first column below m00)",
            f"static const uint64_t POSEIDON_SPARSE[{len(sparse) * (2 * T - 1)}][4] = {{"]
    for m00, v, w in sparse:
        out += [f"    {{{limbs64(mont(c))}}}," for c in [m00] + v + w]
    out += ["};", "// dense matrix of the last partial round", f"static const uint64_t POSEIDON_M_LAST[{T * T}][4] = {{"]
    out += [f"    {{{limbs64(mont(m_last[i][j]))}}}," for i in range(T) for j in range(T)]
    out += ["};"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon_consts.h")
    text = "\n".join(out) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    print("wrote", path)


if __name__ == "__main__":
    main()
