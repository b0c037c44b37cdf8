// Host-side BN254 pairing for the `verify` path: the final KZG check
//     e(F + u W', [1]_2) * e(-W', [tau]_2) == 1
// that halo2-axiom's `VerifierSHPLONK` hands to `MultiMillerLoop` [UPSTREAM, un-vendored; reached from
// the reference's `verify` subcommand, README.md:48-54, SURVEY.md §3.4 / §8(f) rank 1].
//
// One check per proof, a few hundred thousand base-field products: host work, like the transcript.
// The construction is deliberately the plainest correct one (and the same one oracle/pairing.py
// restates, so the two can be compared value for value in the CPU tests): Fq12 = Fq[w]/(w^12 - 18 w^6 + 82),
// G2 mapped through the sextic twist into Fq12, affine Miller loop with loop count 6t + 2 and the two
// Frobenius line corrections, plain final exponentiation by (p^12 - 1)/r.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "inv_bin.cuh"

namespace zkfhe { namespace host {

typedef unsigned __int128 u128p;

struct Fq {
    uint64_t l[4];
    bool operator==(const Fq& o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const Fq& o) const { return !(*this == o); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};
static const Fq FQ_MOD = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const Fq FQ_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const Fq FQ_ONE = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
static const Fq FQ_ZERO = {{0, 0, 0, 0}};
static const uint64_t FQ_INV = 0x87d20782e4866389ULL;

inline bool fq_geq(const Fq& a, const Fq& b) {
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] > b.l[i]) return true;
        if (a.l[i] < b.l[i]) return false;
    }
    return true;
}
inline Fq fq_sub_raw(const Fq& a, const Fq& b) {
    Fq r;
    u128p brw = 0;
    for (int i = 0; i < 4; i++) {
        u128p d = (u128p)a.l[i] - b.l[i] - brw;
        r.l[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
    return r;
}
inline Fq fq_add(const Fq& a, const Fq& b) {
    Fq t;
    u128p c = 0;
    for (int i = 0; i < 4; i++) { c += (u128p)a.l[i] + b.l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    return (c || fq_geq(t, FQ_MOD)) ? fq_sub_raw(t, FQ_MOD) : t;
}
inline Fq fq_sub(const Fq& a, const Fq& b) { return fq_geq(a, b) ? fq_sub_raw(a, b) : fq_sub_raw(FQ_MOD, fq_sub_raw(b, a)); }
inline Fq fq_neg(const Fq& a) { return a.is_zero() ? a : fq_sub_raw(FQ_MOD, a); }
inline Fq fq_mul(const Fq& a, const Fq& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t* p = FQ_MOD.l;
    for (int i = 0; i < 4; i++) {
        u128p c = 0;
        for (int j = 0; j < 4; j++) { c += (u128p)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FQ_INV;
        c = ((u128p)m * p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128p)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fq o = {{t[0], t[1], t[2], t[3]}};
    return (t[4] || fq_geq(o, FQ_MOD)) ? fq_sub_raw(o, FQ_MOD) : o;
}
inline Fq fq_to_mont(const Fq& canon) { return fq_mul(canon, FQ_R2); }
inline Fq fq_from_mont(const Fq& m) { Fq one = {{1, 0, 0, 0}}; return fq_mul(m, one); }
inline Fq fq_from_u64(uint64_t v) { Fq c = {{v, 0, 0, 0}}; return fq_to_mont(c); }
static const Fq FQ_R3 = {{0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL}};
// inverse of a Montgomery-form element: binary extended Euclid on the raw limbs (inv_bin.cuh, the routine the
// kernels use) gives (aR)^-1; one product with R^3 brings it back to a^-1 R.  inv(0) = 0.
inline Fq fq_inv(const Fq& a) {
    Fq t;
    u256_inv_odd(reinterpret_cast<uint32_t*>(t.l), reinterpret_cast<const uint32_t*>(a.l), reinterpret_cast<const uint32_t*>(FQ_MOD.l));
    return fq_mul(t, FQ_R3);
}
inline Fq fq_inv_fermat(const Fq& a) {   // the independent formulation, kept for the self check
    Fq e = FQ_MOD;
    e.l[0] -= 2;
    Fq acc = FQ_ONE;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = fq_mul(acc, acc);
            if ((e.l[i] >> b) & 1) acc = fq_mul(acc, a);
        }
    return acc;
}

// ---- Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), coefficients low degree first --------------------------------
struct Fq12 {
    Fq c[12];
    bool operator==(const Fq12& o) const { return !memcmp(c, o.c, sizeof c); }
};
inline Fq12 fq12_zero() { Fq12 r; for (auto& x : r.c) x = FQ_ZERO; return r; }
inline Fq12 fq12_scalar(const Fq& x) { Fq12 r = fq12_zero(); r.c[0] = x; return r; }
inline Fq12 fq12_one() { return fq12_scalar(FQ_ONE); }
inline Fq12 fq12_add(const Fq12& a, const Fq12& b) { Fq12 r; for (int i = 0; i < 12; i++) r.c[i] = fq_add(a.c[i], b.c[i]); return r; }
inline Fq12 fq12_sub(const Fq12& a, const Fq12& b) { Fq12 r; for (int i = 0; i < 12; i++) r.c[i] = fq_sub(a.c[i], b.c[i]); return r; }
inline Fq12 fq12_neg(const Fq12& a) { Fq12 r; for (int i = 0; i < 12; i++) r.c[i] = fq_neg(a.c[i]); return r; }
inline Fq12 fq12_mul_small(const Fq12& a, uint64_t k) { Fq kk = fq_from_u64(k); Fq12 r; for (int i = 0; i < 12; i++) r.c[i] = fq_mul(a.c[i], kk); return r; }
inline Fq12 fq12_mul(const Fq12& a, const Fq12& b) {
    static const Fq c82 = fq_from_u64(82), c18 = fq_from_u64(18);
    Fq t[23];
    for (auto& x : t) x = FQ_ZERO;
    for (int i = 0; i < 12; i++) {
        if (a.c[i].is_zero()) continue;
        for (int j = 0; j < 12; j++) t[i + j] = fq_add(t[i + j], fq_mul(a.c[i], b.c[j]));
    }
    for (int e = 22; e >= 12; e--) {           // w^12 = 18 w^6 - 82
        if (t[e].is_zero()) continue;
        t[e - 12] = fq_sub(t[e - 12], fq_mul(t[e], c82));
        t[e - 6] = fq_add(t[e - 6], fq_mul(t[e], c18));
    }
    Fq12 r;
    for (int i = 0; i < 12; i++) r.c[i] = t[i];
    return r;
}
// polynomial helpers for the inverse (extended Euclid over Fq[w])
inline int poly_deg(const std::vector<Fq>& p) {
    int d = (int)p.size() - 1;
    while (d > 0 && p[d].is_zero()) d--;
    return d;
}
inline std::vector<Fq> poly_rounded_div(const std::vector<Fq>& a, const std::vector<Fq>& b) {
    const int dega = poly_deg(a), degb = poly_deg(b);
    std::vector<Fq> temp = a, o(a.size(), FQ_ZERO);
    const Fq inv_lead = fq_inv(b[degb]);
    for (int i = dega - degb; i >= 0; i--) {
        const Fq q = fq_mul(temp[degb + i], inv_lead);
        o[i] = fq_add(o[i], q);
        for (int c = 0; c <= degb; c++) temp[c + i] = fq_sub(temp[c + i], fq_mul(b[c], q));
    }
    o.resize(poly_deg(o) + 1);
    return o;
}
inline Fq12 fq12_inv(const Fq12& a) {
    std::vector<Fq> lm(13, FQ_ZERO), hm(13, FQ_ZERO), low(13, FQ_ZERO), high(13, FQ_ZERO);
    lm[0] = FQ_ONE;
    for (int i = 0; i < 12; i++) low[i] = a.c[i];
    high[0] = fq_from_u64(82);
    high[6] = fq_neg(fq_from_u64(18));
    high[12] = FQ_ONE;
    while (poly_deg(low)) {
        std::vector<Fq> r = poly_rounded_div(high, low);
        r.resize(13, FQ_ZERO);
        std::vector<Fq> nm = hm, nw = high;
        for (int i = 0; i < 13; i++)
            for (int j = 0; j < 13 - i; j++) {
                nm[i + j] = fq_sub(nm[i + j], fq_mul(lm[i], r[j]));
                nw[i + j] = fq_sub(nw[i + j], fq_mul(low[i], r[j]));
            }
        hm = lm; high = low;
        lm = nm; low = nw;
    }
    const Fq li = fq_inv(low[0]);
    Fq12 r;
    for (int i = 0; i < 12; i++) r.c[i] = fq_mul(lm[i], li);
    return r;
}
inline Fq12 fq12_div(const Fq12& a, const Fq12& b) { return fq12_mul(a, fq12_inv(b)); }
// a^e, e given as little-endian 64-bit words
inline Fq12 fq12_pow(const Fq12& a, const uint64_t* e, int words) {
    Fq12 acc = fq12_one();
    bool started = false;
    for (int i = words - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            if (started) acc = fq12_mul(acc, acc);
            if ((e[i] >> b) & 1) { acc = started ? fq12_mul(acc, a) : a; started = true; }
        }
    return started ? acc : fq12_one();
}

// ---- points --------------------------------------------------------------------------------------------------
struct G1Aff { Fq x, y; bool inf; };                 // Montgomery coordinates
struct Fq2 { Fq c0, c1; };                           // c0 + c1 u, u^2 = -1
struct G2Aff { Fq2 x, y; bool inf; };
struct Pt12 { Fq12 x, y; };

// ---- Fq2 / G2 group law (affine): only used to derive [tau]_2 for the test SRS --------------------------
inline Fq2 fq2_add(const Fq2& a, const Fq2& b) { return Fq2{fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; }
inline Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return Fq2{fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; }
inline Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
    return Fq2{fq_sub(fq_mul(a.c0, b.c0), fq_mul(a.c1, b.c1)), fq_add(fq_mul(a.c0, b.c1), fq_mul(a.c1, b.c0))};
}
inline Fq2 fq2_inv(const Fq2& a) {
    const Fq d = fq_inv(fq_add(fq_mul(a.c0, a.c0), fq_mul(a.c1, a.c1)));
    return Fq2{fq_mul(a.c0, d), fq_neg(fq_mul(a.c1, d))};
}
inline bool fq2_eq(const Fq2& a, const Fq2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
inline Fq2 fq2_small(uint64_t k) { return Fq2{fq_from_u64(k), FQ_ZERO}; }
inline G2Aff g2_double(const G2Aff& p) {
    if (p.inf || (p.y.c0.is_zero() && p.y.c1.is_zero())) return G2Aff{Fq2{FQ_ZERO, FQ_ZERO}, Fq2{FQ_ZERO, FQ_ZERO}, true};
    const Fq2 m = fq2_mul(fq2_mul(fq2_small(3), fq2_mul(p.x, p.x)), fq2_inv(fq2_add(p.y, p.y)));
    const Fq2 nx = fq2_sub(fq2_mul(m, m), fq2_add(p.x, p.x));
    const Fq2 ny = fq2_sub(fq2_mul(m, fq2_sub(p.x, nx)), p.y);
    return G2Aff{nx, ny, false};
}
inline G2Aff g2_add(const G2Aff& a, const G2Aff& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    if (fq2_eq(a.x, b.x)) {
        if (fq2_eq(a.y, b.y)) return g2_double(a);
        return G2Aff{Fq2{FQ_ZERO, FQ_ZERO}, Fq2{FQ_ZERO, FQ_ZERO}, true};
    }
    const Fq2 m = fq2_mul(fq2_sub(b.y, a.y), fq2_inv(fq2_sub(b.x, a.x)));
    const Fq2 nx = fq2_sub(fq2_sub(fq2_mul(m, m), a.x), b.x);
    const Fq2 ny = fq2_sub(fq2_mul(m, fq2_sub(a.x, nx)), a.y);
    return G2Aff{nx, ny, false};
}
static const Fq G2_GEN_X0_CANON = {{0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL}};
static const Fq G2_GEN_X1_CANON = {{0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL}};
static const Fq G2_GEN_Y0_CANON = {{0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL}};
static const Fq G2_GEN_Y1_CANON = {{0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL}};
inline G2Aff g2_generator() {
    return G2Aff{Fq2{fq_to_mont(G2_GEN_X0_CANON), fq_to_mont(G2_GEN_X1_CANON)},
                 Fq2{fq_to_mont(G2_GEN_Y0_CANON), fq_to_mont(G2_GEN_Y1_CANON)}, false};
}
// k * P, k a canonical 256-bit integer (little-endian words)
inline G2Aff g2_mul(const G2Aff& p, const uint64_t k[4]) {
    G2Aff acc{Fq2{FQ_ZERO, FQ_ZERO}, Fq2{FQ_ZERO, FQ_ZERO}, true};
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = g2_double(acc);
            if ((k[i] >> b) & 1) acc = g2_add(acc, p);
        }
    return acc;
}

inline Pt12 twist(const G2Aff& q) {
    // u = w^6 - 9:  x0 + x1 u = (x0 - 9 x1) + x1 w^6; then scale by w^2 (x) and w^3 (y)
    const Fq nine = fq_from_u64(9);
    Fq12 nx = fq12_zero(), ny = fq12_zero();
    nx.c[2] = fq_sub(q.x.c0, fq_mul(nine, q.x.c1));
    nx.c[8] = q.x.c1;
    ny.c[3] = fq_sub(q.y.c0, fq_mul(nine, q.y.c1));
    ny.c[9] = q.y.c1;
    return Pt12{nx, ny};
}
inline Pt12 pt12_double(const Pt12& p) {
    const Fq12 m = fq12_div(fq12_mul_small(fq12_mul(p.x, p.x), 3), fq12_mul_small(p.y, 2));
    const Fq12 nx = fq12_sub(fq12_mul(m, m), fq12_mul_small(p.x, 2));
    const Fq12 ny = fq12_sub(fq12_mul(m, fq12_sub(p.x, nx)), p.y);
    return Pt12{nx, ny};
}
inline Pt12 pt12_add(const Pt12& a, const Pt12& b) {
    if (a.x == b.x) return pt12_double(a);           // the loop never meets a == -b
    const Fq12 m = fq12_div(fq12_sub(b.y, a.y), fq12_sub(b.x, a.x));
    const Fq12 nx = fq12_sub(fq12_sub(fq12_mul(m, m), a.x), b.x);
    const Fq12 ny = fq12_sub(fq12_mul(m, fq12_sub(a.x, nx)), a.y);
    return Pt12{nx, ny};
}
inline Fq12 linefunc(const Pt12& p1, const Pt12& p2, const Pt12& t) {
    if (!(p1.x == p2.x)) {
        const Fq12 m = fq12_div(fq12_sub(p2.y, p1.y), fq12_sub(p2.x, p1.x));
        return fq12_sub(fq12_mul(m, fq12_sub(t.x, p1.x)), fq12_sub(t.y, p1.y));
    }
    if (p1.y == p2.y) {
        const Fq12 m = fq12_div(fq12_mul_small(fq12_mul(p1.x, p1.x), 3), fq12_mul_small(p1.y, 2));
        return fq12_sub(fq12_mul(m, fq12_sub(t.x, p1.x)), fq12_sub(t.y, p1.y));
    }
    return fq12_sub(t.x, p1.x);
}

// (p^12 - 1) / r and the ate loop count, little-endian 64-bit words (gen_pairing_consts.py)
#include "pairing_consts.h"
static const int LOG_ATE_LOOP_COUNT = 63;

// p as an exponent (little-endian words of the canonical modulus)
inline Fq12 fq12_frobenius(const Fq12& a) { return fq12_pow(a, FQ_MOD.l, 4); }

inline Fq12 miller_loop(const G2Aff& q2, const G1Aff& p1) {
    const Pt12 Q = twist(q2), P = Pt12{fq12_scalar(p1.x), fq12_scalar(p1.y)};
    Pt12 R = Q;
    Fq12 f = fq12_one();
    for (int i = LOG_ATE_LOOP_COUNT; i >= 0; i--) {
        f = fq12_mul(fq12_mul(f, f), linefunc(R, R, P));
        R = pt12_double(R);
        if ((ATE_LOOP_COUNT_LO >> i) & 1) {          // bit 64 of 6t+2, the leading one, is R = Q above
            f = fq12_mul(f, linefunc(R, Q, P));
            R = pt12_add(R, Q);
        }
    }
    const Pt12 Q1 = Pt12{fq12_frobenius(Q.x), fq12_frobenius(Q.y)};
    const Pt12 nQ2 = Pt12{fq12_frobenius(Q1.x), fq12_neg(fq12_frobenius(Q1.y))};
    f = fq12_mul(f, linefunc(R, Q1, P));
    R = pt12_add(R, Q1);
    f = fq12_mul(f, linefunc(R, nQ2, P));
    return f;
}

inline Fq12 final_exponentiate_reference(const Fq12& f) { return fq12_pow(f, FINAL_EXP_WORDS, FINAL_EXP_NWORDS); }

// ---- the production path: same pairing, ~40x fewer field operations ------------------------------------------
// miller_loop / final_exponentiate_reference above are the plain construction oracle/pairing.py restates (Fq12 curve
// arithmetic with Fq12 inversions, one 2790-bit exponentiation).  Below: the point Q stays on the twist over Fq2
// (affine, one Fq inversion per step), a line is the sparse element  -yP + (lambda xP) w + (yA - lambda xA) w^3,
// the Frobenius maps are coefficient-wise products with powers of gamma_1 = xi^((p-1)/6), and the final exponentiation
// is (p^6 - 1)(p^2 + 1) by Frobenius / conjugation followed by the Devegili et al. chain for (p^4 - p^2 + 1)/r
// (three exponentiations by the 63-bit BN parameter u).  miller_loop_fast returns bit-for-bit what miller_loop
// returns, and final_exponentiate what final_exponentiate_reference returns (tests/test_pairing_cpu.py).
inline Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, fq_neg(a.c1)}; }
inline Fq2 fq2_pow(const Fq2& a, const uint64_t* e, int words) {
    Fq2 acc{FQ_ONE, FQ_ZERO};
    for (int i = words - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = fq2_mul(acc, acc);
            if ((e[i] >> b) & 1) acc = fq2_mul(acc, a);
        }
    return acc;
}
// a + b u placed at w^pos:  (a - 9b) w^pos + b w^(pos+6), pos + 6 < 12
inline void fq12_add_fq2_at(Fq12& f, const Fq2& v, int pos) {
    static const Fq nine = fq_from_u64(9);
    f.c[pos] = fq_add(f.c[pos], fq_sub(v.c0, fq_mul(nine, v.c1)));
    f.c[pos + 6] = fq_add(f.c[pos + 6], v.c1);
}
struct PairingConsts {
    Fq2 g12, g13;            // gamma_1^2, gamma_1^3: Frobenius of a twisted G2 point
    Fq12 frob[12];           // gamma_1^i w^i: f^p = sum_i c_i frob[i]
    PairingConsts() {
        const Fq2 xi{fq_from_u64(9), FQ_ONE};
        const Fq2 g1 = fq2_pow(xi, FROB_EXP_WORDS, 4);
        g12 = fq2_mul(g1, g1);
        g13 = fq2_mul(g12, g1);
        Fq2 gi{FQ_ONE, FQ_ZERO};
        Fq12 wi = fq12_one(), w = fq12_zero();
        w.c[1] = FQ_ONE;
        for (int i = 0; i < 12; i++) {
            Fq12 emb = fq12_zero();
            fq12_add_fq2_at(emb, gi, 0);
            frob[i] = fq12_mul(emb, wi);
            gi = fq2_mul(gi, g1);
            wi = fq12_mul(wi, w);
        }
    }
};
inline const PairingConsts& pairing_consts() { static const PairingConsts c; return c; }

inline Fq12 fq12_frobenius_fast(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 r = fq12_zero();
    for (int i = 0; i < 12; i++) {
        if (f.c[i].is_zero()) continue;
        for (int j = 0; j < 12; j++)
            if (!pc.frob[i].c[j].is_zero()) r.c[j] = fq_add(r.c[j], fq_mul(f.c[i], pc.frob[i].c[j]));
    }
    return r;
}
inline Fq12 fq12_conj(const Fq12& f) {               // f^(p^6): w -> -w
    Fq12 r = f;
    for (int i = 1; i < 12; i += 2) r.c[i] = fq_neg(r.c[i]);
    return r;
}

// one Miller step on the twist: A <- A + B (or 2A when B is A), returns the line through them evaluated at P
inline Fq12 line_and_step(G2Aff& A, const G2Aff& B, const G1Aff& P) {
    Fq12 line = fq12_zero();
    Fq2 lambda;
    if (!fq2_eq(A.x, B.x)) {
        lambda = fq2_mul(fq2_sub(B.y, A.y), fq2_inv(fq2_sub(B.x, A.x)));
    } else if (fq2_eq(A.y, B.y)) {
        lambda = fq2_mul(fq2_mul(fq2_small(3), fq2_mul(A.x, A.x)), fq2_inv(fq2_add(A.y, A.y)));
    } else {                                         // vertical line xP - xA w^2 (A + B = identity): not met in the loop
        line.c[0] = P.x;
        fq12_add_fq2_at(line, Fq2{fq_neg(A.x.c0), fq_neg(A.x.c1)}, 2);
        A.inf = true;
        return line;
    }
    line.c[0] = fq_neg(P.y);
    fq12_add_fq2_at(line, Fq2{fq_mul(lambda.c0, P.x), fq_mul(lambda.c1, P.x)}, 1);
    fq12_add_fq2_at(line, fq2_sub(A.y, fq2_mul(lambda, A.x)), 3);
    const Fq2 nx = fq2_sub(fq2_sub(fq2_mul(lambda, lambda), A.x), B.x);
    const Fq2 ny = fq2_sub(fq2_mul(lambda, fq2_sub(A.x, nx)), A.y);
    A.x = nx;
    A.y = ny;
    return line;
}

inline Fq12 miller_loop_fast(const G2Aff& Q, const G1Aff& P) {
    const PairingConsts& pc = pairing_consts();
    G2Aff R = Q;
    Fq12 f = fq12_one();
    for (int i = LOG_ATE_LOOP_COUNT; i >= 0; i--) {
        const Fq12 l = line_and_step(R, R, P);
        f = fq12_mul(l, fq12_mul(f, f));             // the sparse operand first: fq12_mul skips its zero coefficients
        if ((ATE_LOOP_COUNT_LO >> i) & 1) f = fq12_mul(line_and_step(R, Q, P), f);
    }
    const G2Aff Q1{fq2_mul(fq2_conj(Q.x), pc.g12), fq2_mul(fq2_conj(Q.y), pc.g13), false};
    G2Aff nQ2{fq2_mul(fq2_conj(Q1.x), pc.g12), fq2_mul(fq2_conj(Q1.y), pc.g13), false};
    nQ2.y = Fq2{fq_neg(nQ2.y.c0), fq_neg(nQ2.y.c1)};
    f = fq12_mul(line_and_step(R, Q1, P), f);
    f = fq12_mul(line_and_step(R, nQ2, P), f);
    return f;
}

inline Fq12 final_exponentiate(const Fq12& in) {
    // easy part: in^((p^6 - 1)(p^2 + 1))
    Fq12 t1 = fq12_mul(fq12_conj(in), fq12_inv(in));
    t1 = fq12_mul(fq12_frobenius_fast(fq12_frobenius_fast(t1)), t1);
    // hard part (p^4 - p^2 + 1)/r = p^3 + (6u^2 + 1) p^2 + (-36u^3 - 18u^2 - 12u + 1) p + (-36u^3 - 30u^2 - 18u - 2)
    const uint64_t u[1] = {BN_U};
    const Fq12 fp = fq12_frobenius_fast(t1), fp2 = fq12_frobenius_fast(fp), fp3 = fq12_frobenius_fast(fp2);
    const Fq12 fu = fq12_pow(t1, u, 1), fu2 = fq12_pow(fu, u, 1), fu3 = fq12_pow(fu2, u, 1);
    const Fq12 fu2p = fq12_frobenius_fast(fu2), fu3p = fq12_frobenius_fast(fu3);
    const Fq12 y0 = fq12_mul(fq12_mul(fp, fp2), fp3);
    const Fq12 y1 = fq12_conj(t1);
    const Fq12 y2 = fq12_frobenius_fast(fu2p);
    const Fq12 y3 = fq12_conj(fq12_frobenius_fast(fu));
    const Fq12 y4 = fq12_conj(fq12_mul(fu, fu2p));
    const Fq12 y5 = fq12_conj(fu2);
    const Fq12 y6 = fq12_conj(fq12_mul(fu3, fu3p));
    Fq12 t0 = fq12_mul(fq12_mul(fq12_mul(y6, y6), y4), y5);
    Fq12 s1 = fq12_mul(fq12_mul(y3, y5), t0);
    t0 = fq12_mul(t0, y2);
    s1 = fq12_mul(fq12_mul(s1, s1), t0);
    s1 = fq12_mul(s1, s1);
    t0 = fq12_mul(s1, y1);
    s1 = fq12_mul(s1, y0);
    t0 = fq12_mul(t0, t0);
    return fq12_mul(t0, s1);
}

// prod_i e(P_i, Q_i) == 1  (identity points contribute 1)
inline bool pairing_product_is_one(const G1Aff* p, const G2Aff* q, int count, bool reference_construction = false) {
    Fq12 f = fq12_one();
    for (int i = 0; i < count; i++) {
        if (p[i].inf || q[i].inf) continue;
        f = fq12_mul(f, reference_construction ? miller_loop(q[i], p[i]) : miller_loop_fast(q[i], p[i]));
    }
    return (reference_construction ? final_exponentiate_reference(f) : final_exponentiate(f)) == fq12_one();
}

} }  // namespace zkfhe::host
