// Host Poseidon permutation (t = 5, R_F = 8, R_P = 60, x^5 over BN254 Fr) on AVX-512 IFMA.
//
// The Fiat-Shamir sponge of the reference's transcript (snark-verifier `PoseidonTranscript`) is
// sequential and stays on the host: ~1,960 permutations per config-1 proof, 1,281 of them for the public instances.
// The scalar form (host_ff.h: mulx / adcx / adox, sparse partial rounds, lazily reduced dot products) is bound by
// the integer multiplier's issue rate at ~16 us per permutation although its dependent chain is only ~7 us.  Here
// the products that are NOT on that chain move to the vector unit:
//   * field elements in radix 2^52 (5 limbs), eight per zmm register set, Montgomery products with R' = 2^260 by
//     vpmadd52luq / vpmadd52huq (operand scanning, one reduction step per limb, lazily reduced sums of products);
//     values stay loosely reduced (< 2^258) -- a product of a loosely reduced value and a constant comes out < 1.3 r;
//   * full rounds entirely in vectors: the five S-boxes are three products on five lanes (the Montgomery factors of
//     x^2, x^4, x^5 drift to 2^252, 2^244, 2^240; the MDS columns are stored pre-multiplied by 2^276 so that the lazily
//     reduced matrix-vector product lands back on 2^256), the matrix is one five-term dot product on five lanes;
//   * partial rounds (sparse form: s0' = (s0 + k)^5; s0 <- m00 s0' + <v, s[1..4]>; s_j <- s_j + w_j s0'): the S-box
//     chain and m00 s0' stay scalar (that IS the dependent chain: four products); the row product <v, s[1..4]> is one
//     vector product plus a horizontal sum and the column update one more, both fed by the PREVIOUS round's s0' (see
//     partial_rounds), so they run under the scalar chain.
// Constants are derived once from poseidon_consts.h (same tables as the scalar form) and the result is bit-identical
// to host::poseidon_permute_scalar; tests/test_oracle_transcript.py holds both against the oracle.
//
// Compiled by g++ with -mavx512f -mavx512ifma -mavx512vl (this file only); callers dispatch on cpuid.
#include <immintrin.h>
#include <sched.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

// This translation unit is compiled with AVX-512 enabled.  host_ff.h is all `inline` functions, i.e. COMDAT symbols of
// which the linker keeps ONE copy for the whole library: a copy compiled here could carry EVEX-encoded instructions into
// code that runs on CPUs without AVX-512.  So the header is included inside an unnamed namespace (its system headers are
// included above, at global scope): every helper it defines has internal linkage in this file, and the two exported
// functions below take the library's `zkfhe::host::Fr` by pointer and reinterpret it as this file's identical copy.
#define ZKFHE_POSEIDON_IFMA_TU 1
namespace ifma_tu {
namespace {
#include "host_ff.h"
}
}

namespace ifma_tu { namespace { namespace zkfhe { namespace host {

// every CPU with AVX-512 IFMA has BMI2 / ADX (poseidon_ifma_available checks both): no per-call dispatch in here
inline Fr fmul(const Fr& a, const Fr& b) { return mul_adx(a, b); }
inline Fr fsqr(const Fr& a) { return mul_adx(a, a); }

constexpr uint64_t MASK52 = (1ull << 52) - 1;
struct alignas(64) V5 { __m512i l[5]; };

inline void to52(const Fr& a, uint64_t o[5]) {
    o[0] = a.l[0] & MASK52;
    o[1] = ((a.l[0] >> 52) | (a.l[1] << 12)) & MASK52;
    o[2] = ((a.l[1] >> 40) | (a.l[2] << 24)) & MASK52;
    o[3] = ((a.l[2] >> 28) | (a.l[3] << 36)) & MASK52;
    o[4] = a.l[3] >> 16;
}
// sum_k limb[k] 2^(52 k) with limbs < 2^62 and a total < 2^260, reduced mod r.  The quotient estimate comes from the
// top 64 bits (v >> 196) times a 64-bit reciprocal of (r >> 196) + 1: never too large, at most two too small, so two
// conditional subtractions finish it; no division, no data-dependent branch (this runs once per partial round).
inline Fr from52_reduce(const uint64_t t[5]) {
    uint64_t w[5];
    u128 acc = (u128)t[0] + ((u128)t[1] << 52);
    w[0] = (uint64_t)acc; acc >>= 64;
    acc += (u128)t[2] << 40;
    w[1] = (uint64_t)acc; acc >>= 64;
    acc += (u128)t[3] << 28;
    w[2] = (uint64_t)acc; acc >>= 64;
    acc += (u128)t[4] << 16;
    w[3] = (uint64_t)acc;
    w[4] = (uint64_t)(acc >> 64);
    const uint64_t top = (w[4] << 60) | (w[3] >> 4);
    static const uint64_t recip = (uint64_t)(((u128)1 << 121) / ((FR_MOD.l[3] >> 4) + 1));      // rt > 2^57, so this fits
    const uint64_t q = (uint64_t)(((u128)top * recip) >> 64) >> 57;
    {
        u128 m = 0;
        uint64_t brw = 0;
        for (int i = 0; i < 4; i++) {
            m += (u128)FR_MOD.l[i] * q;
            const uint64_t sub = (uint64_t)m;
            m >>= 64;
            const u128 d = (u128)w[i] - sub - brw;
            w[i] = (uint64_t)d;
            brw = (uint64_t)(d >> 64) & 1;
        }
        w[4] = w[4] - (uint64_t)m - brw;
    }
    for (int pass = 0; pass < 2; pass++) {           // v < 3 r here: subtract r while that does not go negative
        uint64_t d[5], brw = 0;
        for (int i = 0; i < 4; i++) {
            const u128 x = (u128)w[i] - FR_MOD.l[i] - brw;
            d[i] = (uint64_t)x;
            brw = (uint64_t)(x >> 64) & 1;
        }
        d[4] = w[4] - brw;
        const uint64_t keep = (uint64_t)0 - (d[4] >> 63);        // all ones: the difference is negative, keep v
        for (int i = 0; i < 5; i++) w[i] = (w[i] & keep) | (d[i] & ~keep);
    }
    return Fr{{w[0], w[1], w[2], w[3]}};
}

struct Consts {
    __m512i N[5], NINV;
    V5 c_first[4], c_second[4];        // lanes 0..4: round constants of the full rounds (Montgomery 2^256, as added)
    V5 mds_col[5];                     // column j of the MDS matrix on lanes 0..4, times 2^276
    V5 mlast_col[5];                   // column j of the last partial round's dense matrix, times 2^260
    V5 vrow[POSEIDON_RP - 1];          // lanes 0..3: first-row entries for s1..s4, lane 4: beta_r = <v_r, w_{r-1}>; times 2^260
    V5 wcol[POSEIDON_RP - 1];          // lanes 0..3: first-column entries for s1..s4, times 2^260
    Fr m00[POSEIDON_RP - 1];
    __m512i bidx[5];                   // broadcast-lane-j index vectors
};

inline void set_lane(V5& v, int lane, const Fr& a) {
    uint64_t t[5];
    to52(a, t);
    for (int k = 0; k < 5; k++) {
        alignas(64) uint64_t w[8];
        _mm512_store_si512(w, v.l[k]);
        w[lane] = t[k];
        v.l[k] = _mm512_load_si512(w);
    }
}
inline V5 zero_v5() {
    V5 v;
    for (int k = 0; k < 5; k++) v.l[k] = _mm512_setzero_si512();
    return v;
}

const Consts& consts() {
    static Consts* C = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        Consts* c = (Consts*)aligned_alloc(64, (sizeof(Consts) + 63) / 64 * 64);
        uint64_t t[5];
        to52(FR_MOD, t);
        for (int k = 0; k < 5; k++) c->N[k] = _mm512_set1_epi64((long long)t[k]);
        uint64_t inv = 1;                                   // -r^-1 mod 2^52 (Newton on the low limb)
        for (int i = 0; i < 6; i++) inv *= 2 - FR_MOD.l[0] * inv;
        c->NINV = _mm512_set1_epi64((long long)((0 - inv) & MASK52));
        const Fr s4 = from_u64(1ull << 4), s20 = from_u64(1ull << 20);
        for (int r = 0; r < 4; r++) {
            c->c_first[r] = zero_v5();
            c->c_second[r] = zero_v5();
            for (int i = 0; i < 5; i++) {
                set_lane(c->c_first[r], i, pc(POSEIDON_C_FIRST, r * 5 + i));
                set_lane(c->c_second[r], i, pc(POSEIDON_C_SECOND, r * 5 + i));
            }
        }
        for (int j = 0; j < 5; j++) {
            c->mds_col[j] = zero_v5();
            c->mlast_col[j] = zero_v5();
            for (int i = 0; i < 5; i++) {
                set_lane(c->mds_col[j], i, mul(pc(POSEIDON_MDS, i * 5 + j), s20));
                set_lane(c->mlast_col[j], i, mul(pc(POSEIDON_M_LAST, i * 5 + j), s4));
            }
            c->bidx[j] = _mm512_set1_epi64(j);
        }
        for (int r = 0; r + 1 < POSEIDON_RP; r++) {
            const size_t b = (size_t)r * 9;
            c->m00[r] = pc(POSEIDON_SPARSE, b);
            c->vrow[r] = zero_v5();
            c->wcol[r] = zero_v5();
            Fr beta = FR_ZERO;
            for (int j = 1; j < 5; j++) {
                set_lane(c->vrow[r], j - 1, mul(pc(POSEIDON_SPARSE, b + j), s4));
                set_lane(c->wcol[r], j - 1, mul(pc(POSEIDON_SPARSE, b + 4 + j), s4));
                if (r) beta = add(beta, mul(pc(POSEIDON_SPARSE, b + j), pc(POSEIDON_SPARSE, b - 9 + 4 + j)));
            }
            set_lane(c->vrow[r], 4, mul(beta, s4));
        }
        C = c;
    });
    return *C;
}

#define MADLO(acc, a, b) acc = _mm512_madd52lo_epu64(acc, a, b)
#define MADHI(acc, a, b) acc = _mm512_madd52hi_epu64(acc, a, b)

struct Acc { __m512i t[6]; };
inline void acc_zero(Acc& A) { for (int k = 0; k < 6; k++) A.t[k] = _mm512_setzero_si512(); }
// A += a * bi  (bi = one limb of b on every lane)
inline void acc_mul(Acc& A, const V5& a, __m512i bi) {
    MADLO(A.t[0], a.l[0], bi); MADLO(A.t[1], a.l[1], bi); MADLO(A.t[2], a.l[2], bi); MADLO(A.t[3], a.l[3], bi); MADLO(A.t[4], a.l[4], bi);
    MADHI(A.t[1], a.l[0], bi); MADHI(A.t[2], a.l[1], bi); MADHI(A.t[3], a.l[2], bi); MADHI(A.t[4], a.l[3], bi); MADHI(A.t[5], a.l[4], bi);
}
// one Montgomery reduction step: make the low limb vanish, shift down by one limb
inline void acc_reduce(Acc& A, const Consts& c) {
    const __m512i m = _mm512_madd52lo_epu64(_mm512_setzero_si512(), A.t[0], c.NINV);
    MADLO(A.t[0], m, c.N[0]); MADLO(A.t[1], m, c.N[1]); MADLO(A.t[2], m, c.N[2]); MADLO(A.t[3], m, c.N[3]); MADLO(A.t[4], m, c.N[4]);
    MADHI(A.t[1], m, c.N[0]); MADHI(A.t[2], m, c.N[1]); MADHI(A.t[3], m, c.N[2]); MADHI(A.t[4], m, c.N[3]); MADHI(A.t[5], m, c.N[4]);
    A.t[0] = _mm512_add_epi64(A.t[1], _mm512_srli_epi64(A.t[0], 52));
    A.t[1] = A.t[2]; A.t[2] = A.t[3]; A.t[3] = A.t[4]; A.t[4] = A.t[5];
    A.t[5] = _mm512_setzero_si512();
}
// limbs < 2^52 again (the value is < 2^260, so nothing leaves the top limb)
inline V5 normalise(const __m512i t[5]) {
    const __m512i M = _mm512_set1_epi64((long long)MASK52);
    V5 o;
    __m512i c = _mm512_srli_epi64(t[0], 52);
    o.l[0] = _mm512_and_si512(t[0], M);
    __m512i x = _mm512_add_epi64(t[1], c);
    c = _mm512_srli_epi64(x, 52); o.l[1] = _mm512_and_si512(x, M);
    x = _mm512_add_epi64(t[2], c);
    c = _mm512_srli_epi64(x, 52); o.l[2] = _mm512_and_si512(x, M);
    x = _mm512_add_epi64(t[3], c);
    c = _mm512_srli_epi64(x, 52); o.l[3] = _mm512_and_si512(x, M);
    o.l[4] = _mm512_add_epi64(t[4], c);
    return o;
}
inline V5 vadd_norm(const V5& a, const V5& b) {
    __m512i t[5];
    for (int k = 0; k < 5; k++) t[k] = _mm512_add_epi64(a.l[k], b.l[k]);
    return normalise(t);
}
// a * b * 2^-260 on every lane (raw accumulator limbs)
inline Acc mont_raw(const V5& a, const V5& b, const Consts& c) {
    Acc A;
    acc_zero(A);
#pragma GCC unroll 5
    for (int i = 0; i < 5; i++) {
        acc_mul(A, a, b.l[i]);
        acc_reduce(A, c);
    }
    return A;
}
inline V5 mont(const V5& a, const V5& b, const Consts& c) { return normalise(mont_raw(a, b, c).t); }
// sum_j col[j] * (lane j of x broadcast) * 2^-260: a 5 x 5 matrix times the vector held on lanes 0..4 of x
inline V5 matvec(const V5 col[5], const V5& x, const Consts& c) {
    Acc A;
    acc_zero(A);
#pragma GCC unroll 5
    for (int i = 0; i < 5; i++) {
#pragma GCC unroll 5
        for (int j = 0; j < 5; j++) acc_mul(A, col[j], _mm512_permutexvar_epi64(c.bidx[j], x.l[i]));
        acc_reduce(A, c);
    }
    return normalise(A.t);
}
inline V5 sbox_scaled(const V5& x, const Consts& c) {        // x^5 with the Montgomery factor 2^240 (see the header comment)
    const V5 x2 = mont(x, x, c);
    const V5 x4 = mont(x2, x2, c);
    return mont(x4, x, c);
}
__attribute__((noinline)) V5 full_round(const V5& s, const V5& rc, const Consts& c) {
    return matvec(c.mds_col, sbox_scaled(vadd_norm(s, rc), c), c);
}
inline Fr lane_to_fr(const V5& v, int lane) {
    uint64_t t[5];
    for (int k = 0; k < 5; k++) {
        alignas(64) uint64_t w[8];
        _mm512_store_si512(w, v.l[k]);
        t[k] = w[lane];
    }
    return from52_reduce(t);
}
inline V5 bcast_fr(const Fr& a) {
    uint64_t t[5];
    to52(a, t);
    V5 v;
    for (int k = 0; k < 5; k++) v.l[k] = _mm512_set1_epi64((long long)t[k]);
    return v;
}

// Partial rounds 0 .. R_P-2 (sparse form) on (s0 scalar, s1..s4 on lanes 0..3 of S).
// The row product lags the column update by one round: with S_r = S_{r-1} + w_{r-1} y_{r-1},
//     <v_r, S_r> = <v_r, S_{r-1}> + beta_r y_{r-1},     beta_r = <v_r, w_{r-1}> a constant (fifth lane of vrow[r]),
// so neither vector product of a round waits for the other and both only need the previous round's s0'.  S is not
// reduced on the way: it grows by < 1.02 r per round, 61.3 r < 2^260 = 84.6 r after all 59.
// Program order matters (the out-of-order window holds ~1.5 scalar products): vector work sits between the products.
// Measured on the B200 box's host (profiles/r02_host_microbench_ifma.txt, us per permutation): scalar form 16.0; this
// arrangement 9.75; the same without the lag (row product, S-box, then a two-term product for the column update) 10.0;
// with m00 s0' computed as (m00 x) x^4 beside the squarings (a three-product chain, five products in all) 10.5 -- the
// loop is bound by instruction throughput (~1,400 per round), not by its dependent chain.
inline void partial_rounds(Fr& s0, V5& S, const Consts& c) {
    Fr y_prev = FR_ZERO;                                                     // s0' of the previous round
    for (int r = 0; r + 1 < POSEIDON_RP; r++) {
        const Fr x = add(s0, pc(POSEIDON_K, r));
        const V5 yv = bcast_fr(y_prev);
        const Fr x2 = fsqr(x);
        V5 B;
        for (int k = 0; k < 5; k++) B.l[k] = _mm512_mask_blend_epi64(0x10, S.l[k], yv.l[k]);
        const Acc rest_raw = mont_raw(c.vrow[r], B, c);                      // lanes 0..3: v_j s_j, lane 4: beta y_prev
        const Fr x4 = fsqr(x2);
        if (r) S = vadd_norm(S, mont(c.wcol[r - 1], yv, c));                 // the column update of the previous round
        uint64_t rl[5];
        for (int k = 0; k < 5; k++) rl[k] = (uint64_t)_mm512_reduce_add_epi64(rest_raw.t[k]);
        const Fr rest = from52_reduce(rl);
        y_prev = fmul(x4, x);
        s0 = add(fmul(c.m00[r], y_prev), rest);
    }
    S = vadd_norm(S, mont(c.wcol[POSEIDON_RP - 2], bcast_fr(y_prev), c));
}

inline void permute(Fr s[POSEIDON_T]) {
    const Consts& c = consts();
    V5 X = zero_v5();
    for (int i = 0; i < 5; i++) set_lane(X, i, s[i]);
    for (int r = 0; r < 4; r++) X = full_round(X, c.c_first[r], c);
    Fr s0 = lane_to_fr(X, 0);
    V5 S;
    for (int k = 0; k < 5; k++) S.l[k] = _mm512_maskz_alignr_epi64(0x0f, X.l[k], X.l[k], 1);     // lanes 1..4 -> 0..3, rest zero
    partial_rounds(s0, S, c);
    {   // last partial round: dense matrix on (s0', s1..s4)
        const Fr y = pow5(add(s0, pc(POSEIDON_K, POSEIDON_RP - 1)));
        const V5 yv = bcast_fr(y);
        for (int k = 0; k < 5; k++) {
            const __m512i up = _mm512_maskz_alignr_epi64(0x1e, S.l[k], S.l[k], 7);                // lanes 0..3 -> 1..4
            X.l[k] = _mm512_mask_blend_epi64(0x01, up, yv.l[k]);
        }
        X = matvec(c.mlast_col, X, c);
    }
    for (int r = 0; r < 4; r++) X = full_round(X, c.c_second[r], c);
    for (int i = 0; i < 5; i++) s[i] = lane_to_fr(X, i);
}

} } } }  // namespace ifma_tu::(unnamed)::zkfhe::host

namespace zkfhe { namespace host {

struct Fr;                     // the library's type (host_ff.h): four little-endian 64-bit limbs, as the copy above

// Same function as host::poseidon_permute_scalar (bit-identical output), inputs and outputs canonical Montgomery.
void poseidon_permute_ifma(Fr s[5]) { ifma_tu::zkfhe::host::permute(reinterpret_cast<ifma_tu::zkfhe::host::Fr*>(s)); }

bool poseidon_ifma_available() {
    static const bool ok = [] {
        if (const char* e = getenv("ZKFHE_POSEIDON_IFMA")) if (atoi(e) == 0) return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512vl") &&
               __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx");
    }();
    return ok;
}

} }  // namespace zkfhe::host
