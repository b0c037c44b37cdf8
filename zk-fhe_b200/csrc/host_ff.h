// Host-side BN254 Fr arithmetic (4 x u64 Montgomery, the ABI layout) and the Poseidon
// Fiat-Shamir transcript.  Only the sequential, tiny part of the prover runs here: hashing
// commitments into challenges and a few hundred scalar operations of the opening argument;
// everything proportional to the domain size runs on the GPU.
#pragma once
#include <sched.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "poseidon_consts.h"

namespace zkfhe { namespace host {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t l[4];
    bool operator==(const Fr& o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

static const Fr FR_MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const Fr FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const Fr FR_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr FR_ZERO = {{0, 0, 0, 0}};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL;
// canonical (non-Montgomery) constants
static const Fr FR_ROOT_OF_UNITY_CANON = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const Fr FR_DELTA_CANON = {{0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL}};
static const Fr FR_ZETA_CANON = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};

inline bool geq(const Fr& a, const Fr& b) {
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] > b.l[i]) return true;
        if (a.l[i] < b.l[i]) return false;
    }
    return true;
}
inline Fr sub_raw(const Fr& a, const Fr& b) {
    Fr r;
    u128 brw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - brw;
        r.l[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
    return r;
}
// t (< 2r, with `hi` = the carry out of bit 256) reduced once: t - r if that does not borrow, else t.  Branch-free.
inline Fr reduce_once(const Fr& t, uint64_t hi = 0) {
    Fr d;
    unsigned char brw = 0;
    unsigned long long x;
    brw = __builtin_usubll_overflow(t.l[0], FR_MOD.l[0], &x); d.l[0] = x;
    for (int i = 1; i < 4; i++) {
        unsigned long long y;
        unsigned char b1 = __builtin_usubll_overflow(t.l[i], FR_MOD.l[i], &y);
        unsigned char b2 = __builtin_usubll_overflow(y, (unsigned long long)brw, &x);
        d.l[i] = x;
        brw = b1 | b2;
    }
    const uint64_t keep = (uint64_t)0 - (uint64_t)(brw && !hi);      // all ones: t < r, keep t
    Fr r;
    for (int i = 0; i < 4; i++) r.l[i] = (t.l[i] & keep) | (d.l[i] & ~keep);
    return r;
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr t;
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    return reduce_once(t, (uint64_t)c);
}
inline Fr sub(const Fr& a, const Fr& b) { return geq(a, b) ? sub_raw(a, b) : sub_raw(FR_MOD, sub_raw(b, a)); }
inline Fr neg(const Fr& a) { return a.is_zero() ? a : sub_raw(FR_MOD, a); }
// Montgomery product, CIOS with the "no-carry" shortcut (valid because the top limb of r is below 2^63 - 1: the
// running sum never needs a sixth limb), fully unrolled: the transcript's Poseidon sponge is ~2 M products per proof
// and is the only host arithmetic on the critical path.
inline Fr mul_portable(const Fr& a, const Fr& b) {
    const uint64_t q0 = FR_MOD.l[0], q1 = FR_MOD.l[1], q2 = FR_MOD.l[2], q3 = FR_MOD.l[3];
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define ZK_MONT_ROUND(bi)                                                             \
    {                                                                                 \
        u128 A = (u128)a.l[0] * (bi) + t0;                                            \
        const uint64_t m = (uint64_t)A * FR_INV;                                      \
        u128 C = (u128)m * q0 + (uint64_t)A;                                          \
        A = (u128)a.l[1] * (bi) + t1 + (uint64_t)(A >> 64);                           \
        C = (u128)m * q1 + (uint64_t)A + (uint64_t)(C >> 64);                         \
        t0 = (uint64_t)C;                                                             \
        A = (u128)a.l[2] * (bi) + t2 + (uint64_t)(A >> 64);                           \
        C = (u128)m * q2 + (uint64_t)A + (uint64_t)(C >> 64);                         \
        t1 = (uint64_t)C;                                                             \
        A = (u128)a.l[3] * (bi) + t3 + (uint64_t)(A >> 64);                           \
        C = (u128)m * q3 + (uint64_t)A + (uint64_t)(C >> 64);                         \
        t2 = (uint64_t)C;                                                             \
        t3 = (uint64_t)(C >> 64) + (uint64_t)(A >> 64);                               \
    }
    ZK_MONT_ROUND(b.l[0]) ZK_MONT_ROUND(b.l[1]) ZK_MONT_ROUND(b.l[2]) ZK_MONT_ROUND(b.l[3])
#undef ZK_MONT_ROUND
    Fr o = {{t0, t1, t2, t3}};
    return reduce_once(o);
}
#if defined(__x86_64__) && defined(__GNUC__)
// The same product with BMI2 / ADX (mulx + the two independent carry chains of adcx / adox): ~130 instructions instead
// of the ~330 gcc makes of the portable form, i.e. ~3x the throughput.  Five accumulator registers rotate through the
// four rounds (after a round's reduction the low limb is exactly zero and becomes the next round's top limb).
#define ZK_MM_MULADD(src, off, T0, T1, T2, T3, A)                                        \
    "xorl %%eax, %%eax\n\t"                                                              \
    "mulxq " #off "+0(%[" #src "]), %%r13, %%r14\n\t adoxq %%r13, " T0 "\n\t adcxq %%r14, " T1 "\n\t"  \
    "mulxq " #off "+8(%[" #src "]), %%r13, %%r14\n\t adoxq %%r13, " T1 "\n\t adcxq %%r14, " T2 "\n\t"  \
    "mulxq " #off "+16(%[" #src "]), %%r13, %%r14\n\t adoxq %%r13, " T2 "\n\t adcxq %%r14, " T3 "\n\t" \
    "mulxq " #off "+24(%[" #src "]), %%r13, %%r14\n\t adoxq %%r13, " T3 "\n\t adcxq %%r14, " A "\n\t"  \
    "adoxq %%rax, " A "\n\t"
#define ZK_MM_ROUND(boff, T0, T1, T2, T3, A)                                             \
    "movq " #boff "(%[b]), %%rdx\n\t"                                                    \
    ZK_MM_MULADD(a, 0, T0, T1, T2, T3, A)                                                \
    "movq " T0 ", %%rdx\n\t imulq %[inv], %%rdx\n\t"                                     \
    ZK_MM_MULADD(q, 0, T0, T1, T2, T3, A)
inline Fr mul_adx(const Fr& a, const Fr& b) {
    Fr o;
    __asm__(
        "xorl %%r8d, %%r8d\n\t xorl %%r9d, %%r9d\n\t xorl %%r10d, %%r10d\n\t xorl %%r11d, %%r11d\n\t xorl %%r12d, %%r12d\n\t"
        ZK_MM_ROUND(0, "%%r8", "%%r9", "%%r10", "%%r11", "%%r12")
        ZK_MM_ROUND(8, "%%r9", "%%r10", "%%r11", "%%r12", "%%r8")
        ZK_MM_ROUND(16, "%%r10", "%%r11", "%%r12", "%%r8", "%%r9")
        ZK_MM_ROUND(24, "%%r11", "%%r12", "%%r8", "%%r9", "%%r10")
        // result in (r12, r8, r9, r10); subtract r once if that does not borrow
        "movq %%r12, %%r13\n\t movq %%r8, %%r14\n\t movq %%r9, %%rax\n\t movq %%r10, %%rdx\n\t"
        "subq 0(%[q]), %%r13\n\t sbbq 8(%[q]), %%r14\n\t sbbq 16(%[q]), %%rax\n\t sbbq 24(%[q]), %%rdx\n\t"
        "cmovcq %%r12, %%r13\n\t cmovcq %%r8, %%r14\n\t cmovcq %%r9, %%rax\n\t cmovcq %%r10, %%rdx\n\t"
        "movq %%r13, 0(%[o])\n\t movq %%r14, 8(%[o])\n\t movq %%rax, 16(%[o])\n\t movq %%rdx, 24(%[o])\n\t"
        :
        : [a] "r"(a.l), [b] "r"(b.l), [q] "r"(FR_MOD.l), [inv] "r"(FR_INV), [o] "r"(o.l)
        : "rax", "rdx", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "cc", "memory");
    return o;
}
#undef ZK_MM_ROUND
#undef ZK_MM_MULADD
// sum_{j<5} m[j] * s[j] (Montgomery) with ONE reduction: the five 512-bit products are accumulated in eight registers
// (5 r^2 < r * 2^256, so the sum is a valid REDC input and the result is < 2r) -- 96 mulx instead of 160 for the rows of
// the MDS / sparse matrices, which are 60 % of the permutation's products.  (text generated by a ten-line script: 24 rows of mulx + adox/adcx with carry ripple).
static const uint64_t FR_ZERO_WORD = 0;
inline Fr dot5_adx(const Fr* m, const Fr* s) {
    Fr o;
    __asm__ volatile(
        "xorl %%r8d, %%r8d\n\t"
        "xorl %%r9d, %%r9d\n\t"
        "xorl %%r10d, %%r10d\n\t"
        "xorl %%r11d, %%r11d\n\t"
        "xorl %%r12d, %%r12d\n\t"
        "xorl %%r13d, %%r13d\n\t"
        "xorl %%r14d, %%r14d\n\t"
        "xorl %%r15d, %%r15d\n\t"
        "movq 0(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 0(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq 8(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 16(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 24(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 8(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 0(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 8(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 16(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 24(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 16(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 0(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 8(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 16(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 24(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 24(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 0(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 8(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 16(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq 24(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 32(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 32(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq 40(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 48(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 56(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 40(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 32(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 40(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 48(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 56(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 48(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 32(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 40(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 48(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 56(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 56(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 32(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 40(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 48(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq 56(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 64(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 64(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq 72(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 80(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 88(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 72(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 64(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 72(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 80(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 88(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 80(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 64(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 72(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 80(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 88(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 88(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 64(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 72(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 80(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq 88(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 96(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 96(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq 104(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 112(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 120(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 104(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 96(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 104(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 112(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 120(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 112(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 96(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 104(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 112(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 120(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 120(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 96(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 104(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 112(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq 120(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 128(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 128(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq 136(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 144(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 152(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 136(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 128(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq 136(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 144(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 152(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 144(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 128(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq 136(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 144(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 152(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq 152(%[s]), %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq 128(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq 136(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq 144(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq 152(%[m]), %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq %%r8, %%rdx\n\t imulq %[inv], %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq %[q0], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r8\n\t"
        "adcxq %%rbx, %%r9\n\t"
        "mulxq %[q1], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq %[q2], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq %[q3], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "adoxq %[zero], %%r12\n\t"
        "adcxq %[zero], %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq %%r9, %%rdx\n\t imulq %[inv], %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq %[q0], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r9\n\t"
        "adcxq %%rbx, %%r10\n\t"
        "mulxq %[q1], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq %[q2], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq %[q3], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "adoxq %[zero], %%r13\n\t"
        "adcxq %[zero], %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq %%r10, %%rdx\n\t imulq %[inv], %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq %[q0], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r10\n\t"
        "adcxq %%rbx, %%r11\n\t"
        "mulxq %[q1], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq %[q2], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq %[q3], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "adoxq %[zero], %%r14\n\t"
        "adcxq %[zero], %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq %%r11, %%rdx\n\t imulq %[inv], %%rdx\n\t"
        "xorl %%eax, %%eax\n\t"
        "mulxq %[q0], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r11\n\t"
        "adcxq %%rbx, %%r12\n\t"
        "mulxq %[q1], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r12\n\t"
        "adcxq %%rbx, %%r13\n\t"
        "mulxq %[q2], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r13\n\t"
        "adcxq %%rbx, %%r14\n\t"
        "mulxq %[q3], %%rax, %%rbx\n\t"
        "adoxq %%rax, %%r14\n\t"
        "adcxq %%rbx, %%r15\n\t"
        "adoxq %[zero], %%r15\n\t"
        "movq %%r12, %%r8\n\t"
        "movq %%r13, %%r9\n\t"
        "movq %%r14, %%r10\n\t"
        "movq %%r15, %%r11\n\t"
        "subq %[q0], %%r8\n\t"
        "sbbq %[q1], %%r9\n\t"
        "sbbq %[q2], %%r10\n\t"
        "sbbq %[q3], %%r11\n\t"
        "cmovcq %%r12, %%r8\n\t"
        "cmovcq %%r13, %%r9\n\t"
        "cmovcq %%r14, %%r10\n\t"
        "cmovcq %%r15, %%r11\n\t"
        "movq %%r8, 0(%[o])\n\t"
        "movq %%r9, 8(%[o])\n\t"
        "movq %%r10, 16(%[o])\n\t"
        "movq %%r11, 24(%[o])\n\t"
        :
        : [m] "r"(m), [s] "r"(s), [o] "r"(o.l), [q0] "m"(FR_MOD.l[0]), [q1] "m"(FR_MOD.l[1]), [q2] "m"(FR_MOD.l[2]), [q3] "m"(FR_MOD.l[3]),
          [inv] "m"(FR_INV), [zero] "m"(FR_ZERO_WORD)
        : "rax", "rbx", "rdx", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "r15", "cc", "memory");
    return o;
}
inline bool cpu_has_adx() {
    static const bool ok = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx");
    return ok;
}
inline Fr mul(const Fr& a, const Fr& b) { return cpu_has_adx() ? mul_adx(a, b) : mul_portable(a, b); }
#define ZKFHE_HAVE_DOT5_ADX 1
#else
inline Fr mul(const Fr& a, const Fr& b) { return mul_portable(a, b); }
#endif
inline Fr sqr(const Fr& a) { return mul(a, a); }
inline Fr to_mont(const Fr& canon) { return mul(canon, FR_R2); }
inline Fr from_mont(const Fr& m) { Fr one = {{1, 0, 0, 0}}; return mul(m, one); }
inline Fr from_u64(uint64_t v) { Fr c = {{v, 0, 0, 0}}; return to_mont(c); }
inline Fr pow_u64(Fr a, uint64_t e) {
    Fr acc = FR_ONE;
    while (e) {
        if (e & 1) acc = mul(acc, a);
        a = sqr(a);
        e >>= 1;
    }
    return acc;
}
inline Fr inv(const Fr& a) {   // Fermat; inv(0) = 0
    Fr e = FR_MOD;
    e.l[0] -= 2;
    Fr acc = FR_ONE;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = sqr(acc);
            if ((e.l[i] >> b) & 1) acc = mul(acc, a);
        }
    return acc;
}
inline Fr omega(uint32_t k) {   // generator of the 2^k-th roots of unity (Montgomery)
    Fr w = to_mont(FR_ROOT_OF_UNITY_CANON);
    for (uint32_t s = k; s < 28; s++) w = sqr(w);
    return w;
}

// ---- Poseidon permutation (t = 5, full rounds 8, partial rounds 60, x^5) ---------------------
// Optimised form (tables derived and checked against the plain form by gen_poseidon.py): the partial rounds add one
// scalar and multiply by a sparse matrix -- 12 field products per partial round instead of 28, 1056 per permutation
// instead of 2000.  `poseidon_permute_plain` is the textbook form over (RC, MDS); the CPU tests hold both against the
// oracle (which is pinned by the published t = 3 / t = 2 vectors).
inline const Fr& pc(const uint64_t (*tbl)[4], size_t i) { return *(const Fr*)tbl[i]; }
inline Fr pow5(const Fr& x) { Fr x2 = sqr(x); return mul(sqr(x2), x); }
// <row of five matrix entries, state>: one lazily reduced dot product where the CPU has BMI2 / ADX
static_assert(POSEIDON_T == 5, "dot5 is written for the t = 5 instance");
inline Fr dot5(const uint64_t (*m)[4], size_t first, const Fr s[POSEIDON_T]) {
#ifdef ZKFHE_HAVE_DOT5_ADX
    if (cpu_has_adx()) return dot5_adx((const Fr*)m[first], s);
#endif
    Fr acc = mul(pc(m, first), s[0]);
    for (int j = 1; j < POSEIDON_T; j++) acc = add(acc, mul(pc(m, first + j), s[j]));
    return acc;
}
inline void poseidon_mds(Fr s[POSEIDON_T], const uint64_t (*m)[4]) {
    Fr n[POSEIDON_T];
    for (int i = 0; i < POSEIDON_T; i++) n[i] = dot5(m, (size_t)i * POSEIDON_T, s);
    for (int i = 0; i < POSEIDON_T; i++) s[i] = n[i];
}
// (measured on the B200 box's host and not kept: issuing the four state[0]-independent products of a partial round's dot
// product before the S-box chain, so that the out-of-order core overlaps them -- 16.3 us either way)
inline void poseidon_permute_scalar(Fr s[POSEIDON_T]) {
    const int T = POSEIDON_T, half = POSEIDON_RF / 2;
    for (int r = 0; r < half; r++) {
        for (int i = 0; i < T; i++) s[i] = pow5(add(s[i], pc(POSEIDON_C_FIRST, r * T + i)));
        poseidon_mds(s, POSEIDON_MDS);
    }
    for (int r = 0; r < POSEIDON_RP; r++) {
        if (r + 1 < POSEIDON_RP) {
            const size_t b = (size_t)r * (2 * T - 1);
            s[0] = pow5(add(s[0], pc(POSEIDON_K, r)));
            const Fr n0 = dot5(POSEIDON_SPARSE, b, s);
            for (int j = 1; j < T; j++) s[j] = add(s[j], mul(pc(POSEIDON_SPARSE, b + T - 1 + j), s[0]));
            s[0] = n0;
        } else {
            s[0] = pow5(add(s[0], pc(POSEIDON_K, r)));
            poseidon_mds(s, POSEIDON_M_LAST);
        }
    }
    for (int r = 0; r < half; r++) {
        for (int i = 0; i < T; i++) s[i] = pow5(add(s[i], pc(POSEIDON_C_SECOND, r * T + i)));
        poseidon_mds(s, POSEIDON_MDS);
    }
}
// AVX-512 IFMA form (poseidon_ifma.cpp, compiled separately with the vector flags): the products off the S-box chain
// run on the vector unit; bit-identical results.  `poseidon_permute` is what the transcript calls.
#ifndef ZKFHE_POSEIDON_IFMA_TU          // (poseidon_ifma.cpp includes this header with internal linkage and defines the two itself)
void poseidon_permute_ifma(Fr s[POSEIDON_T]);
bool poseidon_ifma_available();
inline void poseidon_permute(Fr s[POSEIDON_T]) {
#if defined(__x86_64__) && defined(__GNUC__)
    if (poseidon_ifma_available()) { poseidon_permute_ifma(s); return; }
#endif
    poseidon_permute_scalar(s);
}
#else
inline void poseidon_permute(Fr s[POSEIDON_T]) { poseidon_permute_scalar(s); }
#endif
inline void poseidon_permute_plain(Fr s[POSEIDON_T]) {
    const int T = POSEIDON_T, half = POSEIDON_RF / 2, rounds = POSEIDON_RF + POSEIDON_RP;
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < T; i++) s[i] = add(s[i], pc(POSEIDON_RC, r * T + i));
        const bool full = r < half || r >= half + POSEIDON_RP;
        for (int i = 0; i < (full ? T : 1); i++) s[i] = pow5(s[i]);
        poseidon_mds(s, POSEIDON_MDS);
    }
}

// ---- BLAKE2b-512 (RFC 7693), streaming, clonable ----------------------------------------------
struct Blake2b {
    uint64_t h[8];
    uint8_t buf[128];
    size_t buflen = 0;
    u128 total = 0;
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    Blake2b() {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        memcpy(h, iv, sizeof h);
        h[0] ^= 0x01010000ULL ^ 64;   // digest length 64, no key, fanout = depth = 1
    }
    void compress(const uint8_t* block, bool last) {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = iv[i]; }
        v[12] ^= (uint64_t)total;
        v[13] ^= (uint64_t)(total >> 64);
        if (last) v[14] = ~v[14];
#define ZK_B2G(a, b, c, d, x, y)                                                   \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); v[c] = v[c] + v[d];    \
    v[b] = rotr(v[b] ^ v[c], 24); v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = sigma[r];
            ZK_B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) ZK_B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
            ZK_B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) ZK_B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
            ZK_B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) ZK_B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
            ZK_B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) ZK_B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef ZK_B2G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = (const uint8_t*)data;
        while (len) {
            if (buflen == 128) {            // keep the last block for finalisation
                total += 128;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = 128 - buflen < len ? 128 - buflen : len;
            memcpy(buf + buflen, p, take);
            buflen += take; p += take; len -= take;
        }
    }
    void finalize(uint8_t out[64]) const {   // does not disturb the running state
        Blake2b c = *this;
        c.total += c.buflen;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, 64);
    }
};

static const Fr FR_R3 = {{0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL}};
// 64 uniform bytes (little-endian 512-bit integer) -> Fr (Montgomery), as halo2curves `from_uniform_bytes`
inline Fr from_uniform_bytes(const uint8_t b[64]) {
    Fr d0, d1;
    memcpy(d0.l, b, 32);
    memcpy(d1.l, b + 32, 32);
    return add(mul(d0, FR_R2), mul(d1, FR_R3));
}

// ---- the reference's own test trapdoor -----------------------------------------------------------
// halo2-scaffold `gen_srs(k)` without a params file: `ParamsKZG::setup(k, ChaCha20Rng::from_seed([0; 32]))`, whose
// first draw is tau = `Fr::random(rng)` = the first 64 keystream bytes (eight `next_u64`, little-endian) as a 512-bit
// integer reduced mod r [UPSTREAM-RECALL, SURVEY.md App. C.1].  ChaCha20 with an all-zero key, counter and nonce is
// RFC 7539 appendix A.1 test vector #1, which the CPU tests pin, so this tau (and with it g[i] = tau^i G of the
// `--insecure-test-srs` commitment key) is a function of published constants only.
inline void chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint8_t out[64]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5],
                       key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    auto rotl = [](uint32_t v, int n) { return (v << n) | (v >> (32 - n)); };
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int i = 0; i < 10; i++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) { const uint32_t v = x[i] + in[i]; memcpy(out + 4 * i, &v, 4); }
}
inline Fr reference_test_tau() {                       // Montgomery form
    const uint32_t zero_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t ks[64];
    chacha20_block(zero_key, 0, 0, ks);
    return from_uniform_bytes(ks);
}

// ---- transcript ------------------------------------------------------------------------------
// Two interchangeable Fiat-Shamir hashes over the same message sequence:
//   POSEIDON (default; what the reference's `prove` runs: snark-verifier-sdk `gen_snark_shplonk` ->
//            `PoseidonTranscript<NativeLoader>` over the PSE `poseidon` crate, t = 5, rate 4, R_F = 8, R_P = 60
//            [UPSTREAM-RECALL, SURVEY.md App. C.2]): a sponge over state[1..4] starting from state[0] = 2^64
//            (the crate's `State::default`); absorbed items queue up and a squeeze pads them with a single 1,
//            absorbs them rate-by-rate (an exact multiple of the rate gets one extra block holding only the 1) and
//            returns state[1].  Scalars are absorbed natively, points as two elements, x mod r and y mod r (the
//            transcript's `fe_to_fe` of the affine coordinates).  ~1,700 permutations per config-1 proof (1,281 of
//            them for the 5,121 public instances), on the host.
//   BLAKE2B  halo2's own `Blake2bWrite` / `Challenge255` shape: every absorbed item is fed to a running
//            BLAKE2b-512 with a one-byte tag; a challenge is the digest of the state so far (tag 0 appended),
//            reduced from 512 bits.  Microseconds per proof; selectable (kind 0).
// Every written item is also appended to the proof in canonical little-endian form (scalars 32 bytes; points 32 bytes
// compressed, see write_point), so the verifier replays the same sequence.
enum TranscriptKind { TRANSCRIPT_BLAKE2B = 0, TRANSCRIPT_POSEIDON = 1 };

struct Transcript {
    int kind;
    Blake2b b2;
    Fr state[POSEIDON_T];
    std::vector<Fr> buf;
    std::vector<uint8_t> proof;

    explicit Transcript(int kind_ = TRANSCRIPT_POSEIDON) : kind(kind_) {
        for (auto& s : state) s = FR_ZERO;
        const Fr two64 = {{0, 1, 0, 0}};
        state[0] = to_mont(two64);              // capacity element 2^64
        const char tag[] = "zkfhe-b200-transcript-v1";
        b2.update(tag, sizeof tag - 1);
    }
    void common_scalar(const Fr& x_mont) {
        if (kind == TRANSCRIPT_POSEIDON) { buf.push_back(x_mont); return; }
        Fr c = from_mont(x_mont);
        uint8_t t = 2;
        b2.update(&t, 1);
        b2.update(c.l, 32);
    }
    void write_scalar(const Fr& x_mont) {
        common_scalar(x_mont);
        Fr c = from_mont(x_mont);
        const uint8_t* b = (const uint8_t*)c.l;
        proof.insert(proof.end(), b, b + 32);
    }
    // canonical affine coordinates (4 x u64 each), identity = all zero
    void common_point(const uint64_t x_canon[4], const uint64_t y_canon[4]) {
        if (kind == TRANSCRIPT_POSEIDON) {
            const uint64_t* cs[2] = {x_canon, y_canon};
            for (auto c : cs) {                 // Fq coordinate -> Fr: p < 2r, so one conditional subtraction
                Fr v = {{c[0], c[1], c[2], c[3]}};
                if (geq(v, FR_MOD)) v = sub_raw(v, FR_MOD);
                buf.push_back(to_mont(v));
            }
            return;
        }
        uint8_t t = 1;
        b2.update(&t, 1);
        b2.update(x_canon, 32);
        b2.update(y_canon, 32);
    }
    // The proof carries points the way halo2 writes them (`to_bytes` of halo2curves' bn256 `G1Affine`, SURVEY.md App. C.2
    // [UPSTREAM-RECALL]): 32 bytes, x little-endian (x < p < 2^254 leaves the two top bits free), bit 6 of the last
    // byte = parity of y, bit 7 = the identity (whose other bits are zero).  The hash still absorbs both coordinates.
    static void compress_point(const uint64_t x_canon[4], const uint64_t y_canon[4], uint8_t out[32]) {
        memcpy(out, x_canon, 32);
        if ((x_canon[0] | x_canon[1] | x_canon[2] | x_canon[3] | y_canon[0] | y_canon[1] | y_canon[2] | y_canon[3]) == 0) out[31] |= 0x80;
        else out[31] |= (uint8_t)((y_canon[0] & 1) << 6);
    }
    void write_point(const uint64_t x_canon[4], const uint64_t y_canon[4]) {
        common_point(x_canon, y_canon);
        uint8_t c[32];
        compress_point(x_canon, y_canon, c);
        proof.insert(proof.end(), c, c + 32);
    }
    Fr squeeze() {
        if (kind == TRANSCRIPT_POSEIDON) {
            buf.push_back(FR_ONE);
            while (buf.size() % 4) buf.push_back(FR_ZERO);
            // A long absorb (the 5,121 instances: ~1,300 permutations, ~20 ms) is a compute-bound stretch on a host
            // thread; with many proofs in flight the other proofs' threads need the cores for microseconds at a time to
            // launch kernels.  Offering the core every ~0.5 ms keeps their wake-up latency (and the GPU's idle gaps)
            // short; it costs a system call per 32 permutations.  ZKFHE_TRANSCRIPT_YIELD=0 switches it off.
            static const bool yield_on = [] { const char* e = getenv("ZKFHE_TRANSCRIPT_YIELD"); return !e || atoi(e) != 0; }();
            const bool long_absorb = yield_on && buf.size() > 512;
            for (size_t i = 0; i < buf.size(); i += 4) {
                for (int j = 0; j < 4; j++) state[1 + j] = add(state[1 + j], buf[i + j]);
                poseidon_permute(state);
                if (long_absorb && (i & 127) == 124) sched_yield();
            }
            buf.clear();
            return state[1];
        }
        uint8_t t = 0;
        b2.update(&t, 1);
        uint8_t d[64];
        b2.finalize(d);
        return from_uniform_bytes(d);
    }
};

} }  // namespace zkfhe::host
