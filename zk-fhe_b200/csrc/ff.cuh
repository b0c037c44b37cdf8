// BN254 field arithmetic for sm_100a device code: Fr (scalar field) and Fq
// (base field), 8 x 32-bit limbs, Montgomery form with R = 2^256 -- the
// in-memory layout of halo2curves' `bn256::{Fr,Fq}` (4 x u64 little-endian
// limbs; reference call sites read F::MODULUS at src/poly_chip.rs:90,135,158,199)
// so host buffers pass through the C ABI without conversion.
//
// The hot ops (mul/sqr/add/sub) are generated inline-PTX blocks
// (ff_ptx_gen.cuh, see gen_ff_ptx.py); `mul_c` is an independent plain-C
// Montgomery product used only by the on-device self test.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ff_ptx_gen.cuh"
#include "inv_bin.cuh"

namespace zkfhe {

enum FieldId { FR = 0, FQ = 1 };

template <int F>
struct alignas(32) fe {
    uint32_t v[8];
};
using fr_t = fe<FR>;
using fq_t = fe<FQ>;

template <int F> struct fconst;
template <> struct fconst<FR> {
    static __device__ __forceinline__ fe<FR> mod() { return fe<FR>{ZKFHE_FR_MOD}; }
    static __device__ __forceinline__ fe<FR> one() { return fe<FR>{ZKFHE_FR_ONE}; }
    static __device__ __forceinline__ fe<FR> r2() { return fe<FR>{ZKFHE_FR_R2}; }
    static __device__ __forceinline__ fe<FR> r3() { return fe<FR>{ZKFHE_FR_R3}; }
    static __device__ __forceinline__ fe<FR> mod_minus_2() { return fe<FR>{ZKFHE_FR_MOD_MINUS_2}; }
    static constexpr uint32_t inv32 = ZKFHE_FR_INV32;
};
template <> struct fconst<FQ> {
    static __device__ __forceinline__ fe<FQ> mod() { return fe<FQ>{ZKFHE_FQ_MOD}; }
    static __device__ __forceinline__ fe<FQ> one() { return fe<FQ>{ZKFHE_FQ_ONE}; }
    static __device__ __forceinline__ fe<FQ> r2() { return fe<FQ>{ZKFHE_FQ_R2}; }
    static __device__ __forceinline__ fe<FQ> r3() { return fe<FQ>{ZKFHE_FQ_R3}; }
    static __device__ __forceinline__ fe<FQ> mod_minus_2() { return fe<FQ>{ZKFHE_FQ_MOD_MINUS_2}; }
    static constexpr uint32_t inv32 = ZKFHE_FQ_INV32;
};

template <int F> __device__ __forceinline__ fe<F> fe_zero() {
    fe<F> z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = 0;
    return z;
}
template <int F> __device__ __forceinline__ fe<F> fe_one() { return fconst<F>::one(); }

__device__ __forceinline__ fr_t mul(const fr_t& a, const fr_t& b) { fr_t r; ptx::fr_mul(r.v, a.v, b.v); return r; }
__device__ __forceinline__ fq_t mul(const fq_t& a, const fq_t& b) { fq_t r; ptx::fq_mul(r.v, a.v, b.v); return r; }
__device__ __forceinline__ fr_t sqr(const fr_t& a) { fr_t r; ptx::fr_sqr(r.v, a.v); return r; }
__device__ __forceinline__ fq_t sqr(const fq_t& a) { fq_t r; ptx::fq_sqr(r.v, a.v); return r; }
__device__ __forceinline__ fr_t add(const fr_t& a, const fr_t& b) { fr_t r; ptx::fr_add(r.v, a.v, b.v); return r; }
__device__ __forceinline__ fq_t add(const fq_t& a, const fq_t& b) { fq_t r; ptx::fq_add(r.v, a.v, b.v); return r; }
__device__ __forceinline__ fr_t sub(const fr_t& a, const fr_t& b) { fr_t r; ptx::fr_sub(r.v, a.v, b.v); return r; }
__device__ __forceinline__ fq_t sub(const fq_t& a, const fq_t& b) { fq_t r; ptx::fq_sub(r.v, a.v, b.v); return r; }

template <int F> __device__ __forceinline__ fe<F> operator*(const fe<F>& a, const fe<F>& b) { return mul(a, b); }
template <int F> __device__ __forceinline__ fe<F> operator+(const fe<F>& a, const fe<F>& b) { return add(a, b); }
template <int F> __device__ __forceinline__ fe<F> operator-(const fe<F>& a, const fe<F>& b) { return sub(a, b); }

template <int F> __device__ __forceinline__ bool is_zero(const fe<F>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
template <int F> __device__ __forceinline__ bool eq(const fe<F>& a, const fe<F>& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
template <int F> __device__ __forceinline__ fe<F> neg(const fe<F>& a) {
    return is_zero(a) ? a : sub(fconst<F>::mod(), a);   // mod - a is already reduced for a != 0
}
template <int F> __device__ __forceinline__ fe<F> dbl(const fe<F>& a) { return add(a, a); }

// canonical integer (8 LE limbs) <-> Montgomery
template <int F> __device__ __forceinline__ fe<F> to_mont(const fe<F>& a) { return mul(a, fconst<F>::r2()); }
template <int F> __device__ __forceinline__ fe<F> from_mont(const fe<F>& a) {
    fe<F> one_plain = fe_zero<F>();
    one_plain.v[0] = 1;
    return mul(a, one_plain);
}

// a^e, e given as 8 plain LE limbs (not secret: variable time).
// The exponentiation helpers are cold (setup, one inversion per column / per row at most): they stay
// out of line with rolled loops so that ptxas sees one copy of the product per helper, not hundreds
// (fully inlined they made single kernels of 10^5 PTX lines and a 15-minute build).
template <int F> __device__ __noinline__ fe<F> pow_limbs(const fe<F> a, const fe<F> e) {
    fe<F> acc = fe_one<F>();
    bool started = false;
#pragma unroll 1
    for (int i = 7; i >= 0; i--) {
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = sqr(acc);
            if ((e.v[i] >> bit) & 1) {
                acc = started ? mul(acc, a) : a;
                started = true;
            }
        }
    }
    return acc;
}
// Fermat inverse; inv(0) = 0.  Kept as the independent cross-check of `inv` in the device self test.
template <int F> __device__ __forceinline__ fe<F> inv_fermat(const fe<F>& a) { return pow_limbs(a, fconst<F>::mod_minus_2()); }
// Inverse of a Montgomery-form element: binary extended Euclid on the raw limbs gives (aR)^-1,
// one product with R^3 brings it back to a^-1 R.  inv(0) = 0.
template <int F> __device__ __noinline__ fe<F> inv(const fe<F> a) {
    fe<F> t;
    const fe<F> p = fconst<F>::mod();
    u256_inv_odd(t.v, a.v, p.v);
    return mul(t, fconst<F>::r3());
}

template <int F> __device__ __noinline__ fe<F> pow_u64(const fe<F> a, unsigned long long e) {
    fe<F> acc = fe_one<F>();
    fe<F> base = a;
#pragma unroll 1
    while (e) {
        if (e & 1) acc = mul(acc, base);
        e >>= 1;
        if (e) base = sqr(base);
    }
    return acc;
}

// 32-byte global / shared accesses as two 128-bit transactions
template <int F> __device__ __forceinline__ fe<F> fe_load(const fe<F>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    fe<F> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <int F> __device__ __forceinline__ fe<F> fe_load_nc(const fe<F>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    fe<F> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <int F> __device__ __forceinline__ void fe_store(fe<F>* p, const fe<F>& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// Independent plain-C Montgomery product (CIOS on 32-bit limbs with 64-bit
// accumulators).  Self-test only: checks the generated PTX on the real chip.
template <int F> __device__ fe<F> mul_c(const fe<F>& a, const fe<F>& b) {
    const fe<F> p = fconst<F>::mod();
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
            t[j] = (uint32_t)s;
            c = s >> 32;
        }
        uint64_t s = (uint64_t)t[8] + c;
        t[8] = (uint32_t)s;
        t[9] = (uint32_t)(s >> 32);
        uint32_t m = t[0] * fconst<F>::inv32;
        c = ((uint64_t)m * p.v[0] + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            uint64_t s2 = (uint64_t)m * p.v[j] + t[j] + c;
            t[j - 1] = (uint32_t)s2;
            c = s2 >> 32;
        }
        s = (uint64_t)t[8] + c;
        t[7] = (uint32_t)s;
        t[8] = t[9] + (uint32_t)(s >> 32);
    }
    // conditional subtract
    uint32_t d[8];
    uint64_t brw = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t s = (uint64_t)t[j] - p.v[j] - brw;
        d[j] = (uint32_t)s;
        brw = (s >> 32) & 1;
    }
    bool ge = t[8] != 0 || brw == 0;
    fe<F> r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = ge ? d[j] : t[j];
    return r;
}

}  // namespace zkfhe
