// BN254 G1 (y^2 = x^3 + 3 over Fq) for sm_100a device code.
//
// Replaces the curve arithmetic inside halo2curves `bn256::G1` that
// halo2-axiom's `best_multiexp` runs on the CPU [UPSTREAM, un-vendored;
// SURVEY.md §8 a19].  Affine points use halo2curves' in-memory layout
// (x||y, Montgomery, identity = (0,0)).  Accumulators use extended Jacobian
// "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): a mixed add costs
// 8M+2S and needs no inversion; identity is ZZ = 0.
#pragma once
#include "ff.cuh"

namespace zkfhe {

struct alignas(32) g1_affine {
    fq_t x, y;
};
struct alignas(32) g1_xyzz {
    fq_t x, y, zz, zzz;
};

__device__ __forceinline__ bool is_identity(const g1_affine& p) { return is_zero(p.x) && is_zero(p.y); }
__device__ __forceinline__ bool is_identity(const g1_xyzz& p) { return is_zero(p.zz); }

__device__ __forceinline__ g1_xyzz xyzz_identity() {
    g1_xyzz r;
    r.x = fe_zero<FQ>(); r.y = fe_zero<FQ>(); r.zz = fe_zero<FQ>(); r.zzz = fe_zero<FQ>();
    return r;
}
__device__ __forceinline__ g1_xyzz xyzz_from_affine(const g1_affine& p) {
    if (is_identity(p)) return xyzz_identity();
    g1_xyzz r;
    r.x = p.x; r.y = p.y; r.zz = fe_one<FQ>(); r.zzz = fe_one<FQ>();
    return r;
}

__device__ __forceinline__ g1_affine affine_load(const g1_affine* p) {
    g1_affine r;
    r.x = fe_load_nc(&p->x);
    r.y = fe_load_nc(&p->y);
    return r;
}
__device__ __forceinline__ void affine_store(g1_affine* p, const g1_affine& a) {
    fe_store(&p->x, a.x);
    fe_store(&p->y, a.y);
}
__device__ __forceinline__ g1_xyzz xyzz_load(const g1_xyzz* p) {
    g1_xyzz r;
    r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz);
    return r;
}
__device__ __forceinline__ void xyzz_store(g1_xyzz* p, const g1_xyzz& a) {
    fe_store(&p->x, a.x); fe_store(&p->y, a.y); fe_store(&p->zz, a.zz); fe_store(&p->zzz, a.zzz);
}

// 2 * (affine p), p != identity  (mdbl-2008-s-1, a = 0)
__device__ __forceinline__ g1_xyzz xyzz_dbl_affine(const g1_affine& p) {
    g1_xyzz r;
    fq_t u = dbl(p.y);
    fq_t v = sqr(u);
    fq_t w = mul(u, v);
    fq_t s = mul(p.x, v);
    fq_t xx = sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * p  (dbl-2008-s-1, a = 0)
__device__ __forceinline__ g1_xyzz xyzz_dbl(const g1_xyzz& p) {
    if (is_identity(p)) return p;
    g1_xyzz r;
    fq_t u = dbl(p.y);
    fq_t v = sqr(u);
    fq_t w = mul(u, v);
    fq_t s = mul(p.x, v);
    fq_t xx = sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = mul(v, p.zz);
    r.zzz = mul(w, p.zzz);
    return r;
}

// acc += (neg ? -p : p), p affine  (madd-2008-s)
__device__ __forceinline__ void xyzz_madd(g1_xyzz& acc, const g1_affine& p_in, bool negate) {
    if (is_identity(p_in)) return;
    g1_affine p = p_in;
    if (negate) p.y = neg(p.y);
    if (is_identity(acc)) {
        acc.x = p.x; acc.y = p.y; acc.zz = fe_one<FQ>(); acc.zzz = fe_one<FQ>();
        return;
    }
    fq_t u2 = mul(p.x, acc.zz);
    fq_t s2 = mul(p.y, acc.zzz);
    fq_t pp_ = sub(u2, acc.x);
    fq_t r = sub(s2, acc.y);
    if (is_zero(pp_)) {
        if (is_zero(r)) acc = xyzz_dbl_affine(p);
        else acc = xyzz_identity();
        return;
    }
    fq_t pp = sqr(pp_);
    fq_t ppp = mul(pp_, pp);
    fq_t q = mul(acc.x, pp);
    fq_t x3 = sub(sub(sqr(r), ppp), dbl(q));
    acc.y = sub(mul(r, sub(q, x3)), mul(acc.y, ppp));
    acc.x = x3;
    acc.zz = mul(acc.zz, pp);
    acc.zzz = mul(acc.zzz, ppp);
}

// acc += p  (add-2008-s)
__device__ __forceinline__ void xyzz_add(g1_xyzz& acc, const g1_xyzz& p) {
    if (is_identity(p)) return;
    if (is_identity(acc)) { acc = p; return; }
    fq_t u1 = mul(acc.x, p.zz);
    fq_t u2 = mul(p.x, acc.zz);
    fq_t s1 = mul(acc.y, p.zzz);
    fq_t s2 = mul(p.y, acc.zzz);
    fq_t pp_ = sub(u2, u1);
    fq_t r = sub(s2, s1);
    if (is_zero(pp_)) {
        if (is_zero(r)) acc = xyzz_dbl(acc);
        else acc = xyzz_identity();
        return;
    }
    fq_t pp = sqr(pp_);
    fq_t ppp = mul(pp_, pp);
    fq_t q = mul(u1, pp);
    fq_t x3 = sub(sub(sqr(r), ppp), dbl(q));
    acc.y = sub(mul(r, sub(q, x3)), mul(s1, ppp));
    acc.x = x3;
    acc.zz = mul(mul(acc.zz, p.zz), pp);
    acc.zzz = mul(mul(acc.zzz, p.zzz), ppp);
}

// k * p for a small plain integer k (double-and-add, variable time)
__device__ inline g1_xyzz xyzz_mul_small(const g1_xyzz& p, uint32_t k) {
    g1_xyzz acc = xyzz_identity();
    for (int bit = 31; bit >= 0; bit--) {
        acc = xyzz_dbl(acc);
        if ((k >> bit) & 1) xyzz_add(acc, p);
    }
    return acc;
}

// XYZZ -> affine (one Fermat inversion); identity -> (0,0)
__device__ inline g1_affine xyzz_to_affine(const g1_xyzz& p) {
    g1_affine r;
    if (is_identity(p)) { r.x = fe_zero<FQ>(); r.y = fe_zero<FQ>(); return r; }
    // 1/ZZZ, then 1/ZZ = ZZ^2 / ZZZ^2 ... cheaper: i = 1/(ZZ*ZZZ); 1/ZZ = i*ZZZ; 1/ZZZ = i*ZZ
    fq_t i = inv(mul(p.zz, p.zzz));
    r.x = mul(p.x, mul(i, p.zzz));
    r.y = mul(p.y, mul(i, p.zz));
    return r;
}

}  // namespace zkfhe
