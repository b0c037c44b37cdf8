/* CPU oracle, C restatement.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restates, in plain C with OpenMP threads, the CPU algorithms the GPU path replaces:
 *   - /root/reference/src/poly.rs:75-103   Poly::mul            (O(N^2) schoolbook, exact integers)
 *   - /root/reference/src/poly.rs:180-191  Poly::reduce_by_modulus
 *   - /root/reference/src/poly.rs:113-177  Poly::divide_by_cyclo (literal long division)
 *   - halo2-axiom arithmetic::best_fft      [UPSTREAM-RECALL; un-vendored, version unpinned]
 *       bit-reversal, precomputed twiddles, radix-2 DIT stages, threads over butterflies
 *   - halo2-axiom arithmetic::best_multiexp [UPSTREAM-RECALL]
 *       points split across threads, serial Pippenger per chunk with window
 *       c = ceil(ln(chunk)), (256/c)+1 segments, running-sum bucket reduction
 *   - halo2curves bn256 Fr/Fq (4 x u64 Montgomery, R = 2^256) and G1 Jacobian formulas
 *
 * It is pinned by tests/test_oracle_c.py against the pure-Python oracle (which is itself
 * pinned by the reference's bfv.in / bfv.json known answers).  Used as the checker at
 * sizes where Python is too slow, and as bench.py's `cpu_baseline` ("port": a
 * restatement, not the reference binary, which cannot be built here -- no Rust toolchain).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

#define EXPORT __attribute__((visibility("default")))

/* ---- field constants (recomputed in tests against oracle/field.py) ---------------------- */
static const fe FR_MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const fe FQ_MOD = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL, FQ_INV = 0x87d20782e4866389ULL;
static const fe FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const fe FQ_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const fe FR_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fe FQ_ONE = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
/* 2^28-th root of unity of Fr (canonical): 7^((r-1)/2^28) */
static const fe FR_ROOT_CANON = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const fe FR_ZETA_CANON = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};

typedef struct { const fe* mod; uint64_t inv; const fe* r2; const fe* one; } field_t;
static const field_t FR = {&FR_MOD, 0xc2e1f593efffffffULL, &FR_R2, &FR_ONE};
static const field_t FQ = {&FQ_MOD, 0x87d20782e4866389ULL, &FQ_R2, &FQ_ONE};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int fe_geq(const fe* a, const fe* b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline void fe_sub_raw(fe* r, const fe* a, const fe* b) {
    u128 brw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - brw;
        r->l[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
}
static inline void f_add(const field_t* F, fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    fe t;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    if (c || fe_geq(&t, F->mod)) fe_sub_raw(r, &t, F->mod); else *r = t;
}
static inline void f_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
    if (fe_geq(a, b)) { fe_sub_raw(r, a, b); return; }
    fe t;
    fe_sub_raw(&t, b, a);
    fe_sub_raw(r, F->mod, &t);
}
static inline void f_neg(const field_t* F, fe* r, const fe* a) {
    if (fe_is_zero(a)) *r = *a; else fe_sub_raw(r, F->mod, a);
}
/* Montgomery product, CIOS on 64-bit limbs */
static inline void f_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t* p = F->mod->l;
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = ((u128)m * p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&o, F->mod)) fe_sub_raw(r, &o, F->mod); else *r = o;
}
static inline void f_sqr(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static inline void f_to_mont(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, F->r2); }
static inline void f_from_mont(const field_t* F, fe* r, const fe* a) {
    fe one = {{1, 0, 0, 0}};
    f_mul(F, r, a, &one);
}
static void f_pow(const field_t* F, fe* r, const fe* a, const fe* e /*plain*/) {
    fe acc = *F->one;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            f_sqr(F, &acc, &acc);
            if ((e->l[i] >> b) & 1) f_mul(F, &acc, &acc, a);
        }
    *r = acc;
}
static void f_inv(const field_t* F, fe* r, const fe* a) {
    fe e = *F->mod;
    e.l[0] -= 2;   /* moduli end in ...01 / ...47: no borrow */
    f_pow(F, r, a, &e);
}
static void f_pow_u64(const field_t* F, fe* r, const fe* a, uint64_t e) {
    fe acc = *F->one, base = *a;
    while (e) {
        if (e & 1) f_mul(F, &acc, &acc, &base);
        f_sqr(F, &base, &base);
        e >>= 1;
    }
    *r = acc;
}

/* ---- small exported field helpers (tests) ----------------------------------------------- */
EXPORT void orc_field_mul(int which, uint64_t* r, const uint64_t* a, const uint64_t* b) {
    f_mul(which ? &FQ : &FR, (fe*)r, (const fe*)a, (const fe*)b);
}
EXPORT void orc_field_inv(int which, uint64_t* r, const uint64_t* a) { f_inv(which ? &FQ : &FR, (fe*)r, (const fe*)a); }
EXPORT void orc_to_mont(int which, uint64_t* r, const uint64_t* a) { f_to_mont(which ? &FQ : &FR, (fe*)r, (const fe*)a); }
EXPORT void orc_from_mont(int which, uint64_t* r, const uint64_t* a) { f_from_mont(which ? &FQ : &FR, (fe*)r, (const fe*)a); }

EXPORT void orc_to_mont_array(int which, const uint64_t* in, uint64_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) f_to_mont(which ? &FQ : &FR, (fe*)(out + 4 * i), (const fe*)(in + 4 * i));
}

/* ---- NTT (halo2 best_fft shape) ---------------------------------------------------------- */
static uint32_t bitrev32(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

static void fr_omega(fe* w_mont, uint32_t log_n, int inverse) {
    fe w;
    f_to_mont(&FR, &w, &FR_ROOT_CANON);
    for (uint32_t s = log_n; s < 28; s++) f_sqr(&FR, &w, &w);
    if (inverse) f_inv(&FR, &w, &w);
    *w_mont = w;
}

/* data: batch columns of n Montgomery Fr elements; natural order in and out. */
EXPORT void orc_ntt(uint64_t* data_, uint32_t log_n, uint32_t batch, int inverse, int coset, int threads) {
    fe* data = (fe*)data_;
    const size_t n = (size_t)1 << log_n;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    fe w;
    fr_omega(&w, log_n, inverse);
    fe* tw = (fe*)malloc(sizeof(fe) * (n / 2 ? n / 2 : 1));
    tw[0] = FR_ONE;
    for (size_t i = 1; i < n / 2; i++) f_mul(&FR, &tw[i], &tw[i - 1], &w);
    fe zeta[3], zeta_inv[3], n_inv;
    zeta[0] = FR_ONE;
    f_to_mont(&FR, &zeta[1], &FR_ZETA_CANON);
    f_sqr(&FR, &zeta[2], &zeta[1]);
    zeta_inv[0] = FR_ONE; zeta_inv[1] = zeta[2]; zeta_inv[2] = zeta[1];
    {
        fe nn = {{(uint64_t)n, 0, 0, 0}}, nm;
        f_to_mont(&FR, &nm, &nn);
        f_inv(&FR, &n_inv, &nm);
    }
    for (uint32_t col = 0; col < batch; col++) {
        fe* a = data + (size_t)col * n;
        if (!inverse && coset) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) if (i % 3) f_mul(&FR, &a[i], &a[i], &zeta[i % 3]);
        }
        for (size_t k = 0; k < n; k++) {
            size_t rk = bitrev32((uint32_t)k, log_n);
            if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
        }
        for (uint32_t s = 0; s < log_n; s++) {
            const size_t m = (size_t)1 << s, step = (n / 2) >> s;
#pragma omp parallel for schedule(static)
            for (size_t b = 0; b < n / 2; b++) {
                size_t j = b & (m - 1), i0 = ((b >> s) << (s + 1)) + j, i1 = i0 + m;
                fe t, u = a[i0];
                f_mul(&FR, &t, &a[i1], &tw[j * step]);
                f_add(&FR, &a[i0], &u, &t);
                f_sub(&FR, &a[i1], &u, &t);
            }
        }
        if (inverse) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) {
                f_mul(&FR, &a[i], &a[i], &n_inv);
                if (coset && i % 3) f_mul(&FR, &a[i], &a[i], &zeta_inv[i % 3]);
            }
        }
    }
    free(tw);
}

/* ---- G1 Jacobian (halo2curves formulas: dbl-2009-l, add-2007-bl, madd-2007-bl) ----------- */
typedef struct { fe x, y; } aff;
typedef struct { fe x, y, z; } jac;

static inline int aff_is_id(const aff* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void jac_set_id(jac* p) { memset(p, 0, sizeof *p); }
static inline int jac_is_id(const jac* p) { return fe_is_zero(&p->z); }

static void jac_double(jac* r, const jac* p) {
    if (jac_is_id(p)) { *r = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    f_sqr(&FQ, &a, &p->x);
    f_sqr(&FQ, &b, &p->y);
    f_sqr(&FQ, &c, &b);
    f_add(&FQ, &t, &p->x, &b); f_sqr(&FQ, &t, &t); f_sub(&FQ, &t, &t, &a); f_sub(&FQ, &t, &t, &c);
    f_add(&FQ, &d, &t, &t);
    f_add(&FQ, &e, &a, &a); f_add(&FQ, &e, &e, &a);
    f_sqr(&FQ, &f, &e);
    f_mul(&FQ, &z3, &p->y, &p->z); f_add(&FQ, &z3, &z3, &z3);
    f_sub(&FQ, &x3, &f, &d); f_sub(&FQ, &x3, &x3, &d);
    f_add(&FQ, &c, &c, &c); f_add(&FQ, &c, &c, &c); f_add(&FQ, &c, &c, &c);
    f_sub(&FQ, &t, &d, &x3); f_mul(&FQ, &y3, &e, &t); f_sub(&FQ, &y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
    if (jac_is_id(p)) { *r = *q; return; }
    if (jac_is_id(q)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z); f_sqr(&FQ, &z2z2, &q->z);
    f_mul(&FQ, &u1, &p->x, &z2z2); f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s1, &p->y, &q->z); f_mul(&FQ, &s1, &s1, &z2z2);
    f_mul(&FQ, &s2, &q->y, &p->z); f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(&FQ, &h, &u2, &u1); f_sub(&FQ, &rr, &s2, &s1);
    f_sqr(&FQ, &hh, &h); f_mul(&FQ, &hhh, &h, &hh); f_mul(&FQ, &v, &u1, &hh);
    f_sqr(&FQ, &x3, &rr); f_sub(&FQ, &x3, &x3, &hhh); f_sub(&FQ, &x3, &x3, &v); f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3); f_mul(&FQ, &y3, &rr, &t); f_mul(&FQ, &t, &s1, &hhh); f_sub(&FQ, &y3, &y3, &t);
    f_mul(&FQ, &z3, &p->z, &q->z); f_mul(&FQ, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_mixed(jac* r, const jac* p, const aff* q) {
    if (aff_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { r->x = q->x; r->y = q->y; r->z = FQ_ONE; return; }
    /* madd-2007-bl (Z2 = 1): 7M + 4S */
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s2, &q->y, &p->z); f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&p->x, &u2)) {
        if (fe_eq(&p->y, &s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(&FQ, &h, &u2, &p->x);
    f_sqr(&FQ, &hh, &h);
    f_add(&FQ, &i, &hh, &hh); f_add(&FQ, &i, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &rr, &s2, &p->y); f_add(&FQ, &rr, &rr, &rr);
    f_mul(&FQ, &v, &p->x, &i);
    f_sqr(&FQ, &x3, &rr); f_sub(&FQ, &x3, &x3, &j); f_sub(&FQ, &x3, &x3, &v); f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3); f_mul(&FQ, &y3, &rr, &t);
    f_mul(&FQ, &t, &p->y, &j); f_add(&FQ, &t, &t, &t); f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &h); f_sqr(&FQ, &z3, &z3); f_sub(&FQ, &z3, &z3, &z1z1); f_sub(&FQ, &z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_affine(aff* r, const jac* p) {
    if (jac_is_id(p)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    f_inv(&FQ, &zi, &p->z);
    f_sqr(&FQ, &zi2, &zi);
    f_mul(&FQ, &zi3, &zi2, &zi);
    f_mul(&FQ, &r->x, &p->x, &zi2);
    f_mul(&FQ, &r->y, &p->y, &zi3);
}

/* ---- MSM (halo2 best_multiexp shape) ------------------------------------------------------ */
static size_t get_at(size_t segment, size_t c, const fe* canon) {
    size_t skip_bits = segment * c;
    if (skip_bits >= 256) return 0;
    size_t limb = skip_bits / 64, sh = skip_bits % 64;
    u128 v = canon->l[limb];
    if (limb + 1 < 4) v |= (u128)canon->l[limb + 1] << 64;
    return (size_t)((v >> sh) & (((u128)1 << c) - 1));
}

static void multiexp_serial(const fe* coeffs_canon, const aff* bases, size_t len, jac* acc) {
    size_t c;
    if (len < 4) c = 1; else if (len < 32) c = 3; else c = (size_t)ceil(log((double)len));
    size_t segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    jac* buckets = (jac*)malloc(sizeof(jac) * nb);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t i = 0; i < c; i++) jac_double(acc, acc);
        for (size_t i = 0; i < nb; i++) jac_set_id(&buckets[i]);
        for (size_t i = 0; i < len; i++) {
            size_t d = get_at(seg, c, &coeffs_canon[i]);
            if (d) jac_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        }
        jac running;
        jac_set_id(&running);
        for (size_t i = nb; i-- > 0;) {
            jac_add(&running, &running, &buckets[i]);
            jac_add(acc, acc, &running);
        }
    }
    free(buckets);
}

/* out[col] = sum_i scalars[col][i] * bases[i]; scalars Montgomery Fr, bases/out affine Montgomery Fq. */
EXPORT void orc_msm(const uint64_t* scalars_, const uint64_t* bases_, uint32_t n, uint32_t batch, uint64_t* out_,
                    int threads) {
    const fe* scalars = (const fe*)scalars_;
    const aff* bases = (const aff*)bases_;
    aff* out = (aff*)out_;
    int nt = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    nt = omp_get_max_threads();
#endif
    fe* canon = (fe*)malloc(sizeof(fe) * n);
    for (uint32_t col = 0; col < batch; col++) {
        const fe* sc = scalars + (size_t)col * n;
#pragma omp parallel for schedule(static)
        for (uint32_t i = 0; i < n; i++) f_from_mont(&FR, &canon[i], &sc[i]);
        jac total;
        jac_set_id(&total);
        if (n > (uint32_t)nt) {
            size_t chunk = n / nt;
            size_t nchunks = (n + chunk - 1) / chunk;
            jac* results = (jac*)malloc(sizeof(jac) * nchunks);
#pragma omp parallel for schedule(dynamic, 1)
            for (size_t ci = 0; ci < nchunks; ci++) {
                size_t lo = ci * chunk, len = lo + chunk <= n ? chunk : n - lo;
                jac_set_id(&results[ci]);
                multiexp_serial(canon + lo, bases + lo, len, &results[ci]);
            }
            for (size_t ci = 0; ci < nchunks; ci++) jac_add(&total, &total, &results[ci]);
            free(results);
        } else {
            multiexp_serial(canon, bases, n, &total);
        }
        jac_to_affine(&out[col], &total);
    }
    free(canon);
}

/* ---- toy SRS: g[i] = tau^i G, g_lagrange[i] = l_i(tau) G  (halo2 ParamsKZG::setup shape) --- */
static void fixed_base_mul(aff* out, const aff* table /*[32][255]*/, const fe* k_canon) {
    jac acc;
    jac_set_id(&acc);
    for (int w = 0; w < 32; w++) {
        uint32_t d = (uint32_t)((k_canon->l[w / 8] >> ((w % 8) * 8)) & 0xff);
        if (d) jac_add_mixed(&acc, &acc, &table[w * 255 + d - 1]);
    }
    jac_to_affine(out, &acc);
}

EXPORT void orc_srs(uint32_t k, const uint64_t* tau_canon_, uint64_t* g_, uint64_t* gl_, int threads) {
    const size_t n = (size_t)1 << k;
    aff *g = (aff*)g_, *gl = (aff*)gl_;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    /* table[w][d-1] = d * 2^(8w) * G */
    aff* table = (aff*)malloc(sizeof(aff) * 32 * 255);
    jac base;
    {
        fe one = {{1, 0, 0, 0}}, two = {{2, 0, 0, 0}};
        f_to_mont(&FQ, &base.x, &one);
        f_to_mont(&FQ, &base.y, &two);
        base.z = FQ_ONE;
    }
    for (int w = 0; w < 32; w++) {
        jac acc;
        jac_set_id(&acc);
        for (int d = 1; d <= 255; d++) {
            jac_add(&acc, &acc, &base);
            jac_to_affine(&table[w * 255 + d - 1], &acc);
        }
        for (int i = 0; i < 8; i++) jac_double(&base, &base);
    }
    fe tau, tau_m;
    memcpy(&tau, tau_canon_, sizeof tau);
    f_to_mont(&FR, &tau_m, &tau);
    /* powers of tau (Montgomery) */
    fe* pw = (fe*)malloc(sizeof(fe) * n);
    pw[0] = FR_ONE;
    for (size_t i = 1; i < n; i++) f_mul(&FR, &pw[i], &pw[i - 1], &tau_m);
    if (g) {
#pragma omp parallel for schedule(dynamic, 64)
        for (size_t i = 0; i < n; i++) {
            fe c;
            f_from_mont(&FR, &c, &pw[i]);
            fixed_base_mul(&g[i], table, &c);
        }
    }
    if (gl) {
        /* l_i(tau) = w^i (tau^n - 1) / (n (tau - w^i)) */
        fe w, tn, ninv, nn = {{(uint64_t)n, 0, 0, 0}}, nm;
        fr_omega(&w, k, 0);
        f_mul(&FR, &tn, &pw[n - 1], &tau_m);
        f_sub(&FR, &tn, &tn, &FR_ONE);
        f_to_mont(&FR, &nm, &nn);
        f_inv(&FR, &ninv, &nm);
        f_mul(&FR, &tn, &tn, &ninv);
        fe* wi = (fe*)malloc(sizeof(fe) * n);
        wi[0] = FR_ONE;
        for (size_t i = 1; i < n; i++) f_mul(&FR, &wi[i], &wi[i - 1], &w);
#pragma omp parallel for schedule(dynamic, 64)
        for (size_t i = 0; i < n; i++) {
            fe d, li, c;
            f_sub(&FR, &d, &tau_m, &wi[i]);
            f_inv(&FR, &d, &d);
            f_mul(&FR, &li, &wi[i], &tn);
            f_mul(&FR, &li, &li, &d);
            f_from_mont(&FR, &c, &li);
            fixed_base_mul(&gl[i], table, &c);
        }
        free(wi);
    }
    free(pw);
    free(table);
}

/* ---- stage (1): poly.rs on u64 coefficients with 128-bit exact products ------------------- */
/* c = a * b, schoolbook (poly.rs:86-90); len_a == len_b == len; out has 2*len-1 u128 as (lo,hi) u64 pairs. */
EXPORT void orc_poly_mul(const uint64_t* a, const uint64_t* b, uint32_t len, uint64_t* out_lohi) {
    u128* c = (u128*)calloc(2 * (size_t)len - 1, sizeof(u128));
    for (uint32_t i = 0; i < len; i++) {
        u128 ai = a[i];
        for (uint32_t j = 0; j < len; j++) c[i + j] += ai * b[j];
    }
    for (uint32_t i = 0; i < 2 * len - 1; i++) { out_lohi[2 * i] = (uint64_t)c[i]; out_lohi[2 * i + 1] = (uint64_t)(c[i] >> 64); }
    free(c);
}

/* reduce_by_modulus (poly.rs:180-191) on (lo,hi) pairs */
EXPORT void orc_poly_reduce(const uint64_t* in_lohi, uint32_t len, uint64_t q, uint64_t* out) {
    for (uint32_t i = 0; i < len; i++) {
        u128 v = ((u128)in_lohi[2 * i + 1] << 64) | in_lohi[2 * i];
        out[i] = (uint64_t)(v % q);
    }
}

/* divide_by_cyclo (poly.rs:113-177), literal long division by `cyclo` (len_c coefficients, leading first).
 * dividend: len_d coefficients in [0, q).  quotient_out: len_c entries, remainder_out: 2*(len_c-1)+1 entries.
 * Returns 0, or -1 where the reference would panic (zero leading divisor coefficient / usize underflow). */
EXPORT int orc_divide_by_cyclo(const uint64_t* dividend_, uint32_t len_d, const uint64_t* cyclo, uint32_t len_c,
                               uint64_t q, uint64_t* quotient_out, uint64_t* remainder_out) {
    const uint32_t deg_c = len_c - 1;
    int all_zero = 1;
    for (uint32_t i = 0; i < len_d; i++) if (dividend_[i]) { all_zero = 0; break; }
    if (len_d == 0 || all_zero) {
        memset(quotient_out, 0, sizeof(uint64_t) * (deg_c + 1));
        memset(remainder_out, 0, sizeof(uint64_t) * (2 * deg_c + 1));
        return 0;
    }
    __int128* dividend = (__int128*)malloc(sizeof(__int128) * len_d);
    for (uint32_t i = 0; i < len_d; i++) dividend[i] = dividend_[i];
    __int128* quot = (__int128*)malloc(sizeof(__int128) * (len_d + 1));
    uint32_t nq = 0, pos = 0;
    while (len_d - pos > len_c - 1) {
        if (cyclo[0] == 0) { free(dividend); free(quot); return -1; }
        __int128 ratio = dividend[pos] / (__int128)cyclo[0];
        quot[nq++] = ratio;
        for (uint32_t i = 0; i < len_c; i++) dividend[pos + i] -= ratio * (__int128)cyclo[i];
        pos++;
    }
    uint32_t qs = 0, rs = pos;
    while (qs < nq && quot[qs] == 0) qs++;
    while (rs < len_d && dividend[rs] == 0) rs++;
    uint32_t qlen = nq - qs, rlen = len_d - rs;
    if (qlen == 0 || rlen == 0 || qlen > deg_c + 1 || rlen > 2 * deg_c + 1) { free(dividend); free(quot); return -1; }
    memset(quotient_out, 0, sizeof(uint64_t) * (deg_c + 1));
    memset(remainder_out, 0, sizeof(uint64_t) * (2 * deg_c + 1));
    for (uint32_t i = 0; i < qlen; i++) quotient_out[deg_c + 1 - qlen + i] = (uint64_t)quot[qs + i];
    for (uint32_t i = 0; i < rlen; i++) {
        __int128 v = dividend[rs + i] % (__int128)q;
        if (v < 0) v += q;
        remainder_out[2 * deg_c + 1 - rlen + i] = (uint64_t)v;
    }
    free(dividend);
    free(quot);
    return 0;
}

EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ================================================================================================================
 * Stage (1b): the per-cell witness values of the BFV circuit, on the CPU.
 *
 * Restates, in witness-generation mode (values only: no selectors / copy constraints -- what the reference's `prove`
 * runs once `configs/bfv.json` holds the break points), the calls of
 *   /root/reference/examples/bfv.rs:70-165 (phase 0) and :172-301 (the phase-1 callback) over
 *   /root/reference/src/poly_chip.rs (every method) and
 *   halo2-base v0.3.0-ce GateChip / RangeChip, axiom-eth RlcChip [UPSTREAM-RECALL; the Python
 *   restatement oracle/halo2_base.py is the line-by-line version this follows and is tested against].
 * Single-threaded, like the reference's own code (no rayon in the repo).  Cell values that are small
 * integers are carried as unsigned __int128 and converted to Montgomery form when assigned; the few genuinely
 * field-valued cells (negated powers of two, chi-key factors, RLC accumulators, is_zero inverses) use Fr arithmetic.
 * is_zero inverses are batch-inverted at the end, as halo2's `Assigned::Rational` cells are.
 * ================================================================================================================ */
typedef struct {
    fe* a;      size_t n, cap;        /* advice cells (Montgomery) */
    fe* lk;     size_t nl, lcap;      /* cells_to_lookup values, shared by the contexts, creation order */
    size_t* inv_at; size_t n_inv, inv_cap;   /* advice cells that hold x and must become 1/x */
    int overflow;
} wctx;

static inline void fe_from_u128(fe* r, u128 v) {
    fe c = {{(uint64_t)v, (uint64_t)(v >> 64), 0, 0}};
    f_to_mont(&FR, r, &c);
}
static inline void w_push(wctx* c, const fe* v) {
    if (c->n >= c->cap) { c->overflow = 1; return; }
    c->a[c->n++] = *v;
}
static inline void w_push_u(wctx* c, u128 v) { fe t; fe_from_u128(&t, v); w_push(c, &t); }
static inline void w_look(wctx* c, const fe* v) {
    if (c->nl >= c->lcap) { c->overflow = 1; return; }
    c->lk[c->nl++] = *v;
}
static inline void w_look_u(wctx* c, u128 v) { fe t; fe_from_u128(&t, v); w_look(c, &t); }
static const fe FE_ZERO = {{0, 0, 0, 0}};

/* GateChip (values) */
static void g_add_u(wctx* c, u128 a, u128 b) { w_push_u(c, a); w_push_u(c, b); w_push(c, &FR_ONE); w_push_u(c, a + b); }
static void g_mul_u(wctx* c, u128 a, u128 b) { w_push(c, &FE_ZERO); w_push_u(c, a); w_push_u(c, b); w_push_u(c, a * b); }
/* sub on field values: cells [out, b, 1, a] */
static void g_sub_f(wctx* c, const fe* a, const fe* b, fe* out) {
    f_sub(&FR, out, a, b);
    w_push(c, out); w_push(c, b); w_push(c, &FR_ONE); w_push(c, a);
}
static void g_mul_f(wctx* c, const fe* a, const fe* b, fe* out) {
    f_mul(&FR, out, a, b);
    w_push(c, &FE_ZERO); w_push(c, a); w_push(c, b); w_push(c, out);
}
/* is_zero: [is_zero, a, inv, 1, 0, a, is_zero, 0]; returns is_zero as 0 / 1 */
static int g_is_zero_f(wctx* c, const fe* a) {
    const int z = fe_is_zero(a);
    const fe* zf = z ? &FR_ONE : &FE_ZERO;
    w_push(c, zf); w_push(c, a);
    if (z) w_push(c, &FR_ONE);                     /* Assigned::Trivial(F::one()) */
    else {
        if (c->n_inv < c->inv_cap) c->inv_at[c->n_inv++] = c->n; else c->overflow = 1;
        w_push(c, a);                              /* placeholder: batch-inverted by w_finish */
    }
    w_push(c, &FR_ONE); w_push(c, &FE_ZERO); w_push(c, a); w_push(c, zf); w_push(c, &FE_ZERO);
    return z;
}
static void w_finish(wctx* c) {                    /* Montgomery's trick over the recorded cells */
    const size_t m = c->n_inv;
    if (!m) return;
    fe* pre = (fe*)malloc(sizeof(fe) * m);
    fe run = FR_ONE;
    for (size_t i = 0; i < m; i++) { pre[i] = run; f_mul(&FR, &run, &run, &c->a[c->inv_at[i]]); }
    fe inv;
    f_inv(&FR, &inv, &run);
    for (size_t i = m; i-- > 0;) {
        fe x = c->a[c->inv_at[i]], t;
        f_mul(&FR, &t, &inv, &pre[i]);
        c->a[c->inv_at[i]] = t;
        f_mul(&FR, &inv, &inv, &x);
    }
    free(pre);
    c->n_inv = 0;
}

/* RangeChip */
static uint32_t bitlen128(u128 v) { uint32_t b = 0; while (v) { b++; v >>= 1; } return b; }
static void r_range_check(wctx* c, u128 a, uint32_t range_bits, uint32_t lb) {
    const uint32_t k = (range_bits + lb - 1) / lb, rem = range_bits % lb;
    const u128 mask = ((u128)1 << lb) - 1;
    u128 last;
    if (k == 1) { w_look_u(c, a); last = a; }
    else {
        u128 acc = a & mask;
        w_push_u(c, acc); w_look_u(c, acc);
        last = acc;
        for (uint32_t i = 1; i < k; i++) {
            const u128 limb = (a >> (lb * i)) & mask;
            acc += limb << (lb * i);
            w_push_u(c, limb); w_push_u(c, (u128)1 << (lb * i)); w_push_u(c, acc);
            w_look_u(c, limb);
            last = limb;
        }
    }
    if (rem == 1) { w_push(c, &FE_ZERO); w_push_u(c, last); w_push_u(c, last); w_push_u(c, last); }
    else if (rem > 1) {
        const u128 m = (u128)1 << (lb - rem);
        g_mul_u(c, last, m);
        w_look_u(c, last * m);
    }
}
static void neg_pow2(fe* out, uint32_t bits) { fe p; fe_from_u128(&p, (u128)1 << bits); f_neg(&FR, out, &p); }
/* cells [a + 2^bits - b, b, 1, a + 2^bits, -2^bits, 1, a], then range_check(first, bits) */
static void r_check_less_than(wctx* c, u128 a, u128 b, uint32_t bits, uint32_t lb) {
    const u128 shift_a = ((u128)1 << bits) + a;
    fe np2;
    neg_pow2(&np2, bits);
    w_push_u(c, shift_a - b); w_push_u(c, b); w_push(c, &FR_ONE); w_push_u(c, shift_a); w_push(c, &np2); w_push(c, &FR_ONE); w_push_u(c, a);
    r_range_check(c, shift_a - b, bits, lb);
}
static void r_check_less_than_safe(wctx* c, u128 a, u128 b, uint32_t lb) {
    const uint32_t rb = (bitlen128(b) + lb - 1) / lb * lb;
    r_range_check(c, a, rb, lb);
    r_check_less_than(c, a, b, rb, lb);
}
static int r_is_less_than(wctx* c, u128 a, u128 b, uint32_t num_bits, uint32_t lb) {
    const uint32_t k = (num_bits + lb - 1) / lb, padded = k * lb;
    const u128 shift_a = ((u128)1 << padded) + a, shifted = shift_a - b;
    fe np2;
    neg_pow2(&np2, padded);
    w_push_u(c, shifted); w_push_u(c, b); w_push(c, &FR_ONE); w_push_u(c, shift_a); w_push(c, &np2); w_push(c, &FR_ONE); w_push_u(c, a);
    r_range_check(c, shifted, padded + lb, lb);
    fe top;
    fe_from_u128(&top, (shifted >> padded) & (((u128)1 << lb) - 1));    /* = cells_to_lookup.last() */
    return g_is_zero_f(c, &top);
}
static u128 r_div_mod(wctx* c, u128 a, uint64_t b, uint32_t a_num_bits, uint32_t lb) {
    const u128 div = a / b, rem = a % b;
    w_push_u(c, rem); w_push_u(c, b); w_push_u(c, div); w_push_u(c, a);
    r_check_less_than_safe(c, div, (((u128)1 << a_num_bits) / b) + 1, lb);
    r_check_less_than_safe(c, rem, b, lb);
    return rem;
}
/* RlcChip::compute_rlc_fixed_len: cells [in0, in1, acc1, in2, acc2, ...]; returns the final accumulator */
static void rlc_fixed_len(wctx* c, const u128* in, uint32_t len, const fe* gamma, fe* out) {
    fe run, x;
    fe_from_u128(&run, in[0]);
    w_push(c, &run);
    for (uint32_t i = 1; i < len; i++) {
        fe_from_u128(&x, in[i]);
        f_mul(&FR, &run, &run, gamma);
        f_add(&FR, &run, &run, &x);
        w_push(c, &x); w_push(c, &run);
    }
    *out = run;
}

/* PolyChip methods (poly_chip.rs), values only */
static void pc_constrain_mul(wctx* gate, wctx* rlc, const u128* a, uint32_t la, const u128* b, uint32_t lb_, const u128* cc, uint32_t lc, const fe* gamma) {
    fe ea, eb, ec;                                                            /* poly_chip.rs:97-104 */
    rlc_fixed_len(rlc, a, la, gamma, &ea);
    rlc_fixed_len(rlc, b, lb_, gamma, &eb);
    rlc_fixed_len(rlc, cc, lc, gamma, &ec);
    w_push(gate, &FE_ZERO); w_push(gate, &ea); w_push(gate, &eb); w_push(gate, &ec);   /* :107-115 */
}
static void pc_in_range(wctx* c, const u128* coeff, uint32_t len, uint64_t z, uint64_t y, uint32_t lb) {   /* :270-317 */
    const uint32_t ybits = bitlen128(y);
    for (uint32_t i = 0; i < len; i++) {
        r_check_less_than_safe(c, coeff[i], y, lb);
        const int in1 = r_is_less_than(c, coeff[i], (u128)z + 1, ybits, lb);
        const int not_in2 = r_is_less_than(c, coeff[i], (u128)y - z, ybits, lb);
        const int in2 = 1 - not_in2;                                         /* not = sub(1, a): [out, a, 1, 1] */
        w_push_u(c, in2); w_push_u(c, not_in2); w_push(c, &FR_ONE); w_push(c, &FR_ONE);
        const int nb = 1 - in2, out = in1 | in2;                             /* or: [1-b, 1, b, 1, b, a, 1-b, out] */
        w_push_u(c, nb); w_push(c, &FR_ONE); w_push_u(c, in2); w_push(c, &FR_ONE); w_push_u(c, in2); w_push_u(c, in1); w_push_u(c, nb); w_push_u(c, out);
    }
}
static void pc_chi_key(wctx* c, const u128* coeff, uint32_t len, uint64_t z) {                          /* :320-354 */
    fe zf, one = FR_ONE, zero = FE_ZERO;
    fe_from_u128(&zf, z);
    for (uint32_t i = 0; i < len; i++) {
        fe v, f1, f2, f3, f12, f123;
        fe_from_u128(&v, coeff[i]);
        g_sub_f(c, &v, &zero, &f1);
        g_sub_f(c, &v, &one, &f2);
        g_sub_f(c, &v, &zf, &f3);
        g_mul_f(c, &f1, &f2, &f12);
        g_mul_f(c, &f12, &f3, &f123);
    }
}
static void pc_in_modulus_field(wctx* c, const u128* coeff, uint32_t len, uint64_t q, uint32_t lb) {    /* :357-366 */
    for (uint32_t i = 0; i < len; i++) r_check_less_than_safe(c, coeff[i], q, lb);
}
static void pc_reduce_by_modulo(wctx* c, const u128* coeff, uint32_t len, uint64_t q, uint32_t nbits, uint32_t lb, u128* out) {   /* :226-252 */
    for (uint32_t i = 0; i < len; i++) out[i] = r_div_mod(c, coeff[i], q, nbits, lb);
}
static void pc_add(wctx* c, const u128* a, const u128* b, uint32_t len, u128* out) {                    /* :122-144 */
    for (uint32_t i = 0; i < len; i++) { g_add_u(c, a[i], b[i]); out[i] = a[i] + b[i]; }
}
static void pc_constrain_equality(wctx* c, const u128* a, const u128* b, uint32_t len) {                /* :255-264 */
    for (uint32_t i = 0; i < len; i++) {
        fe x, y, d;
        fe_from_u128(&x, a[i]); fe_from_u128(&y, b[i]);
        g_sub_f(c, &x, &y, &d);
        g_is_zero_f(c, &d);
    }
}

static uint32_t log2_ceil_u64(uint64_t x) { uint32_t b = bitlen128(x); return b - ((x & (x - 1)) == 0 ? 1 : 0); }

/* in[9]: pk0, pk1, m, u, e0, e1, c0, c1 (N values each), cyclo (N + 1), big-endian, already parsed to u64.
 * adv0 / adv1 / adv2: flat advice of the phase-0 gate, phase-1 gate and phase-1 RLC contexts; lk: lookup cells.
 * counts_out = {cells0, cells1, cells2, lookups}.  Returns 0, -1 where the reference asserts / panics, -2 when a
 * capacity is too small. */
EXPORT int orc_bfv_witness(const uint64_t* const* in, uint32_t N, uint64_t Q, uint64_t T, uint64_t B, uint32_t lookup_bits,
                           const uint64_t* gamma_mont, uint64_t* adv0, uint64_t cap0, uint64_t* adv1, uint64_t cap1,
                           uint64_t* adv2, uint64_t cap2, uint64_t* lk, uint64_t capl, uint64_t* counts_out) {
    const uint32_t lb = lookup_bits, L2 = 2 * N - 1, LC = N + 1, LR = 2 * N + 1;
    const uint32_t qbits = bitlen128(Q);
    for (int p = 0; p < 9; p++)
        for (uint32_t i = 0; i < (p == 8 ? LC : N); i++) if (in[p][i] > Q) return -1;                  /* poly.rs:28 */
    size_t* inv_at = (size_t*)malloc(sizeof(size_t) * (size_t)(8 * N + 64));
    wctx c0 = {(fe*)adv0, 0, cap0, (fe*)lk, 0, capl, inv_at, 0, 0, 0};
    /* ---- phase 0 (bfv.rs:70-165) ---- */
    u128 *P[9];
    for (int p = 0; p < 9; p++) {
        const uint32_t len = p == 8 ? LC : N;
        P[p] = (u128*)malloc(sizeof(u128) * len);
        for (uint32_t i = 0; i < len; i++) P[p][i] = in[p][i];
    }
    enum { PK0, PK1, M, U, E0, E1, C0, C1, CY };
    for (int p = 0; p < 9; p++) for (uint32_t i = 0; i < (p == 8 ? LC : N); i++) w_push_u(&c0, P[p][i]);   /* :101-109 */
    w_push_u(&c0, Q / T);                                                                                 /* :115 */
    u128 *pku[2], *quo[2], *qtc[2], *rem[2];
    uint64_t* tmp_lohi = (uint64_t*)malloc(sizeof(uint64_t) * 2 * (size_t)LR);
    uint64_t* red = (uint64_t*)malloc(sizeof(uint64_t) * L2);
    uint64_t* q64 = (uint64_t*)malloc(sizeof(uint64_t) * LC);
    uint64_t* r64 = (uint64_t*)malloc(sizeof(uint64_t) * LR);
    int rc = 0;
    for (int h = 0; h < 2; h++) {
        pku[h] = (u128*)malloc(sizeof(u128) * L2);
        quo[h] = (u128*)malloc(sizeof(u128) * LC);
        qtc[h] = (u128*)malloc(sizeof(u128) * LR);
        rem[h] = (u128*)malloc(sizeof(u128) * LR);
        orc_poly_mul(in[h == 0 ? PK0 : PK1], in[U], N, tmp_lohi);                                         /* :131-132 */
        for (uint32_t i = 0; i < L2; i++) pku[h][i] = ((u128)tmp_lohi[2 * i + 1] << 64) | tmp_lohi[2 * i];
        orc_poly_reduce(tmp_lohi, L2, Q, red);                                                            /* :139-140 */
        if (orc_divide_by_cyclo(red, L2, in[CY], LC, Q, q64, r64) != 0) rc = -1;                          /* :143-146 */
        orc_poly_mul(q64, in[CY], LC, tmp_lohi);                                                          /* :149-150 */
        for (uint32_t i = 0; i < LC; i++) quo[h][i] = q64[i];
        for (uint32_t i = 0; i < LR; i++) { qtc[h][i] = ((u128)tmp_lohi[2 * i + 1] << 64) | tmp_lohi[2 * i]; rem[h][i] = r64[i]; }
    }
    for (int h = 0; h < 2; h++) for (uint32_t i = 0; i < L2; i++) w_push_u(&c0, pku[h][i]);               /* :135-136 */
    for (int h = 0; h < 2; h++) for (uint32_t i = 0; i < LC; i++) w_push_u(&c0, quo[h][i]);               /* :156-157 */
    for (int h = 0; h < 2; h++) for (uint32_t i = 0; i < LR; i++) w_push_u(&c0, qtc[h][i]);               /* :160-161 */
    for (int h = 0; h < 2; h++) for (uint32_t i = 0; i < LR; i++) w_push_u(&c0, rem[h][i]);               /* :164-165 */
    /* ---- phase 1 (bfv.rs:172-301) ---- */
    wctx g = {(fe*)adv1, 0, cap1, (fe*)lk, c0.nl, capl, inv_at, 0, (size_t)(8 * N + 64), 0};
    wctx r = {(fe*)adv2, 0, cap2, (fe*)lk, 0, 0, NULL, 0, 0, 0};
    const fe* gamma = (const fe*)gamma_mont;
    pc_in_range(&g, P[E0], N, B, Q, lb);                                                                  /* :189 */
    pc_in_range(&g, P[E1], N, B, Q, lb);                                                                  /* :190 */
    pc_chi_key(&g, P[U], N, Q - 1);                                                                       /* :201 */
    pc_in_range(&g, P[M], N, T / 2, Q, lb);                                                               /* :210 */
    u128* s = (u128*)malloc(sizeof(u128) * LR);
    u128* s_mod = (u128*)malloc(sizeof(u128) * LR);
    u128* red_pk = (u128*)malloc(sizeof(u128) * L2);
    u128* cres = (u128*)malloc(sizeof(u128) * N);
    u128* t1 = (u128*)malloc(sizeof(u128) * N);
    const uint32_t pku_bits = qbits + qbits + log2_ceil_u64(N);                  /* poly.rs:101 on pk * u */
    const uint32_t qtc_bits = qbits + qbits + log2_ceil_u64((uint64_t)N + 1);    /* poly.rs:101 on quotient * cyclo */
    for (int h = 0; h < 2; h++) {
        pc_constrain_mul(&g, &r, P[h == 0 ? PK0 : PK1], N, P[U], N, pku[h], L2, gamma);                   /* :215 / :264 */
        pc_reduce_by_modulo(&g, pku[h], L2, Q, pku_bits, lb, red_pk);                                     /* :219 / :268 */
        pc_in_modulus_field(&g, quo[h], LC, Q, lb);                                                       /* :225 / :274 */
        pc_in_modulus_field(&g, rem[h], LR, Q, lb);                                                       /* :226 / :275 */
        /* reduce_by_cyclo (poly_chip.rs:183-223) */
        pc_constrain_mul(&g, &r, quo[h], LC, P[CY], LC, qtc[h], LR, gamma);                               /* :205 */
        pc_add(&g, qtc[h], rem[h], LR, s);                                                                /* :208 */
        const uint32_t s_bits = (qtc_bits > qbits ? qtc_bits : qbits) + 1;
        pc_reduce_by_modulo(&g, s, LR, Q, s_bits, lb, s_mod);                                             /* :211 */
        pc_constrain_equality(&g, s_mod + (LR - L2), red_pk, L2);                                         /* :214-217 */
        const u128* pk_u_red = rem[h] + (LR - N);                                                         /* :222 */
        if (h == 0) {
            const uint64_t delta = Q / T;
            for (uint32_t i = 0; i < N; i++) { g_mul_u(&g, P[M][i], delta); t1[i] = P[M][i] * delta; }    /* :243 */
            pc_add(&g, pk_u_red, t1, N, cres);                                                            /* :247 */
            pc_add(&g, cres, P[E0], N, t1);                                                               /* :251 */
            const uint32_t mb = qbits + bitlen128(delta), cb = (qbits > mb ? qbits : mb) + 1;
            pc_reduce_by_modulo(&g, t1, N, Q, (cb > qbits ? cb : qbits) + 1, lb, cres);                   /* :255 */
            pc_constrain_equality(&g, cres, P[C0], N);                                                    /* :259 */
        } else {
            pc_add(&g, pk_u_red, P[E1], N, t1);                                                           /* :292 */
            pc_reduce_by_modulo(&g, t1, N, Q, qbits + 1, lb, cres);                                       /* :296 */
            pc_constrain_equality(&g, cres, P[C1], N);                                                    /* :300 */
        }
    }
    w_finish(&g);
    counts_out[0] = c0.n; counts_out[1] = g.n; counts_out[2] = r.n; counts_out[3] = g.nl;
    if (c0.overflow || g.overflow || r.overflow) rc = -2;
    for (int p = 0; p < 9; p++) free(P[p]);
    for (int h = 0; h < 2; h++) { free(pku[h]); free(quo[h]); free(qtc[h]); free(rem[h]); }
    free(tmp_lohi); free(red); free(q64); free(r64); free(s); free(s_mod); free(red_pk); free(cres); free(t1); free(inv_at);
    return rc;
}

EXPORT void orc_from_mont_array(int which, const uint64_t* in, uint64_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) f_from_mont(which ? &FQ : &FR, (fe*)(out + 4 * i), (const fe*)(in + 4 * i));
}
