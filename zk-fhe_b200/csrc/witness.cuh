// Device-side restatement of the halo2-base v0.3.0 gate semantics used by zk-fhe's
// PolyChip (src/poly_chip.rs): each primitive *emits* the same advice cells, in the
// same order, as the CPU builder does (oracle/halo2_base.py is the
// checker).  One thread owns one polynomial coefficient and writes that coefficient's
// cells contiguously, so a whole chip call is one data-parallel launch.
//
// All cell values are Fr in Montgomery form (the advice-table layout halo2 commits to).
#pragma once
#include "ff.cuh"

namespace zkfhe {

struct Emit {
    fr_t* a;      // advice cursor base for this coefficient
    fr_t* l;      // lookup-cell cursor base for this coefficient
    uint32_t na, nl;
    __device__ __forceinline__ void cell(const fr_t& v) { fe_store(a + na, v); na++; }
    __device__ __forceinline__ void look(const fr_t& v) { fe_store(l + nl, v); nl++; }
};

__device__ __forceinline__ fr_t mont_u64(uint64_t v) {
    fr_t c = fe_zero<FR>();
    c.v[0] = (uint32_t)v;
    c.v[1] = (uint32_t)(v >> 32);
    return to_mont(c);
}
__device__ __forceinline__ fr_t mont_pow2(uint32_t bits) {   // 2^bits, bits < 254
    fr_t c = fe_zero<FR>();
    c.v[bits >> 5] = 1u << (bits & 31);
    return to_mont(c);
}
// bits [off, off+nb) of a canonical integer, nb <= 32
__device__ __forceinline__ uint32_t canon_bits(const fr_t& c, uint32_t off, uint32_t nb) {
    if (off >= 256) return 0;
    uint32_t limb = off >> 5, sh = off & 31;
    uint64_t two = c.v[limb];
    if (limb + 1 < 8) two |= (uint64_t)c.v[limb + 1] << 32;
    return (uint32_t)((two >> sh) & ((nb >= 32) ? 0xffffffffull : ((1ull << nb) - 1)));
}
// c mod 2^bits
__device__ __forceinline__ fr_t canon_low(const fr_t& c, uint32_t bits) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int lo = i * 32;
        if ((int)bits >= lo + 32) r.v[i] = c.v[i];
        else if ((int)bits <= lo) r.v[i] = 0;
        else r.v[i] = c.v[i] & ((1u << (bits - lo)) - 1);
    }
    return r;
}
__device__ __forceinline__ uint32_t canon_bitlen(const fr_t& c) {
    for (int i = 7; i >= 0; i--)
        if (c.v[i]) return 32 * i + (32 - __clz(c.v[i]));
    return 0;
}
// canonical 256-bit integer divided by a 64-bit modulus: quotient (256-bit) and remainder
__device__ inline void canon_divmod_u64(const fr_t& c, uint64_t q, fr_t& quot, uint64_t& rem_out) {
    unsigned __int128 rem = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        unsigned __int128 cur = (rem << 32) | c.v[i];
        quot.v[i] = (uint32_t)(cur / q);     // rem < q  =>  cur / q < 2^32
        rem = cur % q;
    }
    rem_out = (uint64_t)rem;
}

// ---- GateChip (halo2-base flex_gate.rs) ------------------------------------
__device__ __forceinline__ fr_t g_add(Emit& e, const fr_t& a, const fr_t& b) {
    fr_t out = add(a, b);
    e.cell(a); e.cell(b); e.cell(fe_one<FR>()); e.cell(out);
    return out;
}
__device__ __forceinline__ fr_t g_sub(Emit& e, const fr_t& a, const fr_t& b) {
    fr_t out = sub(a, b);
    e.cell(out); e.cell(b); e.cell(fe_one<FR>()); e.cell(a);
    return out;
}
__device__ __forceinline__ fr_t g_mul(Emit& e, const fr_t& a, const fr_t& b) {
    fr_t out = mul(a, b);
    e.cell(fe_zero<FR>()); e.cell(a); e.cell(b); e.cell(out);
    return out;
}
__device__ __forceinline__ fr_t g_not(Emit& e, const fr_t& a) { return g_sub(e, fe_one<FR>(), a); }
__device__ __forceinline__ fr_t g_or(Emit& e, const fr_t& a, const fr_t& b) {
    const fr_t one = fe_one<FR>();
    fr_t not_b = sub(one, b);
    fr_t out = sub(add(a, b), mul(a, b));
    e.cell(not_b); e.cell(one); e.cell(b); e.cell(one); e.cell(b); e.cell(a); e.cell(not_b); e.cell(out);
    return out;
}
__device__ inline fr_t g_is_zero(Emit& e, const fr_t& a) {
    const fr_t one = fe_one<FR>(), zero = fe_zero<FR>();
    bool z = is_zero(a);
    fr_t iz = z ? one : zero;
    fr_t iv = (z || eq(a, one)) ? one : inv(a);    // Assigned::Trivial(1) for zero, else a^-1
    e.cell(iz); e.cell(a); e.cell(iv); e.cell(one); e.cell(zero); e.cell(a); e.cell(iz); e.cell(zero);
    return iz;
}
__device__ __forceinline__ fr_t g_is_equal(Emit& e, const fr_t& a, const fr_t& b) {
    fr_t d = g_sub(e, a, b);
    return g_is_zero(e, d);
}

// ---- RangeChip (halo2-base range.rs) -------------------------------------------------------
// returns the last cell pushed to cells_to_lookup
__device__ inline fr_t r_range_check(Emit& e, const fr_t& a, uint32_t range_bits, uint32_t lb) {
    const uint32_t k = (range_bits + lb - 1) / lb, rem = range_bits % lb;
    fr_t last;
    if (k == 1) {
        e.look(a);
        last = a;
    } else {
        fr_t c = from_mont(a);
        last = mont_u64(canon_bits(c, 0, lb));
        e.cell(last);
        e.look(last);
        for (uint32_t i = 1; i < k; i++) {
            last = mont_u64(canon_bits(c, lb * i, lb));
            e.cell(last);
            e.cell(mont_pow2(lb * i));
            e.cell(to_mont(canon_low(c, lb * (i + 1))));
            e.look(last);
        }
    }
    if (rem == 1) {
        e.cell(fe_zero<FR>()); e.cell(last); e.cell(last); e.cell(last);
    } else if (rem > 1) {
        fr_t m = mont_pow2(lb - rem);
        fr_t chk = mul(last, m);
        e.cell(fe_zero<FR>()); e.cell(last); e.cell(m); e.cell(chk);
        e.look(chk);
        last = chk;
    }
    return last;
}
__device__ inline void r_check_less_than(Emit& e, const fr_t& a, const fr_t& b, uint32_t num_bits, uint32_t lb) {
    const fr_t one = fe_one<FR>();
    fr_t pow2 = mont_pow2(num_bits);
    fr_t shift_a = add(pow2, a);
    fr_t first = sub(shift_a, b);
    e.cell(first); e.cell(b); e.cell(one); e.cell(shift_a); e.cell(neg(pow2)); e.cell(one); e.cell(a);
    r_range_check(e, first, num_bits, lb);
}
// b given in Montgomery form with its bit length (u64 or BigUint bound)
__device__ inline void r_check_less_than_safe(Emit& e, const fr_t& a, const fr_t& b, uint32_t b_bits, uint32_t lb) {
    const uint32_t range_bits = (b_bits + lb - 1) / lb * lb;
    r_range_check(e, a, range_bits, lb);
    r_check_less_than(e, a, b, range_bits, lb);
}
__device__ inline fr_t r_is_less_than(Emit& e, const fr_t& a, const fr_t& b, uint32_t num_bits, uint32_t lb) {
    const fr_t one = fe_one<FR>();
    const uint32_t k = (num_bits + lb - 1) / lb, padded = k * lb;
    fr_t pow_padded = mont_pow2(padded);
    fr_t shift_a = add(pow_padded, a);
    fr_t shifted = sub(shift_a, b);
    e.cell(shifted); e.cell(b); e.cell(one); e.cell(shift_a); e.cell(neg(pow_padded)); e.cell(one); e.cell(a);
    fr_t top = r_range_check(e, shifted, padded + lb, lb);
    return g_is_zero(e, top);
}

// ---- cell-count model (host + device): must agree with the emitters above -------------------
struct CellCount { uint32_t cells, lookups; };
__host__ __device__ inline CellCount cc_range_check(uint32_t bits, uint32_t lb) {
    uint32_t k = (bits + lb - 1) / lb, rem = bits % lb;
    CellCount c{k == 1 ? 0u : 3 * k - 2, k};
    if (rem == 1) c.cells += 4;
    else if (rem > 1) { c.cells += 4; c.lookups += 1; }
    return c;
}
__host__ __device__ inline CellCount cc_check_less_than(uint32_t bits, uint32_t lb) {
    CellCount c = cc_range_check(bits, lb);
    c.cells += 7;
    return c;
}
__host__ __device__ inline CellCount cc_check_less_than_safe(uint32_t b_bits, uint32_t lb) {
    uint32_t rb = (b_bits + lb - 1) / lb * lb;
    CellCount a = cc_range_check(rb, lb), b = cc_check_less_than(rb, lb);
    return CellCount{a.cells + b.cells, a.lookups + b.lookups};
}
__host__ __device__ inline CellCount cc_is_less_than(uint32_t bits, uint32_t lb) {
    uint32_t padded = (bits + lb - 1) / lb * lb;
    CellCount c = cc_range_check(padded + lb, lb);
    c.cells += 7 + 8;
    return c;
}

}  // namespace zkfhe
