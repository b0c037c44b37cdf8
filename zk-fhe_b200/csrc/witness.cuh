// Device-side restatement of the halo2-base v0.3.0 gate semantics used by zk-fhe's
// PolyChip (src/poly_chip.rs): each primitive *emits* the same advice cells, in the
// same order, as the CPU builder does (oracle/halo2_base.py is the
// checker).  One thread owns one polynomial coefficient and writes that coefficient's
// cells contiguously, so a whole chip call is one data-parallel launch.
//
// All cell values are Fr in Montgomery form (the advice-table layout halo2 commits to).
//
// The same emitters also record the circuit *structure* when the witness object was
// created for keygen / mock (META = true): per cell a selector bit, a "constant cell" bit
// and the cell it is copy-constrained to.  Layout therefore has a single source of truth.
#pragma once
#include "ff.cuh"

namespace zkfhe {

// Global cell id: context in the top 4 bits, flat offset below; 0 means "no cell".
static constexpr uint64_t CELL_NONE = 0;
__host__ __device__ __forceinline__ uint64_t cell_id(uint32_t ctx_id, uint64_t off) {
    return ((uint64_t)(ctx_id + 1) << 60) | off;
}
__host__ __device__ __forceinline__ uint32_t cell_ctx(uint64_t id) { return (uint32_t)(id >> 60) - 1; }
__host__ __device__ __forceinline__ uint64_t cell_off(uint64_t id) { return id & ((1ull << 60) - 1); }

static constexpr uint8_t META_SELECTOR = 1;   // a gate starts at this cell
static constexpr uint8_t META_CONSTANT = 2;   // Constant(c) cell: constrained to the fixed constant equal to its value
static constexpr uint8_t META_ASSERT_ZERO = 4;   // gate.assert_is_const(cell, 0)
static constexpr uint8_t META_ASSERT_ONE = 8;    // gate.assert_is_const(cell, 1)
static constexpr uint8_t META_COPY_CONFLICT = 16;   // internal: a second, different copy source was requested for this cell

// A value together with the cell that holds it (CELL_NONE for a fresh witness / constant).
struct Val {
    fr_t v;
    uint64_t cell;
};

struct Emit {
    fr_t* a;            // advice cursor base for this coefficient
    fr_t* l;            // lookup-cell cursor base for this coefficient
    uint32_t na, nl;
    // structure recording (null when the witness object is not in keygen / mock mode)
    uint8_t* flags;     // [cells of this coefficient]
    uint64_t* copy;     // [cells of this coefficient] cell id this cell must equal, or CELL_NONE
    uint64_t* lk_src;   // [lookups of this coefficient] cell id pushed to cells_to_lookup
    uint64_t base_id;   // cell id of this coefficient's first advice cell

    __device__ __forceinline__ uint64_t here() const { return base_id + na; }
    __device__ __forceinline__ void meta(uint8_t f, uint64_t c) {
        if (flags) { flags[na] = f; copy[na] = c; }
    }
    // Witness(v)
    __device__ __forceinline__ Val wit(const fr_t& v) {
        Val r{v, here()};
        meta(0, CELL_NONE);
        fe_store(a + na, v); na++;
        return r;
    }
    // Constant(c)
    __device__ __forceinline__ Val con(const fr_t& v) {
        Val r{v, here()};
        meta(META_CONSTANT, CELL_NONE);
        fe_store(a + na, v); na++;
        return r;
    }
    // Existing(x): new cell, copy-constrained to x
    __device__ __forceinline__ Val ex(const Val& x) {
        Val r{x.v, here()};
        meta(0, x.cell);
        fe_store(a + na, x.v); na++;
        return r;
    }
    // enable the gate selector `back` cells behind the cursor
    __device__ __forceinline__ void gate_at(uint32_t back) {
        if (flags) flags[na - back] |= META_SELECTOR;
    }
    // extra equality between an already emitted cell of this coefficient (`back` behind the cursor) and `to`
    __device__ __forceinline__ void equal_at(uint32_t back, uint64_t to) {
        if (!flags) return;
        if (copy[na - back] == CELL_NONE) copy[na - back] = to;
        else if (copy[na - back] != to) flags[na - back] |= META_COPY_CONFLICT;   // one copy slot per cell: keygen / mock refuse
    }
    // gate.assert_is_const on the cell `back` behind the cursor (value 0 or 1)
    __device__ __forceinline__ void assert_const_at(uint32_t back, bool one) {
        if (flags) flags[na - back] |= one ? META_ASSERT_ONE : META_ASSERT_ZERO;
    }
    __device__ __forceinline__ void look(const Val& x) {
        if (lk_src) lk_src[nl] = x.cell;
        fe_store(l + nl, x.v); nl++;
    }
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
// Operands are `Val`s: an operand with a cell is assigned as Existing(cell), one without as a
// Constant (that is how the chip passes constants: Constant(F::from(z)) etc.).
__device__ __forceinline__ Val q_cell(Emit& e, const Val& x) { return x.cell ? e.ex(x) : e.con(x.v); }
__device__ __forceinline__ Val konst(const fr_t& v) { return Val{v, CELL_NONE}; }

__device__ __forceinline__ Val g_add(Emit& e, const Val& a, const Val& b) {   // [a, b, 1, out]
    q_cell(e, a); q_cell(e, b); e.con(fe_one<FR>());
    Val out = e.wit(add(a.v, b.v));
    e.gate_at(4);
    return out;
}
__device__ __forceinline__ Val g_sub(Emit& e, const Val& a, const Val& b) {   // [out, b, 1, a]
    Val out = e.wit(sub(a.v, b.v));
    q_cell(e, b); e.con(fe_one<FR>()); q_cell(e, a);
    e.gate_at(4);
    return out;
}
__device__ __forceinline__ Val g_mul(Emit& e, const Val& a, const Val& b) {   // [0, a, b, out]
    e.con(fe_zero<FR>()); q_cell(e, a); q_cell(e, b);
    Val out = e.wit(mul(a.v, b.v));
    e.gate_at(4);
    return out;
}
__device__ __forceinline__ Val g_not(Emit& e, const Val& a) { return g_sub(e, konst(fe_one<FR>()), a); }
__device__ __forceinline__ Val g_or(Emit& e, const Val& a, const Val& b) {
    // [1-b, 1, b, 1, b, a, 1-b, out], gates at 0 and 4, equalities (0,6) and (2,4)
    const fr_t one = fe_one<FR>();
    fr_t not_b = sub(one, b.v);
    Val c0 = e.wit(not_b);
    e.con(one); q_cell(e, b); e.con(one); q_cell(e, b); q_cell(e, a);
    e.wit(not_b);
    e.equal_at(1, c0.cell);                               // (0,6); (2,4) is implied: both copy b
    Val out = e.wit(sub(add(a.v, b.v), mul(a.v, b.v)));
    e.gate_at(8); e.gate_at(4);
    return out;
}
__device__ inline Val g_is_zero(Emit& e, const Val& a) {
    // [is_zero, a, inv, 1, 0, a, is_zero, 0], gates at 0 and 4, equality (0,6); returns cell 6
    const fr_t one = fe_one<FR>(), zero = fe_zero<FR>();
    bool z = is_zero(a.v);
    fr_t iz = z ? one : zero;
    fr_t iv = (z || eq(a.v, one)) ? one : inv(a.v);    // Assigned::Trivial(1) for zero, else a^-1
    Val c0 = e.wit(iz);
    q_cell(e, a); e.wit(iv); e.con(one); e.con(zero); q_cell(e, a);
    Val out = e.wit(iz);
    e.equal_at(1, c0.cell);
    e.con(zero);
    e.gate_at(8); e.gate_at(4);
    return out;
}
__device__ __forceinline__ Val g_is_equal(Emit& e, const Val& a, const Val& b) {
    Val d = g_sub(e, a, b);
    return g_is_zero(e, d);
}

// ---- RangeChip (halo2-base range.rs) -------------------------------------------------------
// returns the last cell pushed to cells_to_lookup
__device__ inline Val r_range_check(Emit& e, const Val& a, uint32_t range_bits, uint32_t lb) {
    const uint32_t k = (range_bits + lb - 1) / lb, rem = range_bits % lb;
    Val last;
    if (k == 1) {
        e.look(a);
        last = a;
    } else {
        // inner_product(limbs, [1, 2^lb, ...]) = [l0, l1, 2^lb, acc1, l2, 2^2lb, acc2, ...], gates at 0,3,6,..
        fr_t c = from_mont(a.v);
        last = e.wit(mont_u64(canon_bits(c, 0, lb)));
        e.look(last);
        for (uint32_t i = 1; i < k; i++) {
            last = e.wit(mont_u64(canon_bits(c, lb * i, lb)));
            e.con(mont_pow2(lb * i));
            e.wit(to_mont(canon_low(c, lb * (i + 1))));
            e.gate_at(4);
            e.look(last);
        }
        e.equal_at(1, a.cell);                            // ctx.constrain_equal(a, acc)
    }
    if (rem == 1) {                                       // assert_bit: [0, x, x, x]
        e.con(fe_zero<FR>()); e.ex(last); e.ex(last); e.ex(last);
        e.gate_at(4);
    } else if (rem > 1) {
        Val chk = g_mul(e, last, konst(mont_pow2(lb - rem)));
        e.look(chk);
        last = chk;
    }
    return last;
}
__device__ inline void r_check_less_than(Emit& e, const Val& a, const Val& b, uint32_t num_bits, uint32_t lb) {
    // [a + 2^bits - b, b, 1, a + 2^bits, -2^bits, 1, a], gates at 0 and 3
    const fr_t one = fe_one<FR>();
    fr_t pow2 = mont_pow2(num_bits);
    fr_t shift_a = add(pow2, a.v);
    Val first = e.wit(sub(shift_a, b.v));
    q_cell(e, b); e.con(one); e.wit(shift_a); e.con(neg(pow2)); e.con(one); q_cell(e, a);
    e.gate_at(7); e.gate_at(4);
    r_range_check(e, first, num_bits, lb);
}
// b: constant in Montgomery form with its bit length (u64 or BigUint bound)
__device__ inline void r_check_less_than_safe(Emit& e, const Val& a, const fr_t& b, uint32_t b_bits, uint32_t lb) {
    const uint32_t range_bits = (b_bits + lb - 1) / lb * lb;
    r_range_check(e, a, range_bits, lb);
    r_check_less_than(e, a, konst(b), range_bits, lb);
}
__device__ inline Val r_is_less_than(Emit& e, const Val& a, const Val& b, uint32_t num_bits, uint32_t lb) {
    const fr_t one = fe_one<FR>();
    const uint32_t k = (num_bits + lb - 1) / lb, padded = k * lb;
    fr_t pow_padded = mont_pow2(padded);
    fr_t shift_a = add(pow_padded, a.v);
    Val shifted = e.wit(sub(shift_a, b.v));
    q_cell(e, b); e.con(one); e.wit(shift_a); e.con(neg(pow_padded)); e.con(one); q_cell(e, a);
    e.gate_at(7); e.gate_at(4);
    Val top = r_range_check(e, shifted, padded + lb, lb);
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
