// Stage (1) of the prove path on the GPU: witness generation for the BFV circuit.
//
//  (1a) off-circuit polynomial arithmetic -- mirror of /root/reference/src/poly.rs:
//       Poly::mul (O(N^2) BigInt schoolbook there; here an exact integer convolution done as
//       a 2N-point NTT over Fr, valid because N*Q^2 < r -- SURVEY.md App. D),
//       reduce_by_modulus, divide_by_cyclo (closed form for x^N + 1).
//  (1b) in-circuit cell values -- mirror of /root/reference/src/poly_chip.rs on top of the
//       halo2-base gate semantics in witness.cuh.  Cell order inside a context is call order
//       (examples/bfv.rs:172-301), and inside a call it is coefficient-major, so every chip
//       method is a parallel-for over coefficients with analytically known offsets
//       (SURVEY.md App. E).
//
// Everything stays in HBM between calls; data-dependent reference asserts set bits in a
// device status word that zkfhe_status() / zkfhe_witness_status read back.
#include <new>
#include <vector>
#include "common.cuh"
#include "witness.cuh"
#include "host_ff.h"

using namespace zkfhe;

// ---------------------------------------------------------------------------------------------
// objects behind the opaque handles
// ---------------------------------------------------------------------------------------------
#include "witness_types.cuh"

namespace zkfhe {

enum StatusBit : uint32_t {
    ST_COEFF_GT_MODULUS = 1u << 0,   // poly.rs:28
    ST_COEFF_BITS = 1u << 1,         // poly.rs:51
    ST_CYCLO_SHAPE = 1u << 2,        // divisor is not x^N + 1 (documented assumption, poly.rs:111)
    ST_QUOTIENT_EMPTY = 1u << 3,     // poly.rs:158 usize underflow
    ST_REMAINDER_EMPTY = 1u << 4,    // poly.rs:164 usize underflow
    ST_CELL_COUNT = 1u << 5,         // internal: emitter and cell-count model disagree
    ST_FLAG_D_NONZERO = 1u << 16,    // scratch flags used by divide_by_cyclo
    ST_FLAG_Q_NONZERO = 1u << 17,
    ST_FLAG_R_NONZERO = 1u << 18,
};

static int status_word(zkfhe_ctx* ctx, uint32_t** out) {
    void* p;
    bool fresh = ctx->ws.find("status") == ctx->ws.end();
    ZK_TRY(ws_get(ctx, "status", 256, &p));
    if (fresh) ZK_CUDA(ctx, cudaMemsetAsync(p, 0, 256, ctx->stream));
    *out = (uint32_t*)p;
    return ZKFHE_OK;
}

static int vec_reserve(zkfhe_ctx* ctx, DevVec& v, size_t need) {
    if (need <= v.cap) return ZKFHE_OK;
    size_t ncap = v.cap * 2 > need ? v.cap * 2 : need;
    if (ncap < (1u << 16)) ncap = 1u << 16;
    fr_t* np;
    ZK_CUDA(ctx, cudaMalloc(&np, ncap * sizeof(fr_t)));
    if (v.size) ZK_CUDA(ctx, cudaMemcpyAsync(np, v.p, v.size * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (v.p) {
        ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
        ZK_CUDA(ctx, cudaFree(v.p));
    }
    v.p = np;
    v.cap = ncap;
    return ZKFHE_OK;
}

// metadata arrays follow the capacity of the cell vector they describe
template <class T> static int arr_reserve(zkfhe_ctx* ctx, DevArr<T>& a, size_t cap, size_t used) {
    if (cap <= a.cap) return ZKFHE_OK;
    T* np;
    ZK_CUDA(ctx, cudaMalloc(&np, cap * sizeof(T)));
    ZK_CUDA(ctx, cudaMemsetAsync(np, 0, cap * sizeof(T), ctx->stream));
    if (used && a.p) ZK_CUDA(ctx, cudaMemcpyAsync(np, a.p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    if (a.p) {
        ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
        ZK_CUDA(ctx, cudaFree(a.p));
    }
    a.p = np;
    a.cap = cap;
    return ZKFHE_OK;
}

static inline uint32_t bitlen64(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }
static inline uint32_t log2_ceil_u64(uint64_t x) {   // halo2_base::utils::log2_ceil
    return bitlen64(x) - ((x & (x - 1)) == 0 ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------
// (1a) poly kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_poly_from_u64(const uint64_t* in, fr_t* out, uint32_t len, uint64_t modulus, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    uint64_t v = in[i];
    if (v > modulus) atomicOr(status, ST_COEFF_GT_MODULUS);
    fr_t c = fe_zero<FR>();
    c.v[0] = (uint32_t)v;
    c.v[1] = (uint32_t)(v >> 32);
    fe_store(out + i, c);
}

__global__ void k_poly_check_bits(const fr_t* in, uint32_t len, uint32_t max_bits, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (canon_bitlen(fe_load(in + i)) > max_bits) atomicOr(status, ST_COEFF_BITS);
}

// canonical big-endian (len) -> Montgomery little-endian, zero padded to n2
__global__ void k_poly_ntt_load(const fr_t* in, uint32_t len, fr_t* out, uint32_t n2) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    fe_store(out + j, j < len ? to_mont(fe_load(in + (len - 1 - j))) : fe_zero<FR>());
}
__global__ void k_pointwise_mul(fr_t* a, const fr_t* b, uint32_t n) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fe_store(a + j, mul(fe_load(a + j), fe_load(b + j)));
}
// Montgomery little-endian -> canonical big-endian (out_len), with the from_big_int bits assert
__global__ void k_poly_ntt_store(const fr_t* in, fr_t* out, uint32_t out_len, uint32_t max_bits, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    fr_t c = from_mont(fe_load(in + (out_len - 1 - i)));
    if (canon_bitlen(c) > max_bits) atomicOr(status, ST_COEFF_BITS);
    fe_store(out + i, c);
}
__global__ void k_poly_reduce(const fr_t* in, fr_t* out, uint32_t len, uint64_t q) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    fr_t quot;
    uint64_t rem;
    canon_divmod_u64(fe_load(in + i), q, quot, rem);
    fr_t c = fe_zero<FR>();
    c.v[0] = (uint32_t)rem;
    c.v[1] = (uint32_t)(rem >> 32);
    fe_store(out + i, c);
}

__device__ __forceinline__ uint64_t canon_lo64(const fr_t& c) { return (uint64_t)c.v[0] | ((uint64_t)c.v[1] << 32); }
__device__ __forceinline__ bool canon_fits64(const fr_t& c) { return (c.v[2] | c.v[3] | c.v[4] | c.v[5] | c.v[6] | c.v[7]) == 0; }
__device__ __forceinline__ fr_t canon_from_u64(uint64_t v) {
    fr_t c = fe_zero<FR>();
    c.v[0] = (uint32_t)v;
    c.v[1] = (uint32_t)(v >> 32);
    return c;
}

// Long division by x^N + 1 in closed form.  D has len_d in (N, 2N] coefficients
// already in [0, q).  quotient: N+1 slots, remainder: 2N+1 slots, both left-padded with zeros.
__global__ void k_divide_by_cyclo(const fr_t* D, uint32_t len_d, const fr_t* cyclo, uint32_t N, uint64_t q,
                                  fr_t* quotient, fr_t* remainder, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nq = len_d - N;                 // raw quotient length
    uint32_t flags = 0;
    if (i <= N) {                                  // divisor shape check: 1, 0, ..., 0, 1
        fr_t c = fe_load(cyclo + i);
        bool want_one = (i == 0 || i == N);
        bool ok = canon_fits64(c) && canon_lo64(c) == (want_one ? 1u : 0u);
        if (!ok) flags |= ST_CYCLO_SHAPE;
        // quotient slot i (N+1 slots): raw quotient right-aligned
        uint32_t pad = N + 1 - nq;
        fr_t qv = fe_zero<FR>();
        if (i >= pad) {
            qv = fe_load(D + (i - pad));
            if (!is_zero(qv)) flags |= ST_FLAG_Q_NONZERO | ST_FLAG_D_NONZERO;
        }
        fe_store(quotient + i, qv);
    }
    if (i < 2 * N + 1) {
        // remainder: N raw values (dividend positions nq .. len_d-1) right-aligned in 2N+1 slots
        uint32_t pad = N + 1;
        fr_t rv = fe_zero<FR>();
        if (i >= pad) {
            uint32_t j = nq + (i - pad);           // dividend position
            uint64_t dj = canon_lo64(fe_load(D + j));
            if (dj) flags |= ST_FLAG_D_NONZERO;
            uint64_t r = dj;
            if (j >= N) {                          // feedback of quotient term q[j-N] = D[j-N]
                uint64_t dq = canon_lo64(fe_load(D + (j - N)));
                r = dj >= dq ? dj - dq : dj + (q - dq);      // mod_floor (poly.rs:169-172)
                if (dj != dq) flags |= ST_FLAG_R_NONZERO;
            } else if (dj) {
                flags |= ST_FLAG_R_NONZERO;
            }
            rv = canon_from_u64(r);
        }
        fe_store(remainder + i, rv);
    }
    if (flags) atomicOr(status + 1, flags);
}
// Applies the reference's control flow to the flags gathered above, then clears them.
__global__ void k_divide_by_cyclo_finish(uint32_t* status) {
    uint32_t f = status[1];
    status[1] = 0;
    if (!(f & ST_FLAG_D_NONZERO)) return;          // all-zero dividend: shortcut (poly.rs:118-123), cyclo not inspected
    uint32_t err = f & ST_CYCLO_SHAPE;
    if (!(f & ST_FLAG_Q_NONZERO)) err |= ST_QUOTIENT_EMPTY;
    if (!(f & ST_FLAG_R_NONZERO)) err |= ST_REMAINDER_EMPTY;
    if (err) atomicOr(status, err);
}

// ---------------------------------------------------------------------------------------------
// (1b) chip kernels: one thread per coefficient
// ---------------------------------------------------------------------------------------------
struct View {             // an assigned polynomial: coefficient i at p[i * stride], cell id id0 + i * stride
    const fr_t* p;
    uint32_t stride;
    uint64_t id0;
};
struct OutSpan {          // where coefficient i's cells go
    fr_t* adv;            // + i * cpc
    fr_t* lk;             // + i * lpc
    uint32_t cpc, lpc;
    uint8_t* flags;       // structure recording (null unless the witness records structure)
    uint64_t* copy;
    uint64_t* lk_src;
    uint64_t base_id;     // cell id of adv[0]
};
__device__ __forceinline__ Val view_at(const View& a, uint32_t i) {
    return Val{fe_load(a.p + (size_t)i * a.stride), a.id0 + (uint64_t)i * a.stride};
}
__device__ __forceinline__ Emit make_emit(const OutSpan& o, uint32_t i) {
    Emit e;
    e.a = o.adv + (size_t)i * o.cpc;
    e.l = o.lk + (size_t)i * o.lpc;
    e.na = e.nl = 0;
    e.flags = o.flags ? o.flags + (size_t)i * o.cpc : nullptr;
    e.copy = o.flags ? o.copy + (size_t)i * o.cpc : nullptr;
    e.lk_src = o.flags ? o.lk_src + (size_t)i * o.lpc : nullptr;
    e.base_id = o.base_id + (uint64_t)i * o.cpc;
    return e;
}
#define CHIP_PROLOGUE                                                   \
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;                 \
    if (i >= len) return;                                               \
    Emit e = make_emit(o, i);
#define CHIP_EPILOGUE \
    if (e.na != o.cpc || e.nl != o.lpc) atomicOr(status, ST_CELL_COUNT);

__global__ void k_assign_from_poly(const fr_t* canon, uint32_t len, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    e.wit(to_mont(fe_load(canon + i)));
    CHIP_EPILOGUE
}
__global__ void k_assign_constant(uint64_t v, OutSpan o) {
    Emit e = make_emit(o, 0);
    e.con(mont_u64(v));
}

__global__ void k_chip_in_range(View a, uint32_t len, uint64_t z, uint64_t y, uint32_t lb, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    const uint32_t y_bits = 64 - __clzll(y);
    Val c = view_at(a, i);
    r_check_less_than_safe(e, c, mont_u64(y), y_bits, lb);
    Val in1 = r_is_less_than(e, c, konst(mont_u64(z + 1)), y_bits, lb);
    Val nin2 = r_is_less_than(e, c, konst(mont_u64(y - z)), y_bits, lb);
    Val in2 = g_not(e, nin2);
    g_or(e, in1, in2);
    e.assert_const_at(1, true);
    CHIP_EPILOGUE
}
__global__ void k_chip_chi_key(View a, uint32_t len, uint64_t z, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    Val c = view_at(a, i);
    Val f1 = g_sub(e, c, konst(fe_zero<FR>()));
    Val f2 = g_sub(e, c, konst(fe_one<FR>()));
    Val f3 = g_sub(e, c, konst(mont_u64(z)));
    Val f12 = g_mul(e, f1, f2);
    g_mul(e, f12, f3);
    e.assert_const_at(1, false);
    CHIP_EPILOGUE
}
__global__ void k_chip_check_lt_safe(View a, uint32_t len, uint64_t b, uint32_t lb, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    r_check_less_than_safe(e, view_at(a, i), mont_u64(b), 64 - __clzll(b), lb);
    CHIP_EPILOGUE
}
__global__ void k_chip_div_mod(View a, uint32_t len, uint64_t q, fr_t bound_mont, uint32_t bound_bits, uint32_t lb,
                               OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    Val am = view_at(a, i);
    fr_t quot;
    uint64_t rem;
    canon_divmod_u64(from_mont(am.v), q, quot, rem);
    fr_t q_m = mont_u64(q);
    Val rem_c = e.wit(mont_u64(rem));            // [rem, Q, div, a], gate at 0
    e.con(q_m);
    Val div_c = e.wit(to_mont(quot));
    e.ex(am);
    e.gate_at(4);
    r_check_less_than_safe(e, div_c, bound_mont, bound_bits, lb);
    r_check_less_than_safe(e, rem_c, q_m, 64 - __clzll(q), lb);
    CHIP_EPILOGUE
}
__global__ void k_chip_add(View a, View b, uint32_t len, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    g_add(e, view_at(a, i), view_at(b, i));
    CHIP_EPILOGUE
}
__global__ void k_chip_scalar_mul(View a, View scalar, uint32_t len, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    g_mul(e, view_at(a, i), view_at(scalar, 0));
    CHIP_EPILOGUE
}
__global__ void k_chip_is_equal(View a, View b, uint32_t len, OutSpan o, uint32_t* status) {
    CHIP_PROLOGUE
    g_is_equal(e, view_at(a, i), view_at(b, i));
    e.assert_const_at(2, true);                  // the bool is cell 6 of is_zero's 8
    CHIP_EPILOGUE
}
// constrain_mul's final region [0, a(gamma), b(gamma), c(gamma)], gate at 0
__global__ void k_chip_gate4(View a, View b, View c, OutSpan o) {
    Emit e = make_emit(o, 0);
    e.con(fe_zero<FR>());
    e.ex(view_at(a, 0));
    e.ex(view_at(b, 0));
    e.ex(view_at(c, 0));
    e.gate_at(4);
}
// gate.assert_is_const(cell, 0) on existing cells (safe_trim_leading_zeroes)
__global__ void k_meta_assert_zero(uint8_t* flags, uint32_t stride, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) flags[(size_t)i * stride] |= META_ASSERT_ZERO;
}

// RlcChip::compute_rlc_fixed_len: Horner in gamma as a parallel scan of affine maps
// x -> x * M + V.  One CTA per chain; cells [in0, in1, acc1, in2, acc2, ...], RLC gate
// (a*gamma + b - c on 3 rows) at offsets 0, 2, 4, ...
static constexpr uint32_t RLC_THREADS = 256;
__global__ void __launch_bounds__(RLC_THREADS) k_chip_rlc(View in, uint32_t len, fr_t gamma, OutSpan o) {
    __shared__ fr_t M[2][RLC_THREADS], V[2][RLC_THREADS];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (len + RLC_THREADS - 1) / RLC_THREADS;
    const uint32_t lo = min(t * per, len), hi = min(lo + per, len);
    fr_t* out = o.adv;
    fr_t acc = fe_zero<FR>();
    for (uint32_t j = lo; j < hi; j++) acc = add(mul(acc, gamma), fe_load(in.p + (size_t)j * in.stride));
    fe_store(&M[0][t], pow_u64(gamma, hi - lo));
    fe_store(&V[0][t], acc);
    __syncthreads();
    int cur = 0;
    for (uint32_t d = 1; d < RLC_THREADS; d <<= 1) {
        fr_t m = fe_load(&M[cur][t]), v = fe_load(&V[cur][t]);
        if (t >= d) {
            // compose earlier map (m0, v0) then this one: x -> (x*m0 + v0)*m + v
            fr_t m0 = fe_load(&M[cur][t - d]), v0 = fe_load(&V[cur][t - d]);
            v = add(mul(v0, m), v);
            m = mul(m0, m);
        }
        fe_store(&M[cur ^ 1][t], m);
        fe_store(&V[cur ^ 1][t], v);
        cur ^= 1;
        __syncthreads();
    }
    acc = t ? fe_load(&V[cur][t - 1]) : fe_zero<FR>();
    for (uint32_t j = lo; j < hi; j++) {
        fr_t x = fe_load(in.p + (size_t)j * in.stride);
        acc = add(mul(acc, gamma), x);
        const uint64_t src = in.id0 + (uint64_t)j * in.stride;
        if (j == 0) {
            fe_store(out, x);
            if (o.flags) { o.flags[0] = len > 1 ? META_SELECTOR : 0; o.copy[0] = src; }
        } else {
            fe_store(out + 2 * j - 1, x);
            fe_store(out + 2 * j, acc);
            if (o.flags) {
                o.flags[2 * j - 1] = 0;
                o.copy[2 * j - 1] = src;
                o.flags[2 * j] = j + 1 < len ? META_SELECTOR : 0;
                o.copy[2 * j] = CELL_NONE;
            }
        }
    }
}

__global__ void k_gather_cells(const fr_t* const* bases, const uint64_t* ctx_off, fr_t* out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t co = ctx_off[i];
    fe_store(out + i, fe_load(bases[co >> 60] + (co & ((1ull << 60) - 1))));
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
static inline uint32_t blocks_for(uint32_t n, uint32_t t = 128) { return (n + t - 1) / t; }

static int poly_alloc(zkfhe_ctx* ctx, uint32_t len, uint64_t max_bits, zkfhe_poly** out) {
    zkfhe_poly* p = new (std::nothrow) zkfhe_poly();
    if (!p) return fail(ctx, ZKFHE_ERR_CUDA, "out of host memory");
    p->ctx = ctx;    // stream-ordered allocation: no device-wide synchronisation on alloc / free
    cudaError_t e = cudaMallocAsync((void**)&p->d, (size_t)(len ? len : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) {
        delete p;
        return fail(ctx, ZKFHE_ERR_CUDA, "cudaMalloc(poly): %s", cudaGetErrorString(e));
    }
    p->len = len;
    p->max_bits = max_bits;
    *out = p;
    return ZKFHE_OK;
}

static const char* status_message(uint32_t st) {
    if (st & ST_COEFF_GT_MODULUS) return "assertion failed: coeff <= modulus (src/poly.rs:28)";
    if (st & ST_COEFF_BITS) return "assertion failed: coeff.bits() <= max_bits (src/poly.rs:51)";
    if (st & ST_CYCLO_SHAPE) return "divide_by_cyclo: divisor is not x^N + 1 (assumption at src/poly.rs:111)";
    if (st & ST_QUOTIENT_EMPTY) return "attempt to subtract with overflow: quotient.len() - 1 (src/poly.rs:158)";
    if (st & ST_REMAINDER_EMPTY) return "attempt to subtract with overflow: remainder.len() - 1 (src/poly.rs:164)";
    if (st & ST_CELL_COUNT) return "internal: emitted cell count differs from the layout model";
    return "ok";
}

// 2^nbits / q + 1 as a canonical 256-bit integer (host), nbits < 256
static void bound_pow2_div(uint32_t nbits, uint64_t q, uint32_t out[8], uint32_t* bits) {
    uint32_t num[8] = {0};
    num[nbits >> 5] = 1u << (nbits & 31);
    unsigned __int128 rem = 0;
    for (int i = 7; i >= 0; i--) {
        unsigned __int128 cur = (rem << 32) | num[i];
        out[i] = (uint32_t)(cur / q);
        rem = cur % q;
    }
    for (int i = 0; i < 8; i++) {           // + 1
        if (++out[i] != 0) break;
    }
    *bits = 0;
    for (int i = 7; i >= 0; i--)
        if (out[i]) { *bits = 32 * i + (32 - __builtin_clz(out[i])); break; }
}

__global__ void k_to_mont_one(fr_t* x) { fe_store(x, to_mont(fe_load(x))); }

static int chip_out(zkfhe_witness* w, uint32_t ctx_id, uint32_t len, CellCount per, OutSpan* o, uint64_t* adv_base) {
    zkfhe_ctx* ctx = w->ctx;
    DevVec& A = w->adv[ctx_id];
    DevVec& L = w->lk[ctx_id];
    ZK_TRY(vec_reserve(ctx, A, A.size + (size_t)len * per.cells));
    ZK_TRY(vec_reserve(ctx, L, L.size + (size_t)len * per.lookups));
    o->adv = A.p + A.size;
    o->lk = L.p + L.size;
    o->cpc = per.cells;
    o->lpc = per.lookups;
    o->flags = nullptr;
    o->copy = nullptr;
    o->lk_src = nullptr;
    o->base_id = cell_id(ctx_id, A.size);
    if (w->record) {
        ZK_TRY(arr_reserve(ctx, w->flags[ctx_id], A.cap, A.size));
        ZK_TRY(arr_reserve(ctx, w->copy[ctx_id], A.cap, A.size));
        ZK_TRY(arr_reserve(ctx, w->lk_src[ctx_id], L.cap ? L.cap : 1, L.size));
        o->flags = w->flags[ctx_id].p + A.size;
        o->copy = w->copy[ctx_id].p + A.size;
        o->lk_src = w->lk_src[ctx_id].p + L.size;
    }
    *adv_base = A.size;
    A.size += (size_t)len * per.cells;
    L.size += (size_t)len * per.lookups;
    return ZKFHE_OK;
}
static inline View view_of(zkfhe_witness* w, const zkfhe_assigned_poly* p) {
    return View{w->adv[p->ctx_id].p + p->base, p->stride, cell_id(p->ctx_id, p->base)};
}
static inline View view_of_cell(zkfhe_witness* w, const zkfhe_cell& c) {
    return View{w->adv[c.ctx_id].p + c.offset, 1, cell_id(c.ctx_id, c.offset)};
}
static int check_poly(zkfhe_witness* w, const zkfhe_assigned_poly* p, const char* what) {
    if (!p) return fail(w->ctx, ZKFHE_ERR_ARG, "%s: null polynomial", what);
    if (p->ctx_id > 2 || p->len == 0 || p->stride == 0 ||
        p->base + (uint64_t)(p->len - 1) * p->stride >= w->adv[p->ctx_id].size)
        return fail(w->ctx, ZKFHE_ERR_ARG, "%s: assigned polynomial is out of range of context %u", what, p->ctx_id);
    return ZKFHE_OK;
}
#define W_ENTER(w, gate)                                                                     \
    if (!(w)) return ZKFHE_ERR_ARG;                                                          \
    zkfhe_ctx* ctx = (w)->ctx;                                                               \
    if ((gate) > 2) return fail(ctx, ZKFHE_ERR_ARG, "context id %u out of range", (gate));   \
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));                                                \
    uint32_t* status;                                                                        \
    ZK_TRY(status_word(ctx, &status));

static const uint64_t P_BITS = 254;   // bits of the Fr modulus (poly_chip.rs:90-91)

}  // namespace zkfhe

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int zkfhe_status(zkfhe_ctx* ctx) {
    if (!ctx) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* status;
    ZK_TRY(status_word(ctx, &status));
    uint32_t h = 0;
    ZK_TRY(read_back(ctx, &h, status, 4));
    h &= 0xffffu;
    if (h == 0) return ZKFHE_OK;
    ZK_CUDA(ctx, cudaMemsetAsync(status, 0, 4, ctx->stream));
    return fail(ctx, ZKFHE_ERR_ASSERT, "%s", status_message(h));
}

// ---- Poly -------------------------------------------------------------------------------------
int zkfhe_poly_from_u64(zkfhe_ctx* ctx, const uint64_t* h_coeffs, uint32_t len, uint64_t modulus, zkfhe_poly** out) {
    if (!ctx || !out || (!h_coeffs && len)) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_u64: null pointer");
    if (len == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to subtract with overflow: coefficients.len() - 1 (src/poly.rs:33)");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* status;
    ZK_TRY(status_word(ctx, &status));
    void* stage;
    ZK_TRY(ws_get(ctx, "poly_stage", (size_t)len * 8, &stage));
    // through the page-locked arena: no stream drain per upload (nine uploads open every proof), and the device-side
    // staging buffer is only reused by the NEXT call's copy, which the stream orders after this call's kernel
    ZK_TRY(upload_async(ctx, stage, h_coeffs, (size_t)len * 8));
    ZK_TRY(poly_alloc(ctx, len, bitlen64(modulus), out));
    k_poly_from_u64<<<blocks_for(len), 128, 0, ctx->stream>>>((const uint64_t*)stage, (*out)->d, len, modulus, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

// Poly::from_string itself (poly.rs:21-40): `len` decimal integers separated by single commas.  Parsing 9 x 1024
// strings per proof in the host language of the caller costs more than the transfer; here it is one pass in C.
int zkfhe_poly_from_decimal(zkfhe_ctx* ctx, const char* text, size_t text_len, uint32_t len, uint64_t modulus, zkfhe_poly** out) {
    if (!ctx || !out || !text) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_decimal: null pointer");
    std::vector<uint64_t> vals;
    vals.reserve(len);
    size_t i = 0;
    while (i < text_len) {
        uint64_t v = 0;
        size_t digits = 0;
        while (i < text_len && text[i] >= '0' && text[i] <= '9') {
            const uint64_t d = (uint64_t)(text[i] - '0');
            if (v > (UINT64_MAX - d) / 10) return fail(ctx, ZKFHE_ERR_ARG, "coefficient does not fit u64 (the reference modulus is u64)");
            v = v * 10 + d;
            digits++;
            i++;
        }
        if (!digits) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_decimal: invalid digit found in string (src/poly.rs:25) at offset %zu", i);
        vals.push_back(v);
        if (i < text_len) {
            if (text[i] != ',') return fail(ctx, ZKFHE_ERR_ARG, "poly_from_decimal: invalid digit found in string (src/poly.rs:25) at offset %zu", i);
            i++;
            if (i == text_len) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_decimal: trailing separator");
        }
    }
    if (vals.size() != len) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_decimal: %zu coefficients, expected %u", vals.size(), len);
    return zkfhe_poly_from_u64(ctx, vals.data(), len, modulus, out);
}

int zkfhe_poly_from_u256(zkfhe_ctx* ctx, const uint64_t* h, uint32_t len, uint64_t max_bits, zkfhe_poly** out) {
    if (!ctx || !out || (!h && len)) return fail(ctx, ZKFHE_ERR_ARG, "poly_from_u256: null pointer");
    if (len == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to subtract with overflow: coefficients.len() - 1 (src/poly.rs:48)");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* status;
    ZK_TRY(status_word(ctx, &status));
    ZK_TRY(poly_alloc(ctx, len, max_bits, out));
    ZK_CUDA(ctx, cudaMemcpyAsync((*out)->d, h, (size_t)len * 32, cudaMemcpyHostToDevice, ctx->stream));
    k_poly_check_bits<<<blocks_for(len), 128, 0, ctx->stream>>>((*out)->d, len, (uint32_t)(max_bits > 256 ? 256 : max_bits), status);
    ZK_CHECK_LAUNCH(ctx);
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_poly_mul(zkfhe_ctx* ctx, const zkfhe_poly* a, const zkfhe_poly* b, zkfhe_poly** out) {
    if (!ctx || !a || !b || !out) return fail(ctx, ZKFHE_ERR_ARG, "poly_mul: null pointer");
    if (a->len != b->len) return fail(ctx, ZKFHE_ERR_ASSERT, "assertion failed: deg_a == deg_b (src/poly.rs:78)");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* status;
    ZK_TRY(status_word(ctx, &status));
    const uint32_t len = a->len, out_len = 2 * len - 1;
    const uint64_t max_bits = a->max_bits + b->max_bits + log2_ceil_u64(len);    // poly.rs:101
    if (max_bits >= P_BITS)
        return fail(ctx, ZKFHE_ERR_OVERFLOW, "poly_mul: product coefficients may reach %llu bits >= 254 "
                    "(not exact over Fr; the circuit asserts the same bound at src/poly_chip.rs:94)",
                    (unsigned long long)max_bits);
    uint32_t log_n2 = 1;
    while ((1u << log_n2) < out_len) log_n2++;
    const uint32_t n2 = 1u << log_n2;
    fr_t* buf;
    ZK_TRY(ws_get(ctx, "poly_ntt", (size_t)2 * n2 * sizeof(fr_t), (void**)&buf));
    k_poly_ntt_load<<<blocks_for(n2), 128, 0, ctx->stream>>>(a->d, len, buf, n2);
    ZK_CHECK_LAUNCH(ctx);
    k_poly_ntt_load<<<blocks_for(n2), 128, 0, ctx->stream>>>(b->d, len, buf + n2, n2);
    ZK_CHECK_LAUNCH(ctx);
    ZK_TRY(ntt_run(ctx, buf, n2, n2, buf, n2, log_n2, 2, 0, 0));
    k_pointwise_mul<<<blocks_for(n2), 128, 0, ctx->stream>>>(buf, buf + n2, n2);
    ZK_CHECK_LAUNCH(ctx);
    ZK_TRY(ntt_run(ctx, buf, n2, n2, buf, n2, log_n2, 1, 1, 0));
    ZK_TRY(poly_alloc(ctx, out_len, max_bits, out));
    k_poly_ntt_store<<<blocks_for(out_len), 128, 0, ctx->stream>>>(buf, (*out)->d, out_len, (uint32_t)max_bits, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_poly_reduce_by_modulus(zkfhe_ctx* ctx, const zkfhe_poly* a, uint64_t modulus, zkfhe_poly** out) {
    if (!ctx || !a || !out) return fail(ctx, ZKFHE_ERR_ARG, "poly_reduce_by_modulus: null pointer");
    if (modulus == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to divide by zero (src/poly.rs:185)");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ZK_TRY(poly_alloc(ctx, a->len, bitlen64(modulus), out));
    k_poly_reduce<<<blocks_for(a->len), 128, 0, ctx->stream>>>(a->d, (*out)->d, a->len, modulus);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_poly_divide_by_cyclo(zkfhe_ctx* ctx, const zkfhe_poly* a, const zkfhe_poly* cyclo, uint64_t modulus,
                               zkfhe_poly** quotient, zkfhe_poly** remainder) {
    if (!ctx || !a || !cyclo || !quotient || !remainder) return fail(ctx, ZKFHE_ERR_ARG, "divide_by_cyclo: null pointer");
    if (modulus == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to divide by zero (src/poly.rs:171)");
    if (cyclo->len < 2) return fail(ctx, ZKFHE_ERR_ARG, "divide_by_cyclo: cyclo must have degree >= 1");
    const uint32_t N = cyclo->len - 1;
    if (a->len > 2 * N) return fail(ctx, ZKFHE_ERR_ARG, "divide_by_cyclo: dividend longer than 2N is not supported");
    if (a->max_bits > 64) return fail(ctx, ZKFHE_ERR_ARG, "divide_by_cyclo: dividend must be reduced (max_bits <= 64)");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* status;
    ZK_TRY(status_word(ctx, &status));
    const uint32_t mb = bitlen64(modulus);
    ZK_TRY(poly_alloc(ctx, N + 1, mb, quotient));
    ZK_TRY(poly_alloc(ctx, 2 * N + 1, mb, remainder));
    if (a->len <= N) {
        // the division loop never runs: quotient is empty -> the reference underflows unless a == 0;
        // handled on the device for uniformity: treat as raw quotient of length 0
        ZK_CUDA(ctx, cudaMemsetAsync((*quotient)->d, 0, (size_t)(N + 1) * 32, ctx->stream));
        ZK_CUDA(ctx, cudaMemsetAsync((*remainder)->d, 0, (size_t)(2 * N + 1) * 32, ctx->stream));
        return fail(ctx, ZKFHE_ERR_ARG, "divide_by_cyclo: dividend of degree < N is not supported");
    }
    k_divide_by_cyclo<<<blocks_for(2 * N + 1), 128, 0, ctx->stream>>>(a->d, a->len, cyclo->d, N, modulus,
                                                                     (*quotient)->d, (*remainder)->d, status);
    ZK_CHECK_LAUNCH(ctx);
    k_divide_by_cyclo_finish<<<1, 1, 0, ctx->stream>>>(status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

uint32_t zkfhe_poly_len(const zkfhe_poly* p) { return p ? p->len : 0; }
uint64_t zkfhe_poly_max_bits(const zkfhe_poly* p) { return p ? p->max_bits : 0; }

int zkfhe_poly_download(zkfhe_ctx* ctx, const zkfhe_poly* p, uint64_t* h_out) {
    if (!ctx || !p || !h_out) return fail(ctx, ZKFHE_ERR_ARG, "poly_download: null pointer");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ZK_CUDA(ctx, cudaMemcpyAsync(h_out, p->d, (size_t)p->len * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

void zkfhe_poly_free(zkfhe_poly* p) {
    if (!p) return;
    if (p->d) {
        if (p->ctx) { cudaSetDevice(p->ctx->device); cudaFreeAsync(p->d, p->ctx->stream); }
        else cudaFree(p->d);
    }
    delete p;
}

// ---- witness ------------------------------------------------------------------------------------
int zkfhe_witness_new(zkfhe_ctx* ctx, uint32_t lookup_bits, zkfhe_witness** out) {
    if (!ctx || !out) return ZKFHE_ERR_ARG;
    if (lookup_bits < 1 || lookup_bits > 16) return fail(ctx, ZKFHE_ERR_ARG, "lookup_bits=%u out of range [1,16]", lookup_bits);
    zkfhe_witness* w = new (std::nothrow) zkfhe_witness();
    if (!w) return fail(ctx, ZKFHE_ERR_CUDA, "out of host memory");
    w->ctx = ctx;
    w->lookup_bits = lookup_bits;
    *out = w;
    return ZKFHE_OK;
}

void zkfhe_witness_free(zkfhe_witness* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    for (int i = 0; i < 3; i++) {
        if (w->adv[i].p) cudaFree(w->adv[i].p);
        if (w->lk[i].p) cudaFree(w->lk[i].p);
        if (w->flags[i].p) cudaFree(w->flags[i].p);
        if (w->copy[i].p) cudaFree(w->copy[i].p);
        if (w->lk_src[i].p) cudaFree(w->lk_src[i].p);
    }
    delete w;
}

int zkfhe_witness_set_recording(zkfhe_witness* w, int on) {
    if (!w) return ZKFHE_ERR_ARG;
    for (int i = 0; i < 3; i++)
        if (w->adv[i].size) return fail(w->ctx, ZKFHE_ERR_STATE, "set_recording: cells were already assigned");
    w->record = on != 0;
    return ZKFHE_OK;
}

int zkfhe_witness_download_structure(zkfhe_witness* w, uint32_t ctx_id, uint8_t* h_flags, uint64_t* h_copy) {
    if (!w || ctx_id > 2) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    if (!w->record) return fail(ctx, ZKFHE_ERR_STATE, "download_structure: the witness is not recording structure");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = w->adv[ctx_id].size;
    if (n && h_flags) ZK_CUDA(ctx, cudaMemcpyAsync(h_flags, w->flags[ctx_id].p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n && h_copy) ZK_CUDA(ctx, cudaMemcpyAsync(h_copy, w->copy[ctx_id].p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_witness_download_lookup_sources(zkfhe_witness* w, uint64_t* h_src) {
    if (!w || !h_src) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    if (!w->record) return fail(ctx, ZKFHE_ERR_STATE, "download_lookup_sources: the witness is not recording structure");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t off = 0;
    for (int i = 0; i < 3; i++) {
        if (w->lk[i].size)
            ZK_CUDA(ctx, cudaMemcpyAsync(h_src + off, w->lk_src[i].p, w->lk[i].size * 8, cudaMemcpyDeviceToHost, ctx->stream));
        off += w->lk[i].size;
    }
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_witness_public_cells(const zkfhe_witness* w, uint64_t* h_cell_ids) {
    if (!w || !h_cell_ids) return ZKFHE_ERR_ARG;
    for (size_t i = 0; i < w->make_public.size(); i++)
        h_cell_ids[i] = cell_id(w->make_public[i].ctx_id, w->make_public[i].offset);
    return ZKFHE_OK;
}

// ---- mock: check every recorded constraint on the device (the reference's `mock` subcommand) ----
namespace zkfhe {
struct MockBases {
    const fr_t* adv[3];
};
__global__ void k_mock_cells(MockBases B, uint32_t ctx_id, uint64_t n, const uint8_t* flags, const uint64_t* copy,
                             int rlc, fr_t gamma, unsigned long long* out /*[0]=violations, [1]=first bad cell id*/) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t* a = B.adv[ctx_id];
    uint8_t f = flags[i];
    bool bad = false;
    if (f & META_SELECTOR) {
        if (rlc) {
            if (i + 2 >= n) bad = true;
            else bad = !eq(add(mul(fe_load(a + i), gamma), fe_load(a + i + 1)), fe_load(a + i + 2));
        } else {
            if (i + 3 >= n) bad = true;
            else bad = !eq(add(fe_load(a + i), mul(fe_load(a + i + 1), fe_load(a + i + 2))), fe_load(a + i + 3));
        }
    }
    fr_t v = fe_load(a + i);
    if ((f & META_ASSERT_ZERO) && !is_zero(v)) bad = true;
    if ((f & META_ASSERT_ONE) && !eq(v, fe_one<FR>())) bad = true;
    if (f & META_COPY_CONFLICT) bad = true;          // a second copy source was dropped: the cell is under-constrained
    uint64_t c = copy[i];
    if (c != CELL_NONE && !eq(v, fe_load(B.adv[cell_ctx(c)] + cell_off(c)))) bad = true;
    if (bad) {
        atomicAdd(out, 1ull);
        atomicMin(out + 1, (unsigned long long)cell_id(ctx_id, i));
    }
}
__global__ void k_mock_lookups(const fr_t* lk, uint64_t n, uint32_t lookup_bits, unsigned long long* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t c = from_mont(fe_load(lk + i));
    bool ok = (c.v[1] | c.v[2] | c.v[3] | c.v[4] | c.v[5] | c.v[6] | c.v[7]) == 0 && (c.v[0] >> lookup_bits) == 0;
    if (!ok) {
        atomicAdd(out, 1ull);
        atomicMin(out + 1, (unsigned long long)(0xFull << 60 | i));
    }
}
}  // namespace zkfhe

int zkfhe_witness_mock(zkfhe_witness* w, uint64_t* n_violations, uint64_t* first_bad_cell) {
    if (!w || !n_violations) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    if (!w->record) return fail(ctx, ZKFHE_ERR_STATE, "mock: the witness is not recording structure");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d_out;
    ZK_TRY(ws_get(ctx, "mock_out", 16, (void**)&d_out));
    unsigned long long init[2] = {0, ~0ull};
    ZK_TRY(upload_async(ctx, d_out, init, 16));
    MockBases B{{w->adv[0].p, w->adv[1].p, w->adv[2].p}};
    for (uint32_t c = 0; c < 3; c++) {
        const uint64_t n = w->adv[c].size;
        if (!n) continue;
        if (c == 2 && !w->have_gamma) return fail(ctx, ZKFHE_ERR_STATE, "mock: RLC cells without a challenge");
        k_mock_cells<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(B, c, n, w->flags[c].p, w->copy[c].p, c == 2,
                                                                        w->gamma, d_out);
        ZK_CHECK_LAUNCH(ctx);
        if (w->lk[c].size) {
            k_mock_lookups<<<(uint32_t)((w->lk[c].size + 255) / 256), 256, 0, ctx->stream>>>(w->lk[c].p, w->lk[c].size,
                                                                                             w->lookup_bits, d_out);
            ZK_CHECK_LAUNCH(ctx);
        }
    }
    unsigned long long res[2];
    ZK_CUDA(ctx, cudaMemcpyAsync(res, d_out, 16, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    *n_violations = res[0];
    if (first_bad_cell) *first_bad_cell = res[1];
    if (res[0]) {
        fail(ctx, ZKFHE_ERR_UNSATISFIED, "mock: %llu constraint violations, first at context %u offset %llu", res[0],
             (unsigned)((res[1] >> 60) - 1), (unsigned long long)(res[1] & ((1ull << 60) - 1)));
        return ZKFHE_ERR_UNSATISFIED;
    }
    return ZKFHE_OK;
}

int zkfhe_witness_reset(zkfhe_witness* w) {
    if (!w) return ZKFHE_ERR_ARG;
    for (int i = 0; i < 3; i++) w->adv[i].size = w->lk[i].size = 0;
    w->make_public.clear();
    w->have_gamma = false;
    return ZKFHE_OK;
}

int zkfhe_chip_from_poly(zkfhe_witness* w, uint32_t ctx_id, const zkfhe_poly* p, zkfhe_assigned_poly* out) {
    W_ENTER(w, ctx_id)
    if (!p || !out) return fail(ctx, ZKFHE_ERR_ARG, "from_poly: null pointer");
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_id, p->len, CellCount{1, 0}, &o, &base));
    k_assign_from_poly<<<blocks_for(p->len), 128, 0, ctx->stream>>>(p->d, p->len, o, status);
    ZK_CHECK_LAUNCH(ctx);
    *out = zkfhe_assigned_poly{ctx_id, 1, base, p->len, 0, p->max_bits};
    return ZKFHE_OK;
}

int zkfhe_chip_load_constant(zkfhe_witness* w, uint32_t ctx_id, uint64_t value, zkfhe_cell* out) {
    W_ENTER(w, ctx_id)
    (void)status;
    if (!out) return fail(ctx, ZKFHE_ERR_ARG, "load_constant: null pointer");
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_id, 1, CellCount{1, 0}, &o, &base));
    k_assign_constant<<<1, 1, 0, ctx->stream>>>(value, o);
    ZK_CHECK_LAUNCH(ctx);
    *out = zkfhe_cell{ctx_id, 0, base};
    return ZKFHE_OK;
}

int zkfhe_chip_to_public(zkfhe_witness* w, const zkfhe_assigned_poly* p) {
    if (!w) return ZKFHE_ERR_ARG;
    ZK_TRY(check_poly(w, p, "to_public"));
    for (uint32_t i = 0; i < p->len; i++) w->make_public.push_back(zkfhe_cell{p->ctx_id, 0, p->base + (uint64_t)i * p->stride});
    return ZKFHE_OK;
}

int zkfhe_chip_set_challenge(zkfhe_witness* w, const uint8_t* h_gamma_fr) {
    if (!w || !h_gamma_fr) return ZKFHE_ERR_ARG;
    memcpy(&w->gamma, h_gamma_fr, 32);
    w->have_gamma = true;
    return ZKFHE_OK;
}

static int rlc_chain(zkfhe_witness* w, uint32_t ctx_rlc, const zkfhe_assigned_poly* p, zkfhe_cell* eval) {
    zkfhe_ctx* ctx = w->ctx;
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_rlc, 1, CellCount{2 * p->len - 1, 0}, &o, &base));
    k_chip_rlc<<<1, RLC_THREADS, 0, ctx->stream>>>(view_of(w, p), p->len, w->gamma, o);
    ZK_CHECK_LAUNCH(ctx);
    *eval = zkfhe_cell{ctx_rlc, 0, base + 2 * (uint64_t)p->len - 2};
    return ZKFHE_OK;
}

int zkfhe_chip_constrain_mul(zkfhe_witness* w, uint32_t ctx_gate, uint32_t ctx_rlc, const zkfhe_assigned_poly* a,
                             const zkfhe_assigned_poly* b, const zkfhe_assigned_poly* c) {
    W_ENTER(w, ctx_gate)
    (void)status;
    if (ctx_rlc > 2 || ctx_rlc == ctx_gate) return fail(ctx, ZKFHE_ERR_ARG, "constrain_mul: bad RLC context");
    ZK_TRY(check_poly(w, a, "constrain_mul(a)"));
    ZK_TRY(check_poly(w, b, "constrain_mul(b)"));
    ZK_TRY(check_poly(w, c, "constrain_mul(c)"));
    if (c->max_num_bits >= P_BITS)
        return fail(ctx, ZKFHE_ERR_OVERFLOW, "assertion failed: c_max_bits < p_bits (src/poly_chip.rs:94)");
    if (!w->have_gamma) return fail(ctx, ZKFHE_ERR_STATE, "constrain_mul: the phase-0 challenge has not been set");
    zkfhe_cell ea, eb, ec;
    ZK_TRY(rlc_chain(w, ctx_rlc, a, &ea));
    ZK_TRY(rlc_chain(w, ctx_rlc, b, &eb));
    ZK_TRY(rlc_chain(w, ctx_rlc, c, &ec));
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, 1, CellCount{4, 0}, &o, &base));
    k_chip_gate4<<<1, 1, 0, ctx->stream>>>(view_of_cell(w, ea), view_of_cell(w, eb), view_of_cell(w, ec), o);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_chip_add(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, const zkfhe_assigned_poly* b,
                   zkfhe_assigned_poly* out) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "add(self)"));
    ZK_TRY(check_poly(w, b, "add(other)"));
    if (!out) return fail(ctx, ZKFHE_ERR_ARG, "add: null output");
    if (b->len < a->len) return fail(ctx, ZKFHE_ERR_ASSERT, "index out of bounds: other.assigned_coefficients[i] (src/poly_chip.rs:129)");
    const uint64_t mb = (a->max_num_bits > b->max_num_bits ? a->max_num_bits : b->max_num_bits) + 1;
    if (mb >= P_BITS) return fail(ctx, ZKFHE_ERR_OVERFLOW, "Risk of overflow detected in add (src/poly_chip.rs:138-141)");
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, CellCount{4, 0}, &o, &base));
    k_chip_add<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), view_of(w, b), a->len, o, status);
    ZK_CHECK_LAUNCH(ctx);
    *out = zkfhe_assigned_poly{ctx_gate, 4, base + 3, a->len, 0, mb};
    return ZKFHE_OK;
}

int zkfhe_chip_scalar_mul(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, const zkfhe_cell* scalar,
                          uint64_t scalar_value, zkfhe_assigned_poly* out) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "scalar_mul"));
    if (!scalar || !out || scalar->ctx_id > 2 || scalar->offset >= w->adv[scalar->ctx_id].size)
        return fail(ctx, ZKFHE_ERR_ARG, "scalar_mul: bad scalar cell");
    const uint64_t mb = a->max_num_bits + bitlen64(scalar_value);
    if (mb >= P_BITS) return fail(ctx, ZKFHE_ERR_OVERFLOW, "Risk of overflow detected in scalar_mul (src/poly_chip.rs:161-164)");
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, CellCount{4, 0}, &o, &base));
    k_chip_scalar_mul<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), view_of_cell(w, *scalar), a->len, o, status);
    ZK_CHECK_LAUNCH(ctx);
    *out = zkfhe_assigned_poly{ctx_gate, 4, base + 3, a->len, 0, mb};
    return ZKFHE_OK;
}

int zkfhe_chip_reduce_by_modulo(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, uint64_t modulus,
                                zkfhe_assigned_poly* out) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "reduce_by_modulo"));
    if (!out) return fail(ctx, ZKFHE_ERR_ARG, "reduce_by_modulo: null output");
    if (modulus == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to divide by zero (halo2-base div_mod)");
    if (a->max_num_bits >= 256) return fail(ctx, ZKFHE_ERR_ARG, "reduce_by_modulo: max_num_bits >= 256");
    const uint32_t lb = w->lookup_bits, nbits = (uint32_t)a->max_num_bits;
    uint32_t bound[8], bound_bits;
    bound_pow2_div(nbits, modulus, bound, &bound_bits);
    CellCount c1 = cc_check_less_than_safe(bound_bits, lb), c2 = cc_check_less_than_safe(bitlen64(modulus), lb);
    CellCount per{4 + c1.cells + c2.cells, c1.lookups + c2.lookups};
    // bound -> Montgomery on the host, passed to the kernel by value (this used to be a one-thread kernel and a
    // synchronising read-back per call: six stream drains per proof in the middle of the witness launches)
    fr_t hb;
    {
        host::Fr c;
        memcpy(c.l, bound, 32);
        const host::Fr m = host::to_mont(c);
        memcpy(hb.v, m.l, 32);
    }
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, per, &o, &base));
    k_chip_div_mod<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), a->len, modulus, hb, bound_bits, lb, o, status);
    ZK_CHECK_LAUNCH(ctx);
    *out = zkfhe_assigned_poly{ctx_gate, per.cells, base, a->len, 0, bitlen64(modulus)};   // rem is cell 0 of each block
    return ZKFHE_OK;
}

int zkfhe_chip_constrain_equality(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                  const zkfhe_assigned_poly* b) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "constrain_equality(self)"));
    ZK_TRY(check_poly(w, b, "constrain_equality(other)"));
    if (b->len < a->len) return fail(ctx, ZKFHE_ERR_ASSERT, "index out of bounds: other.assigned_coefficients[i] (src/poly_chip.rs:260)");
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, CellCount{12, 0}, &o, &base));
    k_chip_is_equal<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), view_of(w, b), a->len, o, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_chip_constrain_coefficients_in_range(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                               uint64_t z, uint64_t y) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "constrain_coefficients_in_range"));
    if (!(z < y)) return fail(ctx, ZKFHE_ERR_ASSERT, "assertion failed: z < y (src/poly_chip.rs:278)");
    const uint32_t lb = w->lookup_bits, yb = bitlen64(y);
    CellCount c1 = cc_check_less_than_safe(yb, lb), c2 = cc_is_less_than(yb, lb);
    CellCount per{c1.cells + 2 * c2.cells + 4 + 8, c1.lookups + 2 * c2.lookups};
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, per, &o, &base));
    k_chip_in_range<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), a->len, z, y, lb, o, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_chip_constrain_from_distribution_chi_key(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                                   uint64_t z) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "constrain_from_distribution_chi_key"));
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, CellCount{20, 0}, &o, &base));
    k_chip_chi_key<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), a->len, z, o, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_chip_constrain_coefficients_in_modulus_field(zkfhe_witness* w, uint32_t ctx_gate,
                                                       const zkfhe_assigned_poly* a, uint64_t modulus) {
    W_ENTER(w, ctx_gate)
    ZK_TRY(check_poly(w, a, "constrain_coefficients_in_modulus_field"));
    if (modulus == 0) return fail(ctx, ZKFHE_ERR_ARG, "modulus must be non-zero");
    CellCount per = cc_check_less_than_safe(bitlen64(modulus), w->lookup_bits);
    OutSpan o;
    uint64_t base;
    ZK_TRY(chip_out(w, ctx_gate, a->len, per, &o, &base));
    k_chip_check_lt_safe<<<blocks_for(a->len), 128, 0, ctx->stream>>>(view_of(w, a), a->len, modulus, w->lookup_bits, o, status);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int zkfhe_chip_safe_trim_leading_zeroes(zkfhe_witness* w, const zkfhe_assigned_poly* a, uint32_t degree,
                                        zkfhe_assigned_poly* out) {
    if (!w) return ZKFHE_ERR_ARG;
    ZK_TRY(check_poly(w, a, "safe_trim_leading_zeroes"));
    if (!out) return fail(w->ctx, ZKFHE_ERR_ARG, "safe_trim_leading_zeroes: null output");
    if (degree > a->len - 1) return fail(w->ctx, ZKFHE_ERR_ASSERT, "assertion failed: degree <= self.degree (src/poly_chip.rs:380)");
    const uint32_t drop = a->len - 1 - degree;
    if (w->record && drop) {     // assert_is_const(coeff, 0) on the trimmed leading coefficients (:382-386)
        zkfhe_ctx* ctx = w->ctx;
        k_meta_assert_zero<<<blocks_for(drop), 128, 0, ctx->stream>>>(w->flags[a->ctx_id].p + a->base, a->stride, drop);
        ZK_CHECK_LAUNCH(ctx);
    }
    *out = zkfhe_assigned_poly{a->ctx_id, a->stride, a->base + (uint64_t)drop * a->stride, degree + 1, 0, a->max_num_bits};
    return ZKFHE_OK;
}

int zkfhe_chip_reduce_by_cyclo(zkfhe_witness* w, uint32_t ctx_gate, uint32_t ctx_rlc, const zkfhe_assigned_poly* self,
                               const zkfhe_assigned_poly* cyclo, const zkfhe_assigned_poly* quotient,
                               const zkfhe_assigned_poly* qtc, const zkfhe_assigned_poly* remainder,
                               uint64_t modulus, zkfhe_assigned_poly* out) {
    if (!w) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    ZK_TRY(check_poly(w, self, "reduce_by_cyclo(self)"));
    ZK_TRY(check_poly(w, cyclo, "reduce_by_cyclo(cyclo)"));
    ZK_TRY(check_poly(w, quotient, "reduce_by_cyclo(quotient)"));
    ZK_TRY(check_poly(w, qtc, "reduce_by_cyclo(quotient_times_cyclo)"));
    ZK_TRY(check_poly(w, remainder, "reduce_by_cyclo(remainder)"));
    const uint64_t mbits = bitlen64(modulus);
    if (quotient->max_num_bits > mbits) return fail(ctx, ZKFHE_ERR_ASSERT, "assertion failed: quotient.max_num_bits <= modulus_bits (src/poly_chip.rs:196)");
    if (remainder->max_num_bits > mbits) return fail(ctx, ZKFHE_ERR_ASSERT, "assertion failed: remainder.max_num_bits <= modulus_bits (src/poly_chip.rs:197)");
    const uint64_t mx = qtc->max_num_bits > remainder->max_num_bits ? qtc->max_num_bits : remainder->max_num_bits;
    if (mx + 1 >= P_BITS) return fail(ctx, ZKFHE_ERR_OVERFLOW, "assertion failed: max(...) + 1 < p_bits (src/poly_chip.rs:201)");
    const uint32_t cyclo_deg = cyclo->len - 1;
    ZK_TRY(zkfhe_chip_constrain_mul(w, ctx_gate, ctx_rlc, quotient, cyclo, qtc));                 // :205
    zkfhe_assigned_poly sum, sum_mod, sum_trim;
    ZK_TRY(zkfhe_chip_add(w, ctx_gate, qtc, remainder, &sum));                                    // :208
    ZK_TRY(zkfhe_chip_reduce_by_modulo(w, ctx_gate, &sum, modulus, &sum_mod));                    // :211
    ZK_TRY(zkfhe_chip_safe_trim_leading_zeroes(w, &sum_mod, self->len - 1, &sum_trim));           // :214
    ZK_TRY(zkfhe_chip_constrain_equality(w, ctx_gate, &sum_trim, self));                          // :217
    if (cyclo_deg == 0) return fail(ctx, ZKFHE_ERR_ASSERT, "attempt to subtract with overflow: cyclo_deg - 1 (src/poly_chip.rs:222)");
    return zkfhe_chip_safe_trim_leading_zeroes(w, remainder, cyclo_deg - 1, out);                 // :222
}

int zkfhe_witness_counts(const zkfhe_witness* w, uint64_t advice_cells[3], uint64_t* lookup_cells, uint64_t* instances) {
    if (!w) return ZKFHE_ERR_ARG;
    uint64_t l = 0;
    for (int i = 0; i < 3; i++) {
        if (advice_cells) advice_cells[i] = w->adv[i].size;
        l += w->lk[i].size;
    }
    if (lookup_cells) *lookup_cells = l;
    if (instances) *instances = w->make_public.size();
    return ZKFHE_OK;
}

int zkfhe_witness_device_ptr(zkfhe_witness* w, uint32_t which, uint8_t** d_ptr) {
    if (!w || !d_ptr) return ZKFHE_ERR_ARG;
    if (which > 2) return fail(w->ctx, ZKFHE_ERR_ARG, "device_ptr: only advice contexts 0..2 have a stable buffer");
    *d_ptr = (uint8_t*)w->adv[which].p;
    return ZKFHE_OK;
}

int zkfhe_witness_download(zkfhe_witness* w, uint32_t which, uint8_t* h_out) {
    if (!w || !h_out) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (which <= 2) {
        if (w->adv[which].size)
            ZK_CUDA(ctx, cudaMemcpyAsync(h_out, w->adv[which].p, w->adv[which].size * 32, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (which == 3) {
        size_t off = 0;
        for (int i = 0; i < 3; i++) {
            if (w->lk[i].size)
                ZK_CUDA(ctx, cudaMemcpyAsync(h_out + off, w->lk[i].p, w->lk[i].size * 32, cudaMemcpyDeviceToHost, ctx->stream));
            off += w->lk[i].size * 32;
        }
    } else if (which == 4) {
        const size_t n = w->make_public.size();
        if (n) {
            std::vector<uint64_t> co(n);
            for (size_t i = 0; i < n; i++) co[i] = ((uint64_t)w->make_public[i].ctx_id << 60) | w->make_public[i].offset;
            const fr_t* bases[3] = {w->adv[0].p, w->adv[1].p, w->adv[2].p};
            uint8_t* d;
            ZK_TRY(ws_get(ctx, "gather", n * 8 + 64 + n * 32, (void**)&d));
            uint64_t* d_co = (uint64_t*)(d + 64);
            fr_t* d_out = (fr_t*)(d + 64 + ((n * 8 + 31) / 32) * 32);
            ZK_TRY(ws_get(ctx, "gather", 64 + ((n * 8 + 31) / 32) * 32 + n * 32, (void**)&d));
            d_co = (uint64_t*)(d + 64);
            d_out = (fr_t*)(d + 64 + ((n * 8 + 31) / 32) * 32);
            ZK_CUDA(ctx, cudaMemcpyAsync(d, bases, sizeof bases, cudaMemcpyHostToDevice, ctx->stream));
            ZK_CUDA(ctx, cudaMemcpyAsync(d_co, co.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
            k_gather_cells<<<blocks_for((uint32_t)n), 128, 0, ctx->stream>>>((const fr_t* const*)d, d_co, d_out, (uint32_t)n);
            ZK_CHECK_LAUNCH(ctx);
            ZK_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        }
    } else {
        return fail(ctx, ZKFHE_ERR_ARG, "download: which=%u out of range", which);
    }
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

}  // extern "C"
