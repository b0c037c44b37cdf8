// Batched radix-2 NTT over BN254 Fr for sm_100a (stage (3) of the prove path).
//
// Replaces halo2-axiom `arithmetic::best_fft` and the scaling done by
// `EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff}`
// [UPSTREAM, un-vendored; SURVEY.md §8 a20, App. C].  Semantics (natural order
// in and out, inverse scaled by n^-1, extended coset zeta*H) follow
// oracle/ntt.py, which is what the parity tests compare against.
//
// Shape.  A transform of n = R*C points is done as the classic 4-step:
//   pass A  C independent R-point transforms down the columns of the R x C
//           matrix (input index i1*C + i2), then the twiddle w_n^(i2*k1);
//   pass B  R independent C-point transforms along the rows, stored transposed
//           (output index k1 + R*k2).
// A CTA owns a tile of T = 1024 (or 2048) field elements: in pass A R rows x L
// adjacent columns (L*32-byte contiguous runs), in pass B L whole rows.  n <= 2048
// is a single pass-B launch with L = 1.
//
// Inside a pass the log2(R) radix-2 DIT stages are grouped into ROUNDS done in
// registers (NTT_MAX_LR stages each: radix-4 = 4 elements, 4 butterflies, 3
// twiddles per thread and group; radix-8 is compiled by setting NTT_MAX_LR = 3).
// The first round reads its operands straight from HBM (bit-reversed gather) and
// the last one writes straight back, so a 6-stage pass exchanges data through
// shared memory twice and a 7/8-stage pass three times, instead of once per stage; an element
// moves HBM->SM->HBM once per pass: 2 x 64 B per element per transform against
// 64 B algorithmic.  Shared memory holds the tile as two 16-byte planes (low and
// high halves of every element) so that a quarter-warp's 128-bit accesses cover
// 32 distinct banks.
//
// Fused into the passes: zero extension of a short input (coeff_to_extended),
// the coset pre-multiplication zeta^(i mod 3), the n^-1 scaling of the inverse
// (folded into the 4-step twiddle table for two-pass transforms) and the coset
// post-multiplication zeta^-(i mod 3).
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"

namespace zkfhe {

__global__ void k_build_twiddles(fr_t* tw, uint32_t log_n, int inverse) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_n)) return;
    fr_t w = inverse ? fr_t{ZKFHE_FR_ROOT_OF_UNITY_INV_MONT} : fr_t{ZKFHE_FR_ROOT_OF_UNITY_MONT};
    for (uint32_t s = log_n; s < 28; s++) w = sqr(w);
    fe_store(tw + i, pow_u64(w, i));
}

__global__ void k_n_inv(fr_t* out, uint32_t log_n) {
    // (2^log_n)^-1 = ((r+1)/2)^log_n ; computed by Fermat for simplicity
    fr_t two = add(fe_one<FR>(), fe_one<FR>());
    fr_t x = fe_one<FR>();
    for (uint32_t i = 0; i < log_n; i++) x = mul(x, two);
    fe_store(out, inv(x));
}

struct NttPass {
    const fr_t* in;
    fr_t* out;
    const fr_t* tw;                   // w^i (forward) or w^-i (inverse), i < n: butterfly twiddles
    const fr_t* tw4;                  // 4-step twiddles of pass A: tw, or w^-i * n^-1 for the inverse
    uint64_t in_stride, out_stride;   // elements between consecutive batch columns
    uint32_t log_n, log_r, log_l;
    uint32_t mode;                    // 0: pass A (strided columns), 1: pass B / single (rows)
    uint32_t in_len;                  // input elements with index >= in_len read as zero
    uint32_t pre_coset, post_scale, post_coset;
    uint32_t n_rounds;
    uint32_t zero_stages;             // leading stages of round 0 whose second operand is structurally zero (zero-extended input)
    uint8_t rounds[8];                // stages per round (3, 2, 1; a trailing 0 = plain copy-out)
    fr_t n_inv, zeta, zeta2;
};

// stages per register round.  Measured on B200: 3 = radix-8 (114 registers, 4 CTAs of 128 threads per SM, one shared-
// memory exchange per 6-stage pass) against 2 = radix-4 (80 registers, 6 CTAs per SM, one more exchange per pass):
// radix-4 is 6-9 % faster at 2^13 .. 2^18 -- the pass is bound by dependent-IMAD latency, so warps in flight matter
// more than shared-memory round trips.
static constexpr int NTT_MAX_LR = 2;

extern __shared__ uint4 ntt_smem[];

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

__device__ __forceinline__ fr_t smem_load(const uint4* lo, const uint4* hi, uint32_t i) {
    uint4 a = lo[i], b = hi[i];
    fr_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void smem_store(uint4* lo, uint4* hi, uint32_t i, const fr_t& v) {
    lo[i] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    hi[i] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// LR radix-2 DIT stages (stage indices s .. s+LR-1 of the sub-transform) on 2^LR elements held in
// registers: x[a] sits at transform position t0 + a * 2^s, j = t0 mod 2^s.
template <int LR>
__device__ __forceinline__ void dit_round(fr_t* x, uint32_t j, uint32_t s, uint32_t log_n, const fr_t* __restrict__ tw,
                                          uint32_t zero_stages) {
#pragma unroll
    for (int i = 0; i < LR; i++) {
        const uint32_t h = 1u << i;
#pragma unroll
        for (int a = 0; a < (1 << LR); a++) {
            if (a & h) continue;
            if ((uint32_t)i < zero_stages) { x[a + h] = x[a]; continue; }     // u + 0*w = u - 0*w = u
            const uint32_t lowa = (uint32_t)a & (h - 1);
            fr_t v = x[a + h];
            // exponent of w_(2^(s+i+1)); it is zero for every lane exactly when s == 0 and lowa == 0
            if (!(s == 0 && lowa == 0)) v = mul(v, fe_load_nc(tw + ((uint64_t)(j + (lowa << s)) << (log_n - s - i - 1))));
            const fr_t u = x[a];
            x[a] = add(u, v);
            x[a + h] = sub(u, v);
        }
    }
}

// TMA = true (experiment, ZKFHE_NTT_TMA=1): the pass-A tile was brought into shared memory by one cp.async.bulk.tensor
// (`landing`: R rows x L elements, 32-byte elements, row-major; rows beyond the input are zero-filled by the copy engine)
// and round 0 gathers its bit-reversed operands from there instead of from global memory.
template <int LR, bool TMA = false>
__device__ __forceinline__ void ntt_round(const NttPass& p, uint4* Slo, uint4* Shi, uint32_t s, bool first, bool last,
                                          const fr_t* __restrict__ in, fr_t* __restrict__ out, const fr_t* landing = nullptr) {
    const uint32_t T = 1u << (p.log_r + p.log_l), lmask = (1u << p.log_l) - 1;
    const uint32_t log_c = p.log_n - p.log_r;          // pass A: columns; pass B: rows of the matrix
    const uint32_t base = blockIdx.x << p.log_l;       // first column (A) / first row (B) of this tile
    for (uint32_t q = threadIdx.x; q < (T >> LR); q += blockDim.x) {
        const uint32_t l = q & lmask, u = q >> p.log_l;
        const uint32_t j = u & ((1u << s) - 1);
        const uint32_t t0 = ((u >> s) << (s + LR)) + j;
        fr_t x[1 << LR];
#pragma unroll
        for (int a = 0; a < (1 << LR); a++) {
            const uint32_t t = t0 + ((uint32_t)a << s);
            if (first) {
                const uint32_t tn = brev(t, p.log_r);                 // DIT consumes its input in bit-reversed order
                const uint32_t g = p.mode == 0 ? (tn << log_c) + base + l : ((base + l) << p.log_r) + tn;
                if (g < p.in_len) {
                    fr_t v = TMA ? fe_load(landing + ((tn << p.log_l) + l)) : fe_load(in + g);
                    if (p.pre_coset) {
                        const uint32_t m3 = g % 3u;
                        if (m3 == 1) v = mul(v, p.zeta);
                        else if (m3 == 2) v = mul(v, p.zeta2);
                    }
                    x[a] = v;
                } else {
                    x[a] = fe_zero<FR>();
                }
            } else {
                x[a] = smem_load(Slo, Shi, (t << p.log_l) + l);
            }
        }
        dit_round<LR>(x, j, s, p.log_n, p.tw, first ? p.zero_stages : 0);
#pragma unroll
        for (int a = 0; a < (1 << LR); a++) {
            const uint32_t t = t0 + ((uint32_t)a << s);
            if (last) {
                fr_t v = x[a];
                uint32_t g;
                if (p.mode == 0) {
                    const uint32_t col = base + l;
                    g = (t << log_c) + col;
                    const uint32_t tw_idx = col * t;            // < n
                    if (tw_idx || p.tw4 != p.tw) v = mul(v, fe_load_nc(p.tw4 + tw_idx));
                } else {
                    g = (base + l) + (t << log_c);
                    if (p.post_coset) {
                        const uint32_t m3 = g % 3u;             // zeta^-g = zeta^((3 - g%3) % 3)
                        if (m3 == 1) v = mul(v, p.zeta2);
                        else if (m3 == 2) v = mul(v, p.zeta);
                    }
                    if (p.post_scale) v = mul(v, p.n_inv);
                }
                fe_store(out + g, v);
            } else {
                smem_store(Slo, Shi, (t << p.log_l) + l, x[a]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_ntt_pass(const NttPass p) {
    const uint32_t T = 1u << (p.log_r + p.log_l);
    uint4* Slo = ntt_smem;
    uint4* Shi = ntt_smem + T;
    const fr_t* in = p.in + (uint64_t)blockIdx.y * p.in_stride;
    fr_t* out = p.out + (uint64_t)blockIdx.y * p.out_stride;
    uint32_t s = 0;
    for (uint32_t ri = 0; ri < p.n_rounds; ri++) {
        const uint32_t r = p.rounds[ri];
        const bool first = ri == 0, last = ri + 1 == p.n_rounds;
        if (NTT_MAX_LR >= 3 && r == 3) ntt_round<NTT_MAX_LR>(p, Slo, Shi, s, first, last, in, out);
        else if (r == 2) ntt_round<2>(p, Slo, Shi, s, first, last, in, out);
        else if (r == 1) ntt_round<1>(p, Slo, Shi, s, first, last, in, out);
        else ntt_round<0>(p, Slo, Shi, s, first, last, in, out);
        s += r;
        if (!last) __syncthreads();
    }
}

// ---- experiment: pass A with the tile staged by TMA (cp.async.bulk.tensor + mbarrier) -------------------------------
// VERDICT r1 asked for the measurement: does staging the R x L tile through the copy engine beat the bit-reversed gather
// of round 0 straight from global memory?  The tile needs its own landing buffer (round 0 permutes across threads, so it
// cannot run in place without holding every operand in registers), i.e. 64 KB instead of 32 KB of shared memory per CTA
// and 3 instead of 6 CTAs per SM.  Selected with ZKFHE_NTT_TMA=1; results in profiles/ and DESIGN.md.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) k_ntt_pass_tma(const NttPass p, const __grid_constant__ CUtensorMap tmap) {
    const uint32_t T = 1u << (p.log_r + p.log_l);
    uint4* Slo = ntt_smem;
    uint4* Shi = ntt_smem + T;
    // the copy engine wants a 128-byte aligned destination: round up behind the two planes (128 spare bytes are allocated)
    fr_t* landing = reinterpret_cast<fr_t*>((reinterpret_cast<uintptr_t>(ntt_smem + 2 * T) + 127) & ~(uintptr_t)127);
    __shared__ alignas(8) uint64_t mbar;
    fr_t* out = p.out + (uint64_t)blockIdx.y * p.out_stride;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = T * (uint32_t)sizeof(fr_t);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        // coordinates: (32-bit word inside the row, row, batch column)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(landing)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((blockIdx.x << p.log_l) * 8u), "r"(0u),
                       "r"(blockIdx.y), "r"(smem_u32(&mbar))
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t s = 0;
    for (uint32_t ri = 0; ri < p.n_rounds; ri++) {
        const uint32_t r = p.rounds[ri];
        const bool first = ri == 0, last = ri + 1 == p.n_rounds;
        if (r == 2) ntt_round<2, true>(p, Slo, Shi, s, first, last, nullptr, out, landing);
        else if (r == 1) ntt_round<1, true>(p, Slo, Shi, s, first, last, nullptr, out, landing);
        else ntt_round<0, true>(p, Slo, Shi, s, first, last, nullptr, out, landing);
        s += r;
        if (!last) __syncthreads();
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return (EncodeTiledFn)sym;
    }();
    return fn;
}
// pass A over TMA when the shape allows it (tile rows and 32-byte row segments within the 256-element box limit, the
// input a whole number of matrix rows); returns false to fall back to the plain kernel
static bool try_launch_pass_a_tma(zkfhe_ctx* ctx, const NttPass& p, uint32_t tiles, uint32_t batch, int* rc) {
    const uint32_t R = 1u << p.log_r, L = 1u << p.log_l, C = 1u << (p.log_n - p.log_r);
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || p.mode != 0 || R > 256 || L * 8 > 256 || p.in_len % C || p.in == p.out) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)C * 8, (cuuint64_t)(p.in_len / C), batch};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 32, (cuuint64_t)p.in_stride * 32};      // bytes, dims 1 and 2
    const cuuint32_t box[3] = {L * 8, R, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap tmap;
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)p.in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    const uint32_t T = R * L;
    const size_t smem = 2 * (size_t)T * sizeof(fr_t) + 128;
    cudaFuncSetAttribute(k_ntt_pass_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2048 * (int)sizeof(fr_t) + 128);
    if (T > 2048) return false;
    const uint32_t threads = T / 8 < 32 ? 32 : (T / 8 > 128 ? 128 : T / 8);
    k_ntt_pass_tma<<<dim3(tiles, batch), threads, smem, ctx->stream>>>(p, tmap);
    ctx->launches++;
    *rc = cudaGetLastError() == cudaSuccess ? ZKFHE_OK : fail(ctx, ZKFHE_ERR_CUDA, "k_ntt_pass_tma launch failed");
    return true;
}

// tw4[i] = w^-i * n^-1: the inverse transform's scaling rides on the 4-step twiddle
__global__ void k_scale_twiddles(const fr_t* tw, fr_t* out, uint32_t n, const fr_t* n_inv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_store(out + i, mul(fe_load(tw + i), fe_load(n_inv)));
}

int ntt_domain(zkfhe_ctx* ctx, uint32_t log_n, NttDomain** out) {
    auto it = ctx->domains.find(log_n);
    if (it != ctx->domains.end()) { *out = &it->second; return ZKFHE_OK; }
    NttDomain d;
    size_t n = (size_t)1 << log_n;
    fr_t* d_ninv = nullptr;
    ZK_CUDA(ctx, cudaMalloc(&d.tw_fwd, n * sizeof(fr_t)));
    ZK_CUDA(ctx, cudaMalloc(&d.tw_inv, n * sizeof(fr_t)));
    ZK_CUDA(ctx, cudaMalloc(&d_ninv, sizeof(fr_t)));
    uint32_t blocks = (uint32_t)((n + 255) / 256);
    k_build_twiddles<<<blocks, 256, 0, ctx->stream>>>(d.tw_fwd, log_n, 0);
    ZK_CHECK_LAUNCH(ctx);
    k_build_twiddles<<<blocks, 256, 0, ctx->stream>>>(d.tw_inv, log_n, 1);
    ZK_CHECK_LAUNCH(ctx);
    k_n_inv<<<1, 1, 0, ctx->stream>>>(d_ninv, log_n);
    ZK_CHECK_LAUNCH(ctx);
    ZK_CUDA(ctx, cudaMalloc(&d.tw_inv_s, n * sizeof(fr_t)));
    k_scale_twiddles<<<blocks, 256, 0, ctx->stream>>>(d.tw_inv, d.tw_inv_s, (uint32_t)n, d_ninv);
    ZK_CHECK_LAUNCH(ctx);
    ZK_CUDA(ctx, cudaMemcpyAsync(&d.n_inv, d_ninv, sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    ZK_CUDA(ctx, cudaFree(d_ninv));
    ctx->domains[log_n] = d;
    *out = &ctx->domains[log_n];
    return ZKFHE_OK;
}

static int launch_pass(zkfhe_ctx* ctx, NttPass p, uint32_t tiles, uint32_t batch) {
    const uint32_t T = 1u << (p.log_r + p.log_l);
    // rounds of three stages; a remainder of one is taken as 2 + 2 so no round is a lone radix-2
    uint32_t left = p.log_r, nr = 0;
    while (NTT_MAX_LR >= 3 && (left > 4 || left == 3)) { p.rounds[nr++] = 3; left -= 3; }
    while (NTT_MAX_LR < 3 && left > 2) { p.rounds[nr++] = 2; left -= 2; }
    if (left == 4) { p.rounds[nr++] = 2; p.rounds[nr++] = 2; }
    else if (left) p.rounds[nr++] = (uint8_t)left;
    if (nr < 2) p.rounds[nr++] = 0;      // never load and store in the same round: an in-place pass would race
    p.n_rounds = nr;
    // pass A of a zero-extended transform: rows i1 >= in_len / C are zero, i.e. (bit-reversed) every DIT position
    // whose low z bits are not all zero -- the first z stages only copy
    p.zero_stages = 0;
    if (p.mode == 0) {
        const uint32_t log_c = p.log_n - p.log_r;
        uint32_t z = 0;
        while (z < p.rounds[0] && z < p.log_r && ((uint64_t)p.in_len << (z + 1)) <= ((uint64_t)1 << p.log_n) &&
               (p.in_len >> log_c) << log_c == p.in_len)
            z++;
        p.zero_stages = z;
    }
    // 128 threads per 1024-element tile (two radix-4 groups per thread and round): 80 registers, 6 CTAs per SM;
    // 256-thread CTAs (one group per thread) measured 2-5 % slower
    const uint32_t threads = T / 8 < 32 ? 32 : (T / 8 > 128 ? 128 : T / 8);
    const size_t smem = (size_t)T * sizeof(fr_t);
    // process-wide attribute: always the fixed maximum (4096-element tile), never this call's size
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int)sizeof(fr_t)));
    {   // field products this pass issues (zkfhe_timing_get category 6), in eighths per element: one per butterfly,
        // minus the butterflies of round 0 whose twiddle is 1 or whose operand is structurally zero, plus the
        // per-element 4-step twiddle / coset / n^-1 factors
        const uint64_t elems = (uint64_t)batch << p.log_n;
        const uint32_t r0 = p.rounds[0];
        uint32_t skipped8 = r0 == 3 ? 7 : r0 == 2 ? 6 : r0 == 1 ? 4 : 0;
        if (r0 == 3 && p.zero_stages == 2) skipped8 = 9;
        else if (r0 == 3 && p.zero_stages == 3) skipped8 = 12;
        else if (r0 == 2 && p.zero_stages == 2) skipped8 = 8;
        uint64_t eighths = 4ull * p.log_r - skipped8;
        if (p.mode == 0) eighths += 8;
        if (p.post_scale) eighths += 8;
        if (p.post_coset) eighths += 16 / 3;
        uint64_t prod = elems * eighths / 8;
        if (p.pre_coset) prod += (uint64_t)batch * (p.in_len < (1u << p.log_n) ? p.in_len : (1u << p.log_n)) * 2 / 3;
        ctx->ntt_products += prod;
    }
    {
        static const bool use_tma = [] { const char* e = getenv("ZKFHE_NTT_TMA"); return e && atoi(e) != 0; }();
        int rc = ZKFHE_OK;
        if (use_tma && try_launch_pass_a_tma(ctx, p, tiles, batch, &rc)) return rc;
    }
    dim3 grid(tiles, batch);
    k_ntt_pass<<<grid, threads, smem, ctx->stream>>>(p);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

int ntt_run(zkfhe_ctx* ctx, const fr_t* d_in, uint64_t in_stride, uint32_t in_len, fr_t* d_out,
            uint64_t out_stride, uint32_t log_n, uint32_t batch, int inverse, int coset) {
    if (log_n < 1 || log_n > 22) return fail(ctx, ZKFHE_ERR_ARG, "ntt: log_n=%u out of range [1,22]", log_n);
    if (batch == 0) return ZKFHE_OK;
    if (batch > 65535) return fail(ctx, ZKFHE_ERR_ARG, "ntt: batch=%u > 65535", batch);
    NttDomain* dom;
    ZK_TRY(ntt_domain(ctx, log_n, &dom));
    NttPass p{};
    p.tw = inverse ? dom->tw_inv : dom->tw_fwd;
    p.tw4 = inverse ? dom->tw_inv_s : dom->tw_fwd;
    p.log_n = log_n;
    p.n_inv = dom->n_inv;
    p.zeta = fr_t{ZKFHE_FR_ZETA_MONT};
    p.zeta2 = fr_t{ZKFHE_FR_ZETA2_MONT};
    timed_call_start(ctx);
    ZK_TRY(timed_begin(ctx, ZK_CAT_NTT, (uint64_t)batch << log_n));
    // A lone small transform (stage (1): two or one 2N-point transforms per Poly::mul) on one CTA is a 90-butterfly
    // dependent walk per thread (~80 us at 2^11); cut into 128-element tiles it is two launches of 16 one-warp CTAs.
    const bool small_tiles = log_n >= 8 && log_n <= 11 && ((uint64_t)batch << log_n) <= 8192;
    if (log_n <= 11 && !small_tiles) {
        p.in = d_in; p.out = d_out; p.in_stride = in_stride; p.out_stride = out_stride;
        p.log_r = log_n; p.log_l = 0; p.mode = 1; p.in_len = in_len;
        p.pre_coset = (!inverse && coset); p.post_scale = inverse; p.post_coset = (inverse && coset);
        ZK_TRY(launch_pass(ctx, p, 1, batch));
    } else {
        // tile of 1024 elements (3 CTAs of 128 threads per SM).  Measured on B200: 2048- and 4096-element tiles give
        // longer contiguous runs in pass A but lose 10 % and 55 % to occupancy, at 2^13 .. 2^18 alike.
        uint32_t log_t = 10;
        if (const char* e = getenv("ZKFHE_NTT_LOG_T")) log_t = (uint32_t)atoi(e);      // tuning knob (tools/bench_kernels.py)
        if (log_t < 10) log_t = 10;
        if (log_t > 12) log_t = 12;
        if (small_tiles) log_t = 7;
        // few columns of 2^12 .. 2^16 (the SHPLONK round, single-column transforms): 256-element tiles on one-warp CTAs
        // put 32+ CTAs on the GPU instead of 8 (25-49 us per launch with 1024-element tiles)
        else if (((uint64_t)batch << log_n) <= 65536 && !getenv("ZKFHE_NTT_LOG_T")) log_t = 8;
        if (log_t < (log_n + 1) / 2) log_t = (log_n + 1) / 2;          // a tile holds at least one whole pass-A column
        const uint32_t log_ra = (log_n + 1) / 2, log_c = log_n - log_ra;
        fr_t* tmp;
        ZK_TRY(ws_get(ctx, "ntt_tmp", ((size_t)batch << log_n) * sizeof(fr_t), (void**)&tmp));
        // pass A: in -> tmp
        p.in = d_in; p.out = tmp; p.in_stride = in_stride; p.out_stride = (uint64_t)1 << log_n;
        p.log_r = log_ra; p.log_l = log_t - log_ra < log_c ? log_t - log_ra : log_c; p.mode = 0; p.in_len = in_len;
        p.pre_coset = (!inverse && coset); p.post_scale = 0; p.post_coset = 0;
        ZK_TRY(launch_pass(ctx, p, 1u << (log_c - p.log_l), batch));
        // pass B: tmp -> out
        p.in = tmp; p.out = d_out; p.in_stride = (uint64_t)1 << log_n; p.out_stride = out_stride;
        p.log_r = log_c; p.log_l = log_t - log_c < log_ra ? log_t - log_c : log_ra; p.mode = 1; p.in_len = 1u << log_n;
        p.pre_coset = 0; p.post_scale = 0 /* n^-1 already applied by pass A's twiddles */; p.post_coset = (inverse && coset);
        ZK_TRY(launch_pass(ctx, p, 1u << (log_ra - p.log_l), batch));
    }
    ZK_TRY(timed_end(ctx));
    return ZKFHE_OK;
}

}  // namespace zkfhe
