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
// Each CTA keeps a tile of T = 1024 (or 2048) field elements in shared memory
// and runs all of a pass's butterfly stages there, so one element moves
// HBM->SM->HBM once per pass: 2 x 64 B per element per transform against 64 B
// algorithmic.  In pass A a tile is R rows x L adjacent columns (L*32-byte
// contiguous runs); in pass B it is L whole rows (C*32-byte runs in, L*32-byte
// runs out).  n <= 2048 is a single pass-B launch with L = 1.
//
// Fused into the passes: zero extension of a short input (coeff_to_extended),
// the coset pre-multiplication zeta^(i mod 3), the n^-1 scaling of the inverse
// and the coset post-multiplication zeta^-(i mod 3).
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
    const fr_t* tw;
    uint64_t in_stride, out_stride;   // elements between consecutive batch columns
    uint32_t log_n, log_r, log_l;
    uint32_t mode;                    // 0: pass A (strided columns), 1: pass B / single (rows)
    uint32_t in_len;                  // input elements with index >= in_len read as zero
    uint32_t pre_coset, post_scale, post_coset;
    fr_t n_inv, zeta, zeta2;
};

extern __shared__ uint4 ntt_smem[];

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

__global__ void __launch_bounds__(1024) k_ntt_pass(const NttPass p) {
    fr_t* S = reinterpret_cast<fr_t*>(ntt_smem);
    const uint32_t R = 1u << p.log_r, L = 1u << p.log_l, T = R << p.log_l;
    const uint32_t n = 1u << p.log_n;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const fr_t* in = p.in + (uint64_t)blockIdx.y * p.in_stride;
    fr_t* out = p.out + (uint64_t)blockIdx.y * p.out_stride;
    const uint32_t lmask = L - 1;
    const uint32_t log_c = p.log_n - p.log_r;          // pass A: columns; pass B: rows of the matrix
    const uint32_t base = blockIdx.x << p.log_l;       // first column (A) / first row (B) of this tile

    // ---- load tile, bit-reversing the transform index ---------------------------------
    for (uint32_t e = tid; e < T; e += nt) {
        uint32_t t, l, g;
        if (p.mode == 0) { l = e & lmask; t = e >> p.log_l; g = (t << log_c) + base + l; }
        else             { t = e & (R - 1); l = e >> p.log_r; g = ((base + l) << p.log_r) + t; }
        fr_t v;
        if (g < p.in_len) {
            v = fe_load(in + g);
            if (p.pre_coset) {
                uint32_t m3 = g % 3u;
                if (m3 == 1) v = mul(v, p.zeta);
                else if (m3 == 2) v = mul(v, p.zeta2);
            }
        } else {
            v = fe_zero<FR>();
        }
        fe_store(S + ((brev(t, p.log_r) << p.log_l) + l), v);
    }
    __syncthreads();

    // ---- log_r radix-2 DIT stages in shared memory -------------------------------------
    const uint32_t half = T >> 1;
    for (uint32_t s = 0; s < p.log_r; s++) {
        const uint32_t m = 1u << s;
        for (uint32_t b = tid; b < half; b += nt) {
            uint32_t l = b & lmask;
            uint32_t bf = b >> p.log_l;
            uint32_t j = bf & (m - 1);
            uint32_t i0 = ((bf >> s) << (s + 1)) + j;
            fr_t* p0 = S + ((i0 << p.log_l) + l);
            fr_t* p1 = p0 + (m << p.log_l);
            fr_t u = fe_load(p0);
            fr_t v = fe_load(p1);
            if (j) v = mul(v, fe_load_nc(p.tw + ((uint64_t)j << (p.log_n - s - 1))));
            fe_store(p0, add(u, v));
            fe_store(p1, sub(u, v));
        }
        __syncthreads();
    }

    // ---- store ------------------------------------------------------------------------
    for (uint32_t e = tid; e < T; e += nt) {
        uint32_t l = e & lmask, t = e >> p.log_l, g;
        fr_t v = fe_load(S + e);
        if (p.mode == 0) {
            uint32_t col = base + l;
            g = (t << log_c) + col;
            uint32_t tw_idx = col * t;                  // < n
            if (tw_idx) v = mul(v, fe_load_nc(p.tw + tw_idx));
        } else {
            g = (base + l) + (t << log_c);
            if (p.post_coset) {
                uint32_t m3 = g % 3u;                   // zeta^-g = zeta^((3 - g%3) % 3)
                if (m3 == 1) v = mul(v, p.zeta2);
                else if (m3 == 2) v = mul(v, p.zeta);
            }
            if (p.post_scale) v = mul(v, p.n_inv);
        }
        fe_store(out + g, v);
    }
    (void)n;
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
    ZK_CUDA(ctx, cudaMemcpyAsync(&d.n_inv, d_ninv, sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZK_CUDA(ctx, cudaFree(d_ninv));
    ctx->domains[log_n] = d;
    *out = &ctx->domains[log_n];
    return ZKFHE_OK;
}

static int launch_pass(zkfhe_ctx* ctx, const NttPass& p, uint32_t tiles, uint32_t batch) {
    uint32_t T = 1u << (p.log_r + p.log_l);
    uint32_t threads = T / 2 < 32 ? 32 : (T / 2 > 1024 ? 1024 : T / 2);
    size_t smem = (size_t)T * sizeof(fr_t);
    // process-wide attribute: always the fixed maximum (2048-element tile), never this call's size
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * (int)sizeof(fr_t)));
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
    p.log_n = log_n;
    p.n_inv = dom->n_inv;
    p.zeta = fr_t{ZKFHE_FR_ZETA_MONT};
    p.zeta2 = fr_t{ZKFHE_FR_ZETA2_MONT};
    timed_call_start(ctx);
    ZK_TRY(timed_begin(ctx, ZK_CAT_NTT, (uint64_t)batch << log_n));
    if (log_n <= 11) {
        p.in = d_in; p.out = d_out; p.in_stride = in_stride; p.out_stride = out_stride;
        p.log_r = log_n; p.log_l = 0; p.mode = 1; p.in_len = in_len;
        p.pre_coset = (!inverse && coset); p.post_scale = inverse; p.post_coset = (inverse && coset);
        ZK_TRY(launch_pass(ctx, p, 1, batch));
    } else {
        const uint32_t log_t = log_n > 20 ? 11 : 10;
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
        p.pre_coset = 0; p.post_scale = inverse; p.post_coset = (inverse && coset);
        ZK_TRY(launch_pass(ctx, p, 1u << (log_ra - p.log_l), batch));
    }
    ZK_TRY(timed_end(ctx));
    {   // products issued: one per butterfly, plus the per-element 4-step twiddle, coset and n^-1 factors (x3 = thirds)
        const uint64_t elems = (uint64_t)batch << log_n;
        uint64_t thirds = (log_n > 11 ? 3 : 0) + (coset ? 2 : 0) + (inverse ? 3 : 0);
        ctx->ntt_products += elems / 2 * log_n + elems * thirds / 3;
    }
    return ZKFHE_OK;
}

}  // namespace zkfhe
