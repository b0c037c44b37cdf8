// C ABI of libzkfhe_b200 (include/zkfhe_b200.h): context lifecycle, host<->device
// staging and the on-device self test.  Compute lives in ntt.cu / msm.cu /
// witness.cu; nothing here falls back to the CPU.
#include <chrono>
#include <new>
#include <cstring>
#include "common.cuh"
#include "host_ff.h"

using namespace zkfhe;

namespace zkfhe {

// ---- self test ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int F> __device__ fe<F> random_fe(uint64_t& s) {
    fe<F> a;
    for (int i = 0; i < 8; i += 2) {
        uint64_t z = splitmix(s);
        a.v[i] = (uint32_t)z;
        a.v[i + 1] = (uint32_t)(z >> 32);
    }
    a.v[7] &= 0x1fffffffu;   // < 2^253 < modulus
    return a;
}

template <int F> __device__ uint32_t selftest_field(uint64_t& s) {
    uint32_t bad = 0;
    fe<F> a = random_fe<F>(s), b = random_fe<F>(s), c = random_fe<F>(s);
    if (!eq(mul(a, b), mul_c(a, b))) bad++;
    if (!eq(sqr(a), mul_c(a, a))) bad++;
    if (!eq(sub(add(a, b), b), a)) bad++;
    if (!eq(add(sub(a, b), b), a)) bad++;
    if (!eq(mul(a, add(b, c)), add(mul(a, b), mul(a, c)))) bad++;
    if (!eq(from_mont(to_mont(a)), a)) bad++;
    if (!eq(add(a, neg(a)), fe_zero<F>())) bad++;
    // edge operands
    fe<F> m1 = sub(fe_zero<F>(), fe_one<F>());
    if (!eq(mul(m1, m1), fe_one<F>())) bad++;
    if (!eq(mul(a, fe_one<F>()), a)) bad++;
    if (!is_zero(mul(a, fe_zero<F>()))) bad++;
    return bad;
}

__global__ void k_selftest(uint32_t n_cases, uint64_t seed, uint32_t* mismatches) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cases) return;
    uint64_t s = seed + 0x1234567ull * i;
    uint32_t bad = selftest_field<FR>(s) + selftest_field<FQ>(s);
    if (i < 64) {   // inversions and curve identities on a few threads only
        fr_t a = random_fe<FR>(s);
        if (!is_zero(a) && !eq(mul(a, inv(a)), fe_one<FR>())) bad++;
        if (!eq(inv(a), inv_fermat(a))) bad++;
        fq_t q = random_fe<FQ>(s);
        if (!is_zero(q) && !eq(mul(q, inv(q)), fe_one<FQ>())) bad++;
        if (!eq(inv(q), inv_fermat(q))) bad++;
        if (!is_zero(inv(fe_zero<FQ>())) || !eq(inv(fe_one<FQ>()), fe_one<FQ>())) bad++;
        // G = (1, 2) in Montgomery form; check 2G + G == 3G via two routes and the curve equation
        g1_affine g;
        g.x = fe_one<FQ>();
        g.y = add(fe_one<FQ>(), fe_one<FQ>());
        g1_xyzz d = xyzz_dbl_affine(g);
        g1_xyzz t1 = d;
        xyzz_madd(t1, g, false);                       // 2G + G
        g1_xyzz t2 = xyzz_from_affine(g);
        xyzz_add(t2, d);                               // G + 2G
        g1_affine a1 = xyzz_to_affine(t1), a2 = xyzz_to_affine(t2);
        if (!eq(a1.x, a2.x) || !eq(a1.y, a2.y)) bad++;
        fq_t three = add(add(fe_one<FQ>(), fe_one<FQ>()), fe_one<FQ>());
        if (!eq(sqr(a1.y), add(mul(sqr(a1.x), a1.x), three))) bad++;
        g1_xyzz z = t1;
        xyzz_madd(z, a1, true);                        // 3G - 3G = identity
        if (!is_identity(z)) bad++;
        g1_xyzz dd = t1;
        xyzz_madd(dd, a1, false);                      // 3G + 3G via the doubling branch
        g1_xyzz d2 = xyzz_dbl(t2);
        g1_affine b1 = xyzz_to_affine(dd), b2 = xyzz_to_affine(d2);
        if (!eq(b1.x, b2.x) || !eq(b1.y, b2.y)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

int selftest_run(zkfhe_ctx* ctx, uint32_t n_cases, uint64_t seed, uint32_t* mismatches) {
    uint32_t* d_bad;
    ZK_CUDA(ctx, cudaMalloc(&d_bad, 4));
    ZK_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    k_selftest<<<(n_cases + 127) / 128, 128, 0, ctx->stream>>>(n_cases, seed, d_bad);
    ZK_CHECK_LAUNCH(ctx);
    ZK_CUDA(ctx, cudaMemcpyAsync(mismatches, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    ZK_CUDA(ctx, cudaFree(d_bad));
    return ZKFHE_OK;
}

// ---- micro-benchmarks: the arithmetic ceilings the kernel rooflines are quoted against -----------
// kind 0: Montgomery products/s with every SM full (two independent chains per thread) -- the
//         IMAD-pipe peak that bounds MSM and NTT long before HBM does;
// kind 1..5: latency of a dependent chain on ONE warp: 1 XYZZ add, 2 mixed add, 3 field product,
//         4 inversion (binary Euclid), 5 inversion (Fermat).
__global__ void __launch_bounds__(256) k_bench_mul(fq_t* out, uint32_t iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fq_t a = fe_one<FQ>(), b = fconst<FQ>::r2(), c = fconst<FQ>::r3(), d = fconst<FQ>::r2();
    a.v[0] += t; c.v[1] ^= t;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) { a = mul(a, b); c = mul(c, d); }
    fe_store(out + t, add(a, c));
}
__global__ void __launch_bounds__(32) k_bench_chain(int kind, uint32_t iters, g1_xyzz* out) {
    g1_affine g;
    g.x = fe_one<FQ>();
    g.y = add(fe_one<FQ>(), fe_one<FQ>());
    g1_xyzz p = xyzz_from_affine(g), acc = xyzz_dbl_affine(g);
    acc.x.v[0] += 0;
    if (kind == 1) {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) xyzz_add(acc, p);
    } else if (kind == 2) {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) xyzz_madd(acc, g, false);
    } else if (kind == 3) {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) acc.x = mul(acc.x, acc.y);
    } else if (kind == 4) {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) acc.x = add(inv(acc.x), acc.y);
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) acc.x = add(inv_fermat(acc.x), acc.y);
    }
    if (threadIdx.x == 0) xyzz_store(out, acc);
}

// kind 6 / 7 -- an EXPERIMENT, not used by the prover: the instruction mix of a 256-bit Montgomery product on the FP64
// pipe (Emmart, Zheng, Weems, ARITH 2018): limbs of 52 bits held as doubles, each limb product split into two exact
// doubles by hi = fma_rz(a, b, 2^104), lo = fma_rz(a, b, 2^104 + 2^52 - hi), whose bit patterns are summed as 64-bit
// integers; five reduction steps, each one low product for m and five limb products m * n_j.  110 DFMA + ~70 DADD on the
// FP64 pipe and ~125 64-bit integer additions on the ALU pipe per product, no IMAD.  It answers one question (DESIGN.md
// section 9): is there a second multiply pipe worth feeding beside fmaheavy?  kind 6 runs it on every warp, kind 7 on the
// odd warps while the even warps run the IMAD product.  (The arithmetic is carried through faithfully enough that the
// compiler cannot drop it; the result is not checked against a field product.)
struct d5 { double v[5]; };
__device__ __forceinline__ double ll_to_d52(long long x) {       // integer < 2^52 -> double, no conversion instruction
    return __longlong_as_double((x & 0x000fffffffffffffLL) | 0x4330000000000000LL) - 0x1p52;
}
__device__ __forceinline__ d5 dfma_mont(const d5& a, const d5& b, const d5& n, double ninv) {
    const double C1 = 0x1p104, C3 = 0x1p104 + 0x1p52;
    long long acc[11];
#pragma unroll
    for (int k = 0; k < 11; k++) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double hi = __fma_rz(a.v[i], b.v[j], C1);
            const double lo = __fma_rz(a.v[i], b.v[j], C3 - hi);
            acc[i + j + 1] += __double_as_longlong(hi);
            acc[i + j] += __double_as_longlong(lo);
        }
#pragma unroll
    for (int k = 0; k < 11; k++) acc[k] -= (k < 5 ? k + 1 : 9 - k) * 0x4330000000000000LL + (k ? (k < 6 ? k : 10 - k) * 0x4670000000000000LL : 0);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const double t = ll_to_d52(acc[i]);
        const double mh = __fma_rz(t, ninv, C1);
        const double m = __fma_rz(t, ninv, C3 - mh) - 0x1p52;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double hi = __fma_rz(m, n.v[j], C1);
            const double lo = __fma_rz(m, n.v[j], C3 - hi);
            acc[i + j + 1] += __double_as_longlong(hi) - 0x4670000000000000LL;
            acc[i + j] += __double_as_longlong(lo) - 0x4330000000000000LL;
        }
        acc[i + 1] += acc[i] >> 52;
    }
    d5 r;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        r.v[k] = ll_to_d52(acc[5 + k]);
        if (k < 4) acc[6 + k] += acc[5 + k] >> 52;
    }
    return r;
}
__global__ void __launch_bounds__(256) k_bench_dfma(double* out, uint32_t iters, int mixed, fq_t* out_q) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (mixed && ((threadIdx.x >> 5) & 1) == 0) {                 // even warps: the IMAD product of k_bench_mul
        fq_t a = fe_one<FQ>(), b = fconst<FQ>::r2(), c = fconst<FQ>::r3(), d = fconst<FQ>::r2();
        a.v[0] += t; c.v[1] ^= t;
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) { a = mul(a, b); c = mul(c, d); }
        fe_store(out_q + t, add(a, c));
        return;
    }
    d5 a, b, c, d, n;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        a.v[k] = (double)(0x000f123456789abcLL >> k) + t;
        b.v[k] = (double)(0x0007fedcba987654LL >> k);
        c.v[k] = (double)(0x000a5a5a5a5a5a5aLL >> k) + 3 * t;
        d.v[k] = (double)(0x0003c3c3c3c3c3c3LL >> k);
        n.v[k] = (double)(0x000b85045b681815LL >> (2 * k));
    }
    const double ninv = (double)0x000c2e1f593efffffLL;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) { a = dfma_mont(a, b, n, ninv); c = dfma_mont(c, d, n, ninv); }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) s += a.v[k] + c.v[k];
    out[t] = s;
}

// kind 8 / 9 -- the second experiment of DESIGN.md section 9: point additions per second with every SM full, points
// gathered from the L2-resident window table of the loaded commitment key.
//   kind 8  what k_msm_accumulate does: one XYZZ accumulator per thread, mixed additions (8M + 2S = 10 products each);
//   kind 9  batched affine: each thread adds BA_B independent PAIRS of affine points with ONE inversion
//           (Montgomery's trick: prefix products of the x differences in local memory, binary-Euclid inversion off the
//           multiply pipe, 6 products per addition) -- the inner operation of a pairwise bucket tree.
// Neither result is checked here beyond being kept alive; the prover does not use kind 9.
static constexpr int BA_B = 16;
__global__ void __launch_bounds__(256) k_bench_madd(const g1_affine* __restrict__ table, uint32_t mask, uint32_t iters, g1_xyzz* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    g1_xyzz acc = xyzz_identity();
    uint32_t h = t * 2654435761u;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
        h = h * 1664525u + 1013904223u;
        xyzz_madd(acc, affine_load(table + (h & mask)), (h >> 31) != 0);
    }
    xyzz_store(out + t, acc);
}
__global__ void __launch_bounds__(256) k_bench_batched_affine(const g1_affine* __restrict__ table, uint32_t mask, uint32_t iters, fq_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fq_t check = fe_zero<FQ>();
    uint32_t h = t * 2654435761u;
    fq_t d[BA_B], pre[BA_B];
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it += BA_B) {
        const uint32_t h0 = h;
        fq_t run = fe_one<FQ>();
#pragma unroll 1
        for (int i = 0; i < BA_B; i++) {                          // forward: prefix products of (x2 - x1)
            h = h * 1664525u + 1013904223u;
            const uint32_t ip = h & mask, iq = (ip + 1 + ((h >> 20) & 63)) & mask;      // two distinct table entries
            d[i] = sub(fe_load_nc(&table[iq].x), fe_load_nc(&table[ip].x));
            pre[i] = run;
            run = mul(run, d[i]);
        }
        fq_t inv_run = inv(run);                                   // one inversion for the whole batch
        h = h0;
        uint32_t idx[BA_B];
#pragma unroll 1
        for (int i = 0; i < BA_B; i++) { h = h * 1664525u + 1013904223u; idx[i] = h; }
#pragma unroll 1
        for (int i = BA_B - 1; i >= 0; i--) {                      // backward: 1 / d_i, lambda, the sum
            const uint32_t ip = idx[i] & mask, iq = (ip + 1 + ((idx[i] >> 20) & 63)) & mask;
            const fq_t dinv = mul(inv_run, pre[i]);
            inv_run = mul(inv_run, d[i]);
            const g1_affine P = affine_load(table + ip), Q = affine_load(table + iq);
            const fq_t lam = mul(sub(Q.y, P.y), dinv);
            const fq_t x3 = sub(sub(sqr(lam), P.x), Q.x);
            const fq_t y3 = sub(mul(lam, sub(P.x, x3)), P.y);
            check = add(check, add(x3, y3));
        }
    }
    fe_store(out + t, check);
}

int microbench_run(zkfhe_ctx* ctx, int kind, uint32_t iters, float* ms, uint64_t* ops) {
    if (kind < 0 || kind > 9 || !iters) return fail(ctx, ZKFHE_ERR_ARG, "microbench: kind in [0,9], iters > 0");
    if (kind >= 8 && !ctx->basis[0].loaded) return fail(ctx, ZKFHE_ERR_STATE, "microbench kinds 8, 9 read the window table: load an SRS first");
    if (kind == 9) iters = (iters + BA_B - 1) / BA_B * BA_B;
    cudaDeviceProp prop;
    ZK_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const uint32_t blocks = (uint32_t)prop.multiProcessorCount * 8, threads = 256;
    void* buf;
    ZK_TRY(ws_get(ctx, "microbench", (size_t)blocks * threads * (kind == 8 ? sizeof(g1_xyzz) : sizeof(fq_t) + sizeof(double)) + sizeof(g1_xyzz), &buf));
    // gather range of kinds 8 / 9: the first 16 window rows of the table (W >= 16 for every window width in use), a power of two
    const uint32_t tmask = kind >= 8 ? (1u << (ctx->basis[0].log_n + (ctx->basis[0].W >= 16 ? 4 : 0))) - 1 : 0;
    cudaEvent_t e0, e1;
    ZK_CUDA(ctx, cudaEventCreate(&e0));
    ZK_CUDA(ctx, cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {          // first pass warms up; the second is the one reported
        ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        if (kind == 0) k_bench_mul<<<blocks, threads, 0, ctx->stream>>>((fq_t*)buf, iters);
        else if (kind == 8) k_bench_madd<<<blocks, threads, 0, ctx->stream>>>(ctx->basis[0].table, tmask, iters, (g1_xyzz*)buf);
        else if (kind == 9) k_bench_batched_affine<<<blocks, threads, 0, ctx->stream>>>(ctx->basis[0].table, tmask, iters, (fq_t*)buf);
        else if (kind >= 6) k_bench_dfma<<<blocks, threads, 0, ctx->stream>>>((double*)((fq_t*)buf + (size_t)blocks * threads), iters, kind == 7, (fq_t*)buf);
        else k_bench_chain<<<1, 32, 0, ctx->stream>>>(kind, iters, (g1_xyzz*)buf);
        ZK_CHECK_LAUNCH(ctx);
        ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        ZK_CUDA(ctx, cudaEventSynchronize(e1));
    }
    ZK_CUDA(ctx, cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops = kind >= 8 ? (uint64_t)blocks * threads * iters : kind == 0 || kind >= 6 ? (uint64_t)blocks * threads * iters * 2 : (uint64_t)iters;
    return ZKFHE_OK;
}

}  // namespace zkfhe

// ---- lifecycle ---------------------------------------------------------------------------
extern "C" {

const char* zkfhe_version(void) { return "zkfhe_b200 0.2 (sm_100a)"; }

// ---- host-only hooks: the Fiat-Shamir transcript, testable without a GPU ----------------------------------------
int zkfhe_poseidon_permute(uint8_t* state160, int plain) {
    if (!state160) return ZKFHE_ERR_ARG;
    host::Fr s[POSEIDON_T];
    memcpy(s, state160, sizeof s);
    for (auto& v : s)
        if (host::geq(v, host::FR_MOD)) return ZKFHE_ERR_ARG;
    if (plain == 0) host::poseidon_permute(s);
    else if (plain == 1) host::poseidon_permute_plain(s);
    else if (plain == 2) host::poseidon_permute_scalar(s);
    else if (plain == 3) {
        if (!host::poseidon_ifma_available()) return ZKFHE_ERR_STATE;
        host::poseidon_permute_ifma(s);
    } else return ZKFHE_ERR_ARG;
    memcpy(state160, s, sizeof s);
    return ZKFHE_OK;
}

// The trapdoor of the reference's fallback SRS (host_ff.h reference_test_tau) and the ChaCha20 block it comes from.
int zkfhe_reference_test_tau(uint8_t* tau_fr32, uint8_t* keystream64) {
    if (!tau_fr32) return ZKFHE_ERR_ARG;
    const host::Fr t = host::reference_test_tau();
    memcpy(tau_fr32, t.l, 32);
    if (keystream64) {
        const uint32_t zero_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        host::chacha20_block(zero_key, 0, 0, keystream64);
    }
    return ZKFHE_OK;
}

// ns per operation on the calling host thread: kind 0 = Poseidon permutation (the form the transcript runs), 1 =
// dependent Fr products, 2 = plain-form permutation, 3 = scalar optimised form, 4 = the AVX-512 IFMA form
// (poseidon_ifma.cpp).  `features` (optional, >= 64 bytes) says which code path the host runs.
int zkfhe_host_microbench(int kind, uint32_t iters, double* ns_per_op, char* features, size_t cap) {
    if (!ns_per_op || !iters || kind < 0 || kind > 4) return ZKFHE_ERR_ARG;
    if (kind >= 4 && !host::poseidon_ifma_available()) return ZKFHE_ERR_STATE;
    host::Fr s[POSEIDON_T];
    for (int i = 0; i < POSEIDON_T; i++) s[i] = host::from_u64(i + 1);
    const auto t0 = std::chrono::steady_clock::now();
    if (kind == 0) for (uint32_t i = 0; i < iters; i++) host::poseidon_permute(s);
    else if (kind == 2) for (uint32_t i = 0; i < iters; i++) host::poseidon_permute_plain(s);
    else if (kind == 3) for (uint32_t i = 0; i < iters; i++) host::poseidon_permute_scalar(s);
    else if (kind == 4) for (uint32_t i = 0; i < iters; i++) host::poseidon_permute_ifma(s);
    else for (uint32_t i = 0; i < iters; i++) s[0] = host::mul(s[0], s[1]);
    *ns_per_op = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / iters;
    if (s[0].is_zero() && s[1].is_zero()) *ns_per_op = -1;       // keeps the loop observable
    if (features && cap) {
#if defined(__x86_64__) && defined(__GNUC__)
        snprintf(features, cap, "%s%s", host::cpu_has_adx() ? "fr_mul: mulx/adcx/adox" : "fr_mul: portable (no BMI2/ADX)",
                 host::poseidon_ifma_available() ? "; poseidon: avx512ifma" : "; poseidon: scalar");
#else
        snprintf(features, cap, "fr_mul: portable");
#endif
    }
    return ZKFHE_OK;
}

int zkfhe_transcript_replay(int kind, const uint8_t* script, size_t len, uint8_t* out, size_t cap, size_t* n_challenges) {
    if (!script || !n_challenges || (kind != host::TRANSCRIPT_BLAKE2B && kind != host::TRANSCRIPT_POSEIDON)) return ZKFHE_ERR_ARG;
    host::Transcript tr(kind);
    size_t pos = 0, count = 0;
    while (pos < len) {
        const uint8_t op = script[pos++];
        if (op == 1) {                                   // scalar
            if (pos + 32 > len) return ZKFHE_ERR_ARG;
            host::Fr c;
            memcpy(c.l, script + pos, 32);
            pos += 32;
            if (host::geq(c, host::FR_MOD)) return ZKFHE_ERR_ARG;
            tr.common_scalar(host::to_mont(c));
        } else if (op == 2) {                            // point
            if (pos + 64 > len) return ZKFHE_ERR_ARG;
            uint64_t xy[8];
            memcpy(xy, script + pos, 64);
            pos += 64;
            tr.common_point(xy, xy + 4);
        } else if (op == 3) {                            // squeeze
            const host::Fr c = host::from_mont(tr.squeeze());
            if (out) {
                if ((count + 1) * 32 > cap) return ZKFHE_ERR_ARG;
                memcpy(out + count * 32, c.l, 32);
            }
            count++;
        } else {
            return ZKFHE_ERR_ARG;
        }
    }
    *n_challenges = count;
    return ZKFHE_OK;
}

int zkfhe_init(int device, zkfhe_ctx** out) {
    if (!out) return ZKFHE_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return ZKFHE_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ZKFHE_ERR_CUDA;
    zkfhe_ctx* ctx = new (std::nothrow) zkfhe_ctx();
    if (!ctx) return ZKFHE_ERR_CUDA;
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return ZKFHE_ERR_CUDA;
    }
    ctx->own_stream = true;
    {   // Poly handles come from the stream-ordered allocator (~20 per proof).  By default the pool hands freed memory back
        // to the driver at every synchronisation and maps it again on the next allocation; keep it (the polynomials of
        // the next proof have the same sizes).  Process-wide for this device, idempotent.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    *out = ctx;
    return ZKFHE_OK;
}

void zkfhe_destroy(zkfhe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    zkfhe_comm_destroy(ctx);
    for (auto& kv : ctx->domains) { cudaFree(kv.second.tw_fwd); cudaFree(kv.second.tw_inv); cudaFree(kv.second.tw_inv_s); }
    for (auto& b : ctx->basis) if (b.table && !b.shared) { cudaFree(b.table); if (b.table_s) cudaFree(b.table_s); }
    for (auto& kv : ctx->ws) if (kv.second.p) cudaFree(kv.second.p);
    for (auto& pr : ctx->ev_pairs) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* zkfhe_last_error(const zkfhe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int zkfhe_set_stream(zkfhe_ctx* ctx, void* cuda_stream) {
    if (!ctx) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;     // NULL is the CUDA legacy default stream
    ctx->own_stream = false;
    return ZKFHE_OK;
}

int zkfhe_set_blocking_sync(zkfhe_ctx* ctx, int on) {
    if (!ctx) return ZKFHE_ERR_ARG;
    ctx->blocking_sync = on != 0;
    return ZKFHE_OK;
}

int zkfhe_sync(zkfhe_ctx* ctx) {
    if (!ctx) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    ctx->h_stage_used = 0;                   // the stream is empty: every staged upload has been consumed
    return ZKFHE_OK;
}

uint64_t zkfhe_launch_count(const zkfhe_ctx* ctx) { return ctx ? ctx->launches : 0; }

float zkfhe_last_kernel_ms(const zkfhe_ctx* ctx) {
    if (!ctx) return -1.f;
    // dominant kernel of the last call: the accumulate kernel for MSM, the butterfly passes for NTT
    float total = 0.f;
    for (size_t i = ctx->call_mark; i < ctx->ev_used; i++) {
        if (ctx->ev_info[i].cat >= ZK_CAT_MSM_OTHER) continue;
        if (cudaEventSynchronize(ctx->ev_pairs[i].second) != cudaSuccess) return -1.f;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pairs[i].first, ctx->ev_pairs[i].second) != cudaSuccess) return -1.f;
        total += ms;
    }
    return total;
}

int zkfhe_timing_reset(zkfhe_ctx* ctx) {
    if (!ctx) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    ctx->ev_used = ctx->call_mark = 0;
    ctx->ntt_products = 0;
    auto it = ctx->ws.find("msm_refs");
    if (it != ctx->ws.end()) ZK_CUDA(ctx, cudaMemset(it->second.p, 0, 8));
    return ZKFHE_OK;
}

int zkfhe_timing_get(zkfhe_ctx* ctx, int category, float* ms, uint32_t* spans, uint64_t* units) {
    if (!ctx || category < 0 || category >= ZK_CAT_COUNT) return ZKFHE_ERR_ARG;
    if (category == ZK_CAT_MSM_REFS) {
        unsigned long long v = 0;
        auto it = ctx->ws.find("msm_refs");
        if (it != ctx->ws.end()) {
            ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
            ZK_CUDA(ctx, cudaMemcpy(&v, it->second.p, 8, cudaMemcpyDeviceToHost));
        }
        if (ms) *ms = 0.f;
        if (spans) *spans = 0;
        if (units) *units = v;
        return ZKFHE_OK;
    }
    if (category == ZK_CAT_NTT_PRODUCTS) {
        if (ms) *ms = 0.f;
        if (spans) *spans = 0;
        if (units) *units = ctx->ntt_products;
        return ZKFHE_OK;
    }
    float total = 0.f;
    uint32_t cnt = 0;
    uint64_t un = 0;
    for (size_t i = 0; i < ctx->ev_used; i++) {
        if (ctx->ev_info[i].cat != category) continue;
        ZK_CUDA(ctx, cudaEventSynchronize(ctx->ev_pairs[i].second));
        float t = 0.f;
        ZK_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev_pairs[i].first, ctx->ev_pairs[i].second));
        total += t;
        cnt++;
        un += ctx->ev_info[i].units;
    }
    if (ms) *ms = total;
    if (spans) *spans = cnt;
    if (units) *units = un;
    return ZKFHE_OK;
}

int zkfhe_microbench(zkfhe_ctx* ctx, int kind, uint32_t iters, float* ms, uint64_t* ops) {
    if (!ctx || !ms || !ops) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return microbench_run(ctx, kind, iters, ms, ops);
}

int zkfhe_selftest(zkfhe_ctx* ctx, uint32_t n_cases, uint64_t seed, uint32_t* mismatches) {
    if (!ctx || !mismatches) return ZKFHE_ERR_ARG;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return selftest_run(ctx, n_cases, seed, mismatches);
}

// ---- NTT ---------------------------------------------------------------------------------
int zkfhe_ntt_fr_dev(zkfhe_ctx* ctx, uint8_t* d_data, uint32_t log_n, uint32_t batch, int inverse, int coset) {
    if (!ctx || !d_data) return fail(ctx, ZKFHE_ERR_ARG, "ntt: null pointer");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    fr_t* d = reinterpret_cast<fr_t*>(d_data);
    uint64_t n = 1ull << log_n;
    return ntt_run(ctx, d, n, (uint32_t)n, d, n, log_n, batch, inverse, coset);
}

int zkfhe_ntt_fr(zkfhe_ctx* ctx, uint8_t* h_data, uint32_t log_n, uint32_t batch, int inverse, int coset) {
    if (!ctx || !h_data) return fail(ctx, ZKFHE_ERR_ARG, "ntt: null pointer");
    if (log_n < 1 || log_n > 22) return fail(ctx, ZKFHE_ERR_ARG, "ntt: log_n=%u out of range [1,22]", log_n);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = ((size_t)batch << log_n) * sizeof(fr_t);
    if (bytes == 0) return ZKFHE_OK;
    void* d;
    ZK_TRY(ws_get(ctx, "ntt_io", bytes, &d));
    ZK_CUDA(ctx, cudaMemcpyAsync(d, h_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(zkfhe_ntt_fr_dev(ctx, (uint8_t*)d, log_n, batch, inverse, coset));
    ZK_CUDA(ctx, cudaMemcpyAsync(h_data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_coeff_to_extended_dev(zkfhe_ctx* ctx, const uint8_t* d_coeffs, uint32_t log_n_in, uint8_t* d_ext,
                                uint32_t log_n_out, uint32_t batch) {
    if (!ctx || !d_coeffs || !d_ext) return fail(ctx, ZKFHE_ERR_ARG, "coeff_to_extended: null pointer");
    if (log_n_in > log_n_out) return fail(ctx, ZKFHE_ERR_ARG, "coeff_to_extended: log_n_in > log_n_out");
    if (log_n_out < 12 && (const void*)d_coeffs == (const void*)d_ext && log_n_in != log_n_out)
        return fail(ctx, ZKFHE_ERR_ARG, "coeff_to_extended: in-place needs equal sizes");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, reinterpret_cast<const fr_t*>(d_coeffs), 1ull << log_n_in, 1u << log_n_in,
                   reinterpret_cast<fr_t*>(d_ext), 1ull << log_n_out, log_n_out, batch, 0, 1);
}

// ---- MSM ---------------------------------------------------------------------------------
int zkfhe_load_srs(zkfhe_ctx* ctx, uint32_t k, const uint8_t* h_g, const uint8_t* h_g_lagrange) {
    if (!ctx) return ZKFHE_ERR_ARG;
    if (k < 1 || k > 22) return fail(ctx, ZKFHE_ERR_ARG, "load_srs: k=%u out of range [1,22]", k);
    if (!h_g && !h_g_lagrange) return fail(ctx, ZKFHE_ERR_ARG, "load_srs: both bases are null");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = ((size_t)1 << k) * sizeof(g1_affine);
    void* d;
    ZK_TRY(ws_get(ctx, "srs_stage", bytes, &d));
    const uint8_t* src[2] = {h_g, h_g_lagrange};
    for (int which = 0; which < 2; which++) {
        if (!src[which]) continue;
        ZK_CUDA(ctx, cudaMemcpyAsync(d, src[which], bytes, cudaMemcpyHostToDevice, ctx->stream));
        ZK_TRY(msm_load_basis(ctx, which, (const g1_affine*)d, k, which == 1));
    }
    ctx->srs_k = k;
    return ZKFHE_OK;
}

int zkfhe_srs_setup(zkfhe_ctx* ctx, uint32_t k, const uint8_t* h_tau_fr, uint8_t* h_g_out, uint8_t* h_g_lagrange_out) {
    if (!ctx || !h_tau_fr) return fail(ctx, ZKFHE_ERR_ARG, "srs_setup: null pointer");
    if (k < 1 || k > 22) return fail(ctx, ZKFHE_ERR_ARG, "srs_setup: k=%u out of range [1,22]", k);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = ((size_t)1 << k) * sizeof(g1_affine);
    g1_affine* d;
    ZK_TRY(ws_get(ctx, "srs_setup", 2 * bytes, (void**)&d));
    fr_t tau;
    memcpy(&tau, h_tau_fr, 32);
    ZK_TRY(srs_setup(ctx, k, tau, d, d + ((size_t)1 << k)));
    ZK_TRY(msm_load_basis(ctx, 0, d, k, false));
    ZK_TRY(msm_load_basis(ctx, 1, d + ((size_t)1 << k), k, true));
    ctx->srs_k = k;
    if (h_g_out) ZK_CUDA(ctx, cudaMemcpyAsync(h_g_out, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_g_lagrange_out)
        ZK_CUDA(ctx, cudaMemcpyAsync(h_g_lagrange_out, d + ((size_t)1 << k), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_share_srs(zkfhe_ctx* dst, const zkfhe_ctx* src) {
    if (!dst || !src) return ZKFHE_ERR_ARG;
    if (dst->device != src->device) return fail(dst, ZKFHE_ERR_ARG, "share_srs: contexts are on different devices");
    if (src->srs_k == 0) return fail(dst, ZKFHE_ERR_STATE, "share_srs: the source context has no SRS");
    for (int i = 0; i < 2; i++) {
        if (dst->basis[i].table && !dst->basis[i].shared) {
            cudaFree(dst->basis[i].table);
            if (dst->basis[i].table_s) cudaFree(dst->basis[i].table_s);
        }
        dst->basis[i] = src->basis[i];
        dst->basis[i].shared = true;
    }
    dst->srs_k = src->srs_k;
    return ZKFHE_OK;
}

int zkfhe_fr_convert_dev(zkfhe_ctx* ctx, uint8_t* d_data, uint64_t count, int to_montgomery) {
    if (!ctx || (!d_data && count)) return fail(ctx, ZKFHE_ERR_ARG, "fr_convert: null pointer");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return fr_convert(ctx, reinterpret_cast<fr_t*>(d_data), count, to_montgomery);
}

int zkfhe_msm_g1_dev(zkfhe_ctx* ctx, const uint8_t* d_scalars, uint32_t batch, int basis, uint8_t* d_out_affine) {
    if (!ctx || !d_scalars || !d_out_affine) return fail(ctx, ZKFHE_ERR_ARG, "msm: null pointer");
    if (ctx->srs_k == 0) return fail(ctx, ZKFHE_ERR_STATE, "msm: zkfhe_load_srs has not been called");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return msm_run(ctx, reinterpret_cast<const fr_t*>(d_scalars), 1ull << ctx->srs_k, ctx->srs_k, batch, basis,
                   reinterpret_cast<g1_affine*>(d_out_affine));
}

int zkfhe_msm_g1_dev_ex(zkfhe_ctx* ctx, const uint8_t* d_scalars, uint32_t batch, int basis, int small_values,
                        uint8_t* d_out_affine) {
    if (!ctx || !d_scalars || !d_out_affine) return fail(ctx, ZKFHE_ERR_ARG, "msm: null pointer");
    if (ctx->srs_k == 0) return fail(ctx, ZKFHE_ERR_STATE, "msm: zkfhe_load_srs has not been called");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    return msm_run(ctx, reinterpret_cast<const fr_t*>(d_scalars), 1ull << ctx->srs_k, ctx->srs_k, batch, basis,
                   reinterpret_cast<g1_affine*>(d_out_affine), small_values);
}

int zkfhe_msm_g1(zkfhe_ctx* ctx, const uint8_t* h_scalars, uint32_t batch, int basis, uint8_t* h_out_affine) {
    if (!ctx || !h_scalars || !h_out_affine) return fail(ctx, ZKFHE_ERR_ARG, "msm: null pointer");
    if (ctx->srs_k == 0) return fail(ctx, ZKFHE_ERR_STATE, "msm: zkfhe_load_srs has not been called");
    if (batch == 0) return ZKFHE_OK;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t sbytes = ((size_t)batch << ctx->srs_k) * sizeof(fr_t), obytes = (size_t)batch * sizeof(g1_affine);
    void *d_s, *d_o;
    ZK_TRY(ws_get(ctx, "msm_in", sbytes, &d_s));
    ZK_TRY(ws_get(ctx, "msm_out", obytes, &d_o));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_s, h_scalars, sbytes, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(zkfhe_msm_g1_dev(ctx, (const uint8_t*)d_s, batch, basis, (uint8_t*)d_o));
    ZK_CUDA(ctx, cudaMemcpyAsync(h_out_affine, d_o, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

}  // extern "C"
