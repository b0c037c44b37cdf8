// Shared host-side plumbing for libzkfhe_b200: the context object behind the C
// ABI (include/zkfhe_b200.h), error reporting and device workspace management.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ec.cuh"
#include "../../include/zkfhe_b200.h"

namespace zkfhe {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// Pippenger layout for one SRS basis (resident across proofs).
struct MsmBasis {
    g1_affine* table = nullptr;   // [W][n] : table[w*n + i] = 2^(c*w) * P_i  (affine, Montgomery)
    uint32_t log_n = 0, c = 0, W = 0;
    // second expansion with a narrow window for columns of small values (witness cells, lookup inputs):
    // a 29-bit value has 3 signed digits at c = 10 or at c = 13 alike, but c = 10 has 512 buckets to
    // reduce per column instead of 4096, and the bucket reduction is what such columns pay for
    g1_affine* table_s = nullptr;
    uint32_t c_s = 0, W_s = 0;
    // prefix sums behind each table (Lagrange basis only): entry W*n + w*(n+1) + i of the same allocation holds
    // sum_{j<i} table[w*n + j], so a run of equal scalars over rows [a, b] costs two references per window
    // (+prefix[b+1], -prefix[a]) instead of b - a + 1.  Grand-product columns are constant over thousands of rows
    // wherever a permutation chunk has no copy-constrained cell, and 0/1 witness columns are full of runs.
    bool prefix = false;
    bool loaded = false;
    bool shared = false;          // table owned by another context (zkfhe_share_srs)
};

struct NttDomain {
    fr_t* tw_fwd = nullptr;   // omega^i,  i < n
    fr_t* tw_inv = nullptr;   // omega^-i, i < n
    fr_t* tw_inv_s = nullptr; // omega^-i / n, i < n (4-step twiddles of the inverse transform)
    fr_t n_inv;               // host copy (Montgomery), passed by value to kernels
};

}  // namespace zkfhe

struct zkfhe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;                       // kernels launched through this ctx
    std::map<uint32_t, zkfhe::NttDomain> domains;   // by log_n
    zkfhe::MsmBasis basis[2];                    // 0: coefficient basis g, 1: Lagrange basis
    uint32_t srs_k = 0;
    std::map<std::string, zkfhe::DevBuf> ws;     // named, grow-only workspaces
    // CUDA-event pairs bracketing the dominant kernel(s) of the last NTT / MSM call
    // (a pool: spans accumulate until zkfhe_timing_reset; `call_mark` is where the last call started)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs;
    struct SpanInfo { int cat; uint64_t units; };
    std::vector<SpanInfo> ev_info;
    size_t ev_used = 0, call_mark = 0;
    uint64_t ntt_products = 0;                   // butterflies + twiddle / coset / scaling products since timing_reset
    // one proof sharded over several GPUs (comm.cu): rank / size of the NCCL communicator bound to this context, or
    // -- for tests on one GPU -- `virtual_ranks` shards computed one after the other by this context
    void* nccl_comm = nullptr;
    int rank = 0, n_ranks = 1;
    int virtual_ranks = 0;
    float comm_ms = 0;                           // device time spent in collectives since timing_reset (category 7)
    uint32_t comm_calls = 0;
    // host waits: by default the calling thread SLEEPS on a blocking-sync event until the stream has drained (many
    // proofs in flight share the host cores with the transcript's Poseidon sponge; cudaStreamSynchronize would spin a
    // core per waiting thread); zkfhe_set_blocking_sync(ctx, 0) switches back to the spinning wait for lowest latency
    bool blocking_sync = true;
    cudaEvent_t sync_event = nullptr;
    // page-locked landing buffer for the small device -> host read-backs of a proof (commitments, evaluations, status):
    // a copy into PAGEABLE memory makes cudaMemcpyAsync itself wait -- spinning -- for everything queued before it
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    // ... and a page-locked arena the small host -> device uploads are staged in (a copy FROM pageable memory drains the
    // stream before it starts); bump-allocated, rewound whenever the stream is known to be empty
    uint8_t* h_stage = nullptr;
    size_t h_stage_used = 0;
};
enum { ZK_CAT_MSM_ACCUMULATE = 0, ZK_CAT_NTT = 1, ZK_CAT_MSM_OTHER = 2, ZK_CAT_MSM_FOLD = 3, ZK_CAT_MSM_FINAL = 4,
       ZK_CAT_MSM_REFS = 5 /* no time: units = point additions issued by the accumulate kernel */, 
       ZK_CAT_NTT_PRODUCTS = 6 /* no time: units = field products issued by the NTT passes */,
       ZK_CAT_COMM = 7 /* NCCL collectives of a sharded proof (units = bytes gathered) */, ZK_CAT_COUNT = 8 };

namespace zkfhe {

inline cudaError_t stream_wait(zkfhe_ctx* ctx) {
    if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->stream);
    if (!ctx->sync_event) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->sync_event);
}

inline int fail(zkfhe_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define ZK_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return zkfhe::fail(ctx, ZKFHE_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,      \
                               cudaGetErrorString(e__));                                            \
    } while (0)

#define ZK_CHECK_LAUNCH(ctx)                                                                        \
    do {                                                                                            \
        (ctx)->launches++;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return zkfhe::fail(ctx, ZKFHE_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,         \
                               cudaGetErrorString(e__));                                            \
    } while (0)

#define ZK_TRY(expr)                  \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != ZKFHE_OK) return rc__; \
    } while (0)

inline void timed_call_start(zkfhe_ctx* ctx) {
    if (ctx->ev_used > 8192) ctx->ev_used = 0;      // bounded pool: old spans are dropped
    ctx->call_mark = ctx->ev_used;
}
inline int timed_begin(zkfhe_ctx* ctx, int cat = 0, uint64_t units = 0) {
    if (ctx->ev_used == ctx->ev_pairs.size()) {
        cudaEvent_t a, b;
        ZK_CUDA(ctx, cudaEventCreate(&a));
        ZK_CUDA(ctx, cudaEventCreate(&b));
        ctx->ev_pairs.emplace_back(a, b);
        ctx->ev_info.push_back({cat, units});
    }
    ctx->ev_info[ctx->ev_used] = {cat, units};
    ZK_CUDA(ctx, cudaEventRecord(ctx->ev_pairs[ctx->ev_used].first, ctx->stream));
    return ZKFHE_OK;
}
inline int timed_end(zkfhe_ctx* ctx) {
    ZK_CUDA(ctx, cudaEventRecord(ctx->ev_pairs[ctx->ev_used].second, ctx->stream));
    ctx->ev_used++;
    return ZKFHE_OK;
}

// Page-locked host scratch of at least `bytes` (grow-only; contents are only valid until the next call).
inline int pinned_get(zkfhe_ctx* ctx, size_t bytes, void** out) {
    if (ctx->h_pinned_bytes < bytes) {
        if (ctx->h_pinned) {
            ZK_CUDA(ctx, stream_wait(ctx));
            ZK_CUDA(ctx, cudaFreeHost(ctx->h_pinned));
            ctx->h_pinned = nullptr;
            ctx->h_pinned_bytes = 0;
        }
        const size_t want = bytes < (1u << 16) ? (1u << 16) : bytes + bytes / 4;
        ZK_CUDA(ctx, cudaHostAlloc(&ctx->h_pinned, want, cudaHostAllocDefault));
        ctx->h_pinned_bytes = want;
    }
    *out = ctx->h_pinned;
    return ZKFHE_OK;
}
// device -> host through the page-locked scratch, then one (sleeping) wait; `h_dst` may be pageable
inline int read_back(zkfhe_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    void* pin;
    ZK_TRY(pinned_get(ctx, bytes, &pin));
    ZK_CUDA(ctx, cudaMemcpyAsync(pin, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, stream_wait(ctx));
    ctx->h_stage_used = 0;                      // the stream is empty: every staged upload has been consumed
    memcpy(h_dst, pin, bytes);
    return ZKFHE_OK;
}
// host -> device, asynchronous, from pageable memory: staged through the page-locked arena (uploads larger than a
// quarter of the arena go straight through cudaMemcpyAsync, which then synchronises by itself)
static constexpr size_t H_STAGE_BYTES = 4u << 20;
inline int upload_async(zkfhe_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (!bytes) return ZKFHE_OK;
    if (bytes > H_STAGE_BYTES / 4) {
        ZK_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return ZKFHE_OK;
    }
    if (!ctx->h_stage) ZK_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_stage, H_STAGE_BYTES, cudaHostAllocDefault));
    const size_t need = (bytes + 63) & ~(size_t)63;
    if (ctx->h_stage_used + need > H_STAGE_BYTES) {
        ZK_CUDA(ctx, stream_wait(ctx));
        ctx->h_stage_used = 0;
    }
    uint8_t* slot = ctx->h_stage + ctx->h_stage_used;
    ctx->h_stage_used += need;
    memcpy(slot, h_src, bytes);
    ZK_CUDA(ctx, cudaMemcpyAsync(d_dst, slot, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return ZKFHE_OK;
}

// Grow-only named workspace.  Reallocation synchronises the stream first.
inline int ws_get(zkfhe_ctx* ctx, const char* name, size_t bytes, void** out) {
    DevBuf& b = ctx->ws[name];
    if (b.bytes < bytes) {
        if (b.p) {
            ZK_CUDA(ctx, stream_wait(ctx));
            ZK_CUDA(ctx, cudaFree(b.p));
            b.p = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8;
        ZK_CUDA(ctx, cudaMalloc(&b.p, want));
        b.bytes = want;
    }
    *out = b.p;
    return ZKFHE_OK;
}

// ---- entry points implemented in the .cu files (device-pointer level) ----------
int ntt_domain(zkfhe_ctx* ctx, uint32_t log_n, NttDomain** out);
int ntt_run(zkfhe_ctx* ctx, const fr_t* d_in, uint64_t in_stride, uint32_t in_len, fr_t* d_out,
            uint64_t out_stride, uint32_t log_n, uint32_t batch, int inverse, int coset);
int msm_load_basis(zkfhe_ctx* ctx, int which, const g1_affine* d_bases, uint32_t log_n, bool prefix);
int msm_run(zkfhe_ctx* ctx, const fr_t* d_scalars, uint64_t stride, uint32_t log_n, uint32_t batch,
            int which, g1_affine* d_out, int small_values = 0);
int srs_setup(zkfhe_ctx* ctx, uint32_t log_n, const fr_t& tau_mont, g1_affine* d_g, g1_affine* d_gl);
int fr_convert(zkfhe_ctx* ctx, fr_t* d, uint64_t count, int to_montgomery);
int points_to_canonical(zkfhe_ctx* ctx, g1_affine* d_pts, uint32_t count);
int selftest_run(zkfhe_ctx* ctx, uint32_t n_cases, uint64_t seed, uint32_t* mismatches);

// ---- sharding of one proof over ranks (comm.cu) ----------------------------------------------------------------
// G shards; this process computes shards [first, last): its own rank under NCCL, all of them in virtual mode, the
// single shard 0 otherwise.
struct Shards { uint32_t G, first, last; };
inline Shards shards_of(const zkfhe_ctx* ctx) {
    if (ctx->nccl_comm) return Shards{(uint32_t)ctx->n_ranks, (uint32_t)ctx->rank, (uint32_t)ctx->rank + 1};
    if (ctx->virtual_ranks > 1) return Shards{(uint32_t)ctx->virtual_ranks, 0, (uint32_t)ctx->virtual_ranks};
    return Shards{1, 0, 1};
}
// contiguous block of shard v when `count` items are split over G shards: per = ceil(count / G) items per shard
inline uint32_t shard_per(uint32_t count, uint32_t G) { return (count + G - 1) / G; }
inline void shard_range(uint32_t count, uint32_t G, uint32_t v, uint32_t* lo, uint32_t* hi) {
    const uint32_t per = shard_per(count, G);
    *lo = v * per < count ? v * per : count;
    *hi = *lo + per < count ? *lo + per : count;
}
// in place: every rank has written its block [rank * bytes_per_rank, (rank + 1) * bytes_per_rank) of `d_buf`; afterwards
// every rank holds all blocks.  No-op without a communicator (virtual shards were written by this process).
int comm_allgather(zkfhe_ctx* ctx, void* d_buf, size_t bytes_per_rank);

}  // namespace zkfhe
