// One proof over several GPUs: the NCCL communicator behind the C ABI's zkfhe_comm_* calls and the single collective
// the sharded prover uses (an in-place all-gather on the context's stream).
//
// The reference has no distributed code at all; this is the B200 side of SURVEY.md section 8(e):
// commitment phases shard by column, the quotient by expression (each rank transforms and evaluates only the columns
// its block of gates / permutation chunks / lookups reads), the openings by column, and what crosses NVLink is the
// 64-byte commitments of a phase, one 4n x 32-byte share of the quotient numerator, ~900 evaluations and six n x 32-byte
// opening sums per rank.  Curve points cannot be summed by
// ncclAllReduce, and gathering (never reducing) keeps every rank's transcript -- hence the proof -- byte-identical to
// the single-GPU one.
//
// libnccl.so.2 is loaded at run time with dlopen: inside a torch.distributed process that is the copy torch already
// loaded, otherwise the system library.  Only the C API subset below is used (stable since NCCL 2.0).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    const char* error = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = "libnccl.so.2 not found (dlopen)"; return; }
        api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
            api.error = "libnccl.so.2 lacks a required symbol";
    });
    return &api;
}

}  // namespace

namespace zkfhe {

int comm_allgather(zkfhe_ctx* ctx, void* d_buf, size_t bytes_per_rank) {
    if (!ctx->nccl_comm || !bytes_per_rank) return ZKFHE_OK;
    NcclApi* api = nccl();
    ZK_TRY(timed_begin(ctx, ZK_CAT_COMM, (uint64_t)bytes_per_rank * ctx->n_ranks));
    const int rc = api->AllGather((const char*)d_buf + (size_t)ctx->rank * bytes_per_rank, d_buf, bytes_per_rank, ncclInt8,
                                  (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != ncclSuccess) return fail(ctx, ZKFHE_ERR_CUDA, "ncclAllGather: %s", api->GetErrorString(rc));
    ZK_TRY(timed_end(ctx));
    ctx->comm_calls++;
    return ZKFHE_OK;
}

}  // namespace zkfhe

extern "C" {

int zkfhe_comm_unique_id(uint8_t* out128) {
    if (!out128) return ZKFHE_ERR_ARG;
    NcclApi* api = nccl();
    if (api->error) return ZKFHE_ERR_STATE;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return ZKFHE_ERR_CUDA;
    memcpy(out128, id.internal, 128);
    return ZKFHE_OK;
}

int zkfhe_comm_init(zkfhe_ctx* ctx, int rank, int n_ranks, const uint8_t* id128) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return zkfhe::fail(ctx, ZKFHE_ERR_ARG, "comm_init: bad rank %d of %d", rank, n_ranks);
    if (ctx->nccl_comm) return zkfhe::fail(ctx, ZKFHE_ERR_STATE, "comm_init: this context already has a communicator");
    if (ctx->virtual_ranks > 1) return zkfhe::fail(ctx, ZKFHE_ERR_STATE, "comm_init: the context is in virtual-rank mode");
    NcclApi* api = nccl();
    if (api->error) return zkfhe::fail(ctx, ZKFHE_ERR_STATE, "comm_init: %s", api->error);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    const int rc = api->CommInitRank(&comm, n_ranks, id, rank);
    if (rc != ncclSuccess) return zkfhe::fail(ctx, ZKFHE_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(rc));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    return ZKFHE_OK;
}

int zkfhe_comm_destroy(zkfhe_ctx* ctx) {
    if (!ctx) return ZKFHE_ERR_ARG;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        zkfhe::stream_wait(ctx);
        nccl()->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->n_ranks = 1;
    return ZKFHE_OK;
}

int zkfhe_set_virtual_ranks(zkfhe_ctx* ctx, int n_ranks) {
    if (!ctx || n_ranks < 0 || n_ranks > 64) return zkfhe::fail(ctx, ZKFHE_ERR_ARG, "set_virtual_ranks: %d out of range [0, 64]", n_ranks);
    if (ctx->nccl_comm) return zkfhe::fail(ctx, ZKFHE_ERR_STATE, "set_virtual_ranks: the context has an NCCL communicator");
    ctx->virtual_ranks = n_ranks;
    return ZKFHE_OK;
}

int zkfhe_comm_info(const zkfhe_ctx* ctx, int* rank, int* n_ranks, int* virtual_ranks) {
    if (!ctx) return ZKFHE_ERR_ARG;
    if (rank) *rank = ctx->rank;
    if (n_ranks) *n_ranks = ctx->n_ranks;
    if (virtual_ranks) *virtual_ranks = ctx->virtual_ranks;
    return ZKFHE_OK;
}

// The block of `count` items that shard `rank` of `n_ranks` owns (contiguous, ceil(count / n_ranks) each): the rule the
// prover uses for the columns of a commitment phase.  Pure host arithmetic.
int zkfhe_shard_range(uint32_t count, uint32_t n_ranks, uint32_t rank, uint32_t* lo, uint32_t* hi) {
    if (!lo || !hi || !n_ranks || rank >= n_ranks) return ZKFHE_ERR_ARG;
    zkfhe::shard_range(count, n_ranks, rank, lo, hi);
    return ZKFHE_OK;
}

}  // extern "C"
