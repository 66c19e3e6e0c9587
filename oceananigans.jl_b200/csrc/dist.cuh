// dist.cuh -- single-node multi-GPU plumbing: one process per GPU, NCCL communicator over NVLink/NVSwitch.
// Bootstrap mirrors ext/OceananigansNCCLExt/nccl_communicator.jl:25-63 (rank 0 makes the unique id, the host
// broadcasts it, every rank calls ncclCommInitRank).
#pragma once
#include <nccl.h>

extern "C" int32_t ob_dist_unique_id(void *out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return OB_OK;
}
extern "C" int32_t ob_dist_init(ob_ctx *ctx, int32_t rank, int32_t world, const void *id128) {
    if (!ctx) return fail(OB_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    NCCL_TRY(ncclCommInitRank(&comm, world, id, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return OB_OK;
}
extern "C" int32_t ob_dist_finalize(ob_ctx *ctx) {
    if (ctx && ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        ncclCommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
        ctx->world = 1;
        ctx->rank = 0;
    }
    return OB_OK;
}
