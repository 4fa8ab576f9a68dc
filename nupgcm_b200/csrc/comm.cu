// Communicator of the row-block sharded solves (multi-GPU).  The reference is single-GPU
// (ext/nuPGCMCUDAExt.jl); BASELINE.json's north_star shards the Krylov path across the B200s of
// one NVSwitch box.  A communicator is nothing but a device-memory arena per rank that all other
// ranks can store into: peers are mapped with CUDA IPC (one process per GPU) or used directly
// (several ranks in one process).  The persistent solver kernels (krylov.cu) do all the
// communication themselves with remote stores and flagged words; there is no host-side collective
// on the solve path.
#include "common.cuh"

extern "C" int32_t nupgcm_comm_create(nupgcm_ctx *ctx, int32_t rank, int32_t nranks, int64_t max_n,
                                      nupgcm_comm **out) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, out, "comm_create: out is NULL");
    NUPGCM_REQUIRE(ctx, nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks,
                   "comm_create: rank / nranks out of range (at most 8 ranks)");
    NUPGCM_REQUIRE(ctx, max_n >= 1 && max_n < INT32_MAX, "comm_create: max_n out of range");
    nupgcm_comm *c = (nupgcm_comm *)calloc(1, sizeof(nupgcm_comm));
    if (!c) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    c->ctx = ctx;
    c->rank = rank;
    c->nranks = nranks;
    c->max_n = max_n;
    c->xgen = 1;
    c->n_pad = (max_n + 15) & ~(int64_t)15;
    c->arena_bytes = kArenaVecOffset + 3 * (size_t)c->n_pad * (sizeof(double) + 16);   // plain + flagged forms
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(&c->arena, c->arena_bytes);
    if (e != cudaSuccess) {
        free(c);
        return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "device allocation failed: %s", cudaGetErrorString(e));
    }
    NUPGCM_CUDA(ctx, cudaMemset(c->arena, 0, c->arena_bytes));
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());
    c->peer[rank] = c->arena;
    if (nranks == 1) c->connected = 1;
    *out = c;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_comm_destroy(nupgcm_comm *c) {
    if (!c) return NUPGCM_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (int p = 0; p < c->nranks; ++p)
        if (c->ipc_mapped[p] && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    cudaFree(c->arena);
    free(c);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_comm_ipc_handle(nupgcm_comm *c, void *handle_out) {
    NUPGCM_REQUIRE(nullptr, c, "comm is NULL");
    NUPGCM_REQUIRE(c->ctx, handle_out, "comm_ipc_handle: NULL buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == NUPGCM_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    NUPGCM_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    NUPGCM_CUDA(c->ctx, cudaIpcGetMemHandle(&h, c->arena));
    memcpy(handle_out, &h, sizeof(h));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_comm_connect_ipc(nupgcm_comm *c, const void *handles) {
    NUPGCM_REQUIRE(nullptr, c, "comm is NULL");
    NUPGCM_REQUIRE(c->ctx, handles, "comm_connect_ipc: NULL handles");
    NUPGCM_REQUIRE(c->ctx, !c->connected || c->nranks == 1, "comm_connect: already connected");
    NUPGCM_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    for (int p = 0; p < c->nranks; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)p * NUPGCM_IPC_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        NUPGCM_CUDA(c->ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = (char *)ptr;
        c->ipc_mapped[p] = 1;
    }
    c->connected = 1;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_comm_connect_local(nupgcm_comm *c, nupgcm_comm *const *all) {
    NUPGCM_REQUIRE(nullptr, c, "comm is NULL");
    NUPGCM_REQUIRE(c->ctx, all, "comm_connect_local: NULL list");
    NUPGCM_REQUIRE(c->ctx, !c->connected || c->nranks == 1, "comm_connect: already connected");
    NUPGCM_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    for (int p = 0; p < c->nranks; ++p) {
        const nupgcm_comm *o = all[p];
        NUPGCM_REQUIRE(c->ctx, o && o->rank == p && o->nranks == c->nranks && o->max_n == c->max_n,
                       "comm_connect_local: communicators do not form one group ordered by rank");
        if (p == c->rank) continue;
        if (o->ctx->device != c->ctx->device) {
            int can = 0;
            NUPGCM_CUDA(c->ctx, cudaDeviceCanAccessPeer(&can, c->ctx->device, o->ctx->device));
            if (!can) return nupgcm_fail(c->ctx, NUPGCM_ERR_CUDA, "%s", "comm_connect_local: no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else NUPGCM_CUDA(c->ctx, e);
        }
        c->peer[p] = o->arena;
    }
    c->connected = 1;
    return NUPGCM_OK;
}
