// Context, error reporting, timing.  Replaces what `using CUDA` + `GPU()` provide the reference
// (ext/nuPGCMCUDAExt.jl:8-16,33).
#include "common.cuh"

char g_nupgcm_err[512] = "";

extern "C" int32_t nupgcm_version(void) { return NUPGCM_B200_VERSION; }

extern "C" int64_t nupgcm_solve_stats_size(void) { return (int64_t)sizeof(nupgcm_solve_stats); }

extern "C" const char *nupgcm_last_error(const nupgcm_ctx *ctx) {
    return ctx ? ctx->err : g_nupgcm_err;
}

extern "C" int32_t nupgcm_create(int32_t device, nupgcm_ctx **out) {
    if (!out) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return nupgcm_fail(nullptr, NUPGCM_ERR_NO_DEVICE,
                           "no CUDA device available (%s); libnupgcm_b200 has no CPU fallback",
                           e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "device index out of range");
    cudaDeviceProp prop;
    NUPGCM_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return nupgcm_fail(nullptr, NUPGCM_ERR_NO_DEVICE,
                           "device %s is not sm_100 (B200); this library is built for sm_100a only",
                           prop.name);
    NUPGCM_CUDA(nullptr, cudaSetDevice(device));
    nupgcm_ctx *ctx = (nupgcm_ctx *)calloc(1, sizeof(nupgcm_ctx));
    if (!ctx) return nupgcm_fail(nullptr, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    snprintf(ctx->name, sizeof(ctx->name), "%s", prop.name);
    ctx->coop_grid = ctx->sm_count;   // one persistent CTA per SM
    NUPGCM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    NUPGCM_CUDA(ctx, cudaEventCreate(&ctx->ev0));
    NUPGCM_CUDA(ctx, cudaEventCreate(&ctx->ev1));
    NUPGCM_CUDA(ctx, cudaEventCreate(&ctx->sev0));
    NUPGCM_CUDA(ctx, cudaEventCreate(&ctx->sev1));
    NUPGCM_CUDA(ctx, cudaMalloc(&ctx->d_barrier, 4 * sizeof(unsigned long long)));
    // reduction scratch of the persistent solvers: per-CTA slots + replicated broadcast words
    NUPGCM_CUDA(ctx, cudaMalloc(&ctx->d_partials, 2 * ((size_t)(8 + kPartialSlots) * ((ctx->coop_grid + 7) & ~7) + 8 * 32) * 16 + 65536));
    NUPGCM_CUDA(ctx, cudaMalloc(&ctx->d_scalars, 64 * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaMallocHost(&ctx->h_scalars, 64 * sizeof(double)));
    ctx->hist_cap = 1 << 16;
    NUPGCM_CUDA(ctx, cudaMalloc(&ctx->d_hist, ctx->hist_cap * sizeof(double)));
    *out = ctx;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_destroy(nupgcm_ctx *ctx) {
    if (!ctx) return NUPGCM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_barrier);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_scalars);
    cudaFreeHost(ctx->h_scalars);
    cudaFree(ctx->d_hist);
    cudaFree(ctx->d_ws);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->sev0);
    cudaEventDestroy(ctx->sev1);
    cudaStreamDestroy(ctx->stream);
    free(ctx);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_set_grid(nupgcm_ctx *ctx, int32_t grid) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, grid >= 1 && grid <= ctx->sm_count, "set_grid: grid must be in 1..SM count");
    ctx->coop_grid = grid;
    return NUPGCM_OK;
}

// Page-locked host memory for the buffers a host keeps its copy of the state in: uploads and
// downloads from/to it are plain DMA transfers (no staging copy inside the driver).
extern "C" int32_t nupgcm_host_alloc(nupgcm_ctx *ctx, int64_t bytes, void **out) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, out && bytes > 0, "host_alloc: bad argument");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_host_free(nupgcm_ctx *ctx, void *p) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    if (p) NUPGCM_CUDA(ctx, cudaFreeHost(p));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_synchronize(nupgcm_ctx *ctx) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_mem_status(nupgcm_ctx *ctx, size_t *free_bytes, size_t *total_bytes) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    size_t f = 0, t = 0;
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_device_info(nupgcm_ctx *ctx, int32_t *sm_count, int32_t *cc_major,
                                      int32_t *cc_minor, char *name) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (name) snprintf(name, 64, "%s", ctx->name);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_timer_start(nupgcm_ctx *ctx) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_timer_stop(nupgcm_ctx *ctx, float *ms) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NUPGCM_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float t = 0.f;
    NUPGCM_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
    if (ms) *ms = t;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_launch_count(nupgcm_ctx *ctx, int64_t *count) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    if (count) *count = ctx->launches;
    return NUPGCM_OK;
}
