// CSR matrices on the device and the stand-alone FP64 SpMV.
//
// Replaces `CuSparseMatrixCSR(a)` (ext/nuPGCMCUDAExt.jl:27) and the `cusparseSpMV` behind
// `mul!(y, A, x)`; the same row kernel is instantiated inside the persistent Krylov kernels.
// Rows are processed by sub-warps of T = 2..32 lanes ("vector" CSR); T is chosen from the mean
// row length so that a lane sees ~4-8 entries (inversion matrix: ~50-77 per row -> T = 16;
// evolution matrix: ~23 per row -> T = 4).  Summation order inside a row is fixed (lane-strided
// partial sums, then an xor-shuffle tree), so results are run-to-run reproducible.
#include <algorithm>
#include <vector>

#include "common.cuh"

// ---- kernels ------------------------------------------------------------------------------

template <int T>
__global__ void __launch_bounds__(256)
k_spmv(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
       const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
       int64_t n_rows, double alpha, double beta) {
    const int lane = threadIdx.x & (T - 1);
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
    // the trip count is uniform over the warp so that the full-mask shuffles stay legal
    for (int64_t base = 0; base < n_rows; base += ngroups) {
        const int64_t row = base + group;
        const bool active = row < n_rows;
        double acc = 0.0;
        if (active) {
            const int32_t beg = rowptr[row], end = rowptr[row + 1];
            for (int32_t k = beg + lane; k < end; k += T)
                acc = fma(vals[k], __ldg(x + colidx[k]), acc);
        }
        acc = group_sum<T>(acc);
        if (active && lane == 0)
            y[row] = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * y[row]);
    }
}

__global__ void k_scatter_vals(double *vals, const double *stage, const int32_t *keep, int64_t nnz) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x)
        vals[i] = stage[keep[i]];
}

__global__ void k_combine(double *out, const double *m, const double *kh, const double *kv,
                          double theta, int64_t nnz) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = m[i] + theta * (kh[i] + kv[i]);     // M + θ*(Kₕ + Kᵥ), evolution.jl:145
}

__global__ void k_inv_diag(const int32_t *rowptr, const int32_t *colidx, const double *vals,
                           double *dinv, int64_t n_rows) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        double d = 0.0;
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
            if (colidx[k] == r) d += vals[k];
        dinv[r] = 1.0 / d;
    }
}

// ---- host helpers -------------------------------------------------------------------------

static int choose_tpr(double mean_row) {
    if (mean_row >= 96.0) return 32;
    if (mean_row >= 40.0) return 16;
    if (mean_row >= 20.0) return 8;
    if (mean_row >= 8.0) return 4;
    return 2;
}

// Row ranges of the persistent kernels: contiguous, balanced on (nnz + 4*rows).
static void build_partition(const std::vector<int32_t> &rowptr, int64_t n_rows, int parts,
                            std::vector<int32_t> &part) {
    part.assign(parts + 1, 0);
    const double total = (double)rowptr[n_rows] + 4.0 * (double)n_rows;
    int64_t r = 0;
    for (int p = 1; p < parts; ++p) {
        const double target = total * p / parts;
        while (r < n_rows && (double)rowptr[r] + 4.0 * (double)r < target) ++r;
        part[p] = (int32_t)r;
    }
    part[parts] = (int32_t)n_rows;
    for (int p = 1; p <= parts; ++p) part[p] = std::max(part[p], part[p - 1]);
}

// Row partition and SM-resident tables of the persistent solvers for a grid of `grid` CTAs.
// Cached in the handle; rebuilt only when a solve asks for a different grid.
int32_t nupgcm_csr_prepare(nupgcm_csr *A, int grid) {
    nupgcm_ctx *ctx = A->ctx;
    if (A->prepared_grid == grid) return NUPGCM_OK;
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(A->d_part); A->d_part = nullptr;
    cudaFree(A->d_loc); A->d_loc = nullptr;
    cudaFree(A->d_foot_ptr); A->d_foot_ptr = nullptr;
    cudaFree(A->d_foot); A->d_foot = nullptr;
    A->res_max_nnz = A->res_max_foot = A->res_max_rows = 0;
    const int64_t n_rows = A->n_rows, kept = A->nnz;
    std::vector<int32_t> h_rowptr(A->h_rowptr, A->h_rowptr + n_rows + 1), part;
    build_partition(h_rowptr, n_rows, grid, part);
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_part, part.size() * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemcpy(A->d_part, part.data(), part.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (A->n_rows == A->n_cols && kept > 0) {
        const int32_t *h_col = A->h_col;
        std::vector<uint16_t> loc(kept);
        std::vector<int32_t> foot_ptr(grid + 1, 0), foot, tmp;
        bool ok = true;
        int max_nnz = 0, max_foot = 0, max_rows = 0;
        for (int p = 0; p < grid && ok; ++p) {
            const int32_t k0 = h_rowptr[part[p]], k1 = h_rowptr[part[p + 1]];
            tmp.assign(h_col + k0, h_col + k1);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            if (tmp.size() > 65535) { ok = false; break; }
            for (int32_t k = k0; k < k1; ++k)
                loc[k] = (uint16_t)(std::lower_bound(tmp.begin(), tmp.end(), h_col[k]) - tmp.begin());
            foot.insert(foot.end(), tmp.begin(), tmp.end());
            foot_ptr[p + 1] = (int32_t)foot.size();
            max_nnz = std::max(max_nnz, (int)(k1 - k0));
            max_foot = std::max(max_foot, (int)tmp.size());
            max_rows = std::max(max_rows, (int)(part[p + 1] - part[p]));
        }
        if (ok) {
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_loc, ((size_t)kept + 16) * sizeof(uint16_t)));
            NUPGCM_CUDA(ctx, cudaMemset(A->d_loc, 0, ((size_t)kept + 16) * sizeof(uint16_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_loc, loc.data(), (size_t)kept * sizeof(uint16_t), cudaMemcpyHostToDevice));
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_foot_ptr, foot_ptr.size() * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_foot_ptr, foot_ptr.data(), foot_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_foot, (foot.size() + 8) * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemset(A->d_foot, 0, (foot.size() + 8) * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_foot, foot.data(), foot.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            A->res_max_nnz = max_nnz;
            A->res_max_foot = max_foot;
            A->res_max_rows = max_rows;
        }
    }
    A->prepared_grid = grid;
    return NUPGCM_OK;
}

// ---- C ABI --------------------------------------------------------------------------------

extern "C" int32_t nupgcm_csr_create(nupgcm_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                     const int64_t *rowptr, const int64_t *colidx,
                                     const double *vals, int32_t index_base, int32_t drop_zeros,
                                     nupgcm_csr **out) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, out && rowptr && (nnz == 0 || (colidx && vals)), "csr_create: NULL argument");
    NUPGCM_REQUIRE(ctx, n_rows >= 0 && n_cols >= 0 && nnz >= 0, "csr_create: negative size");
    NUPGCM_REQUIRE(ctx, n_rows < INT32_MAX && n_cols < INT32_MAX && nnz < INT32_MAX,
                   "csr_create: sizes exceed the int32 device index range");
    NUPGCM_REQUIRE(ctx, index_base == 0 || index_base == 1, "csr_create: index_base must be 0 or 1");
    NUPGCM_REQUIRE(ctx, rowptr[0] - index_base == 0 && rowptr[n_rows] - index_base == nnz,
                   "csr_create: rowptr does not span nnz");

    std::vector<int32_t> h_rowptr(n_rows + 1), h_col, h_keep;
    std::vector<double> h_val;
    h_col.reserve(nnz);
    h_val.reserve(nnz);
    if (drop_zeros) h_keep.reserve(nnz);
    h_rowptr[0] = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t b = rowptr[r] - index_base, e = rowptr[r + 1] - index_base;
        if (e < b || e > nnz)
            return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s", "csr_create: rowptr not monotone");
        for (int64_t k = b; k < e; ++k) {
            const int64_t c = colidx[k] - index_base;
            if (c < 0 || c >= n_cols)
                return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s", "csr_create: column index out of range");
            if (drop_zeros && vals[k] == 0.0) continue;
            h_col.push_back((int32_t)c);
            h_val.push_back(vals[k]);
            if (drop_zeros) h_keep.push_back((int32_t)k);
        }
        h_rowptr[r + 1] = (int32_t)h_col.size();
    }
    const int64_t kept = (int64_t)h_col.size();

    nupgcm_csr *A = (nupgcm_csr *)calloc(1, sizeof(nupgcm_csr));
    if (!A) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    A->ctx = ctx;
    A->n_rows = n_rows;
    A->n_cols = n_cols;
    A->nnz_given = nnz;
    A->nnz = kept;
    A->dropped = drop_zeros != 0;
    A->tpr = choose_tpr(n_rows ? (double)kept / (double)n_rows : 0.0);
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nz = (size_t)(kept > 0 ? kept : 1);
    // a few elements of padding let the persistent kernels bulk-copy 16-byte aligned windows
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_rowptr, (size_t)(n_rows + 1 + 8) * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemset(A->d_rowptr, 0, (size_t)(n_rows + 1 + 8) * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_colidx, nz * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_vals, (nz + 4) * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaMemset(A->d_vals, 0, (nz + 4) * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaMemcpy(A->d_rowptr, h_rowptr.data(), (size_t)(n_rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (kept) {
        NUPGCM_CUDA(ctx, cudaMemcpy(A->d_colidx, h_col.data(), kept * sizeof(int32_t), cudaMemcpyHostToDevice));
        NUPGCM_CUDA(ctx, cudaMemcpy(A->d_vals, h_val.data(), kept * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (drop_zeros) {
        NUPGCM_CUDA(ctx, cudaMalloc(&A->d_keep, nz * sizeof(int32_t)));
        if (kept)
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_keep, h_keep.data(), kept * sizeof(int32_t), cudaMemcpyHostToDevice));
        NUPGCM_CUDA(ctx, cudaMalloc(&A->d_stage, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double)));
    }
    // host copies of the structure: the persistent solvers derive their row partition and
    // SM-resident tables from them for whatever grid size a solve uses (nupgcm_csr_prepare)
    A->h_rowptr = (int32_t *)malloc((size_t)(n_rows + 1) * sizeof(int32_t));
    A->h_col = (int32_t *)malloc(nz * sizeof(int32_t));
    if (!A->h_rowptr || !A->h_col) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    memcpy(A->h_rowptr, h_rowptr.data(), (size_t)(n_rows + 1) * sizeof(int32_t));
    if (kept) memcpy(A->h_col, h_col.data(), (size_t)kept * sizeof(int32_t));
    A->prepared_grid = 0;
    *out = A;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_csr_destroy(nupgcm_csr *A) {
    if (!A) return NUPGCM_OK;
    cudaStreamSynchronize(A->ctx->stream);
    cudaFree(A->d_rowptr);
    cudaFree(A->d_colidx);
    cudaFree(A->d_vals);
    cudaFree(A->d_keep);
    cudaFree(A->d_stage);
    cudaFree(A->d_part);
    cudaFree(A->d_loc);
    cudaFree(A->d_foot_ptr);
    cudaFree(A->d_foot);
    free(A->h_rowptr);
    free(A->h_col);
    free(A);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_csr_info(const nupgcm_csr *A, int64_t *n_rows, int64_t *n_cols,
                                   int64_t *nnz_given, int64_t *nnz_stored) {
    NUPGCM_REQUIRE(nullptr, A, "csr is NULL");
    if (n_rows) *n_rows = A->n_rows;
    if (n_cols) *n_cols = A->n_cols;
    if (nnz_given) *nnz_given = A->nnz_given;
    if (nnz_stored) *nnz_stored = A->nnz;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_csr_update_values(nupgcm_csr *A, const double *vals, int64_t nnz) {
    NUPGCM_REQUIRE(nullptr, A, "csr is NULL");
    nupgcm_ctx *ctx = A->ctx;
    NUPGCM_REQUIRE(ctx, vals && nnz == A->nnz_given, "csr_update_values: NULL values or nnz mismatch");
    if (!A->dropped) {
        NUPGCM_CUDA(ctx, cudaMemcpyAsync(A->d_vals, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // pattern of kept entries is frozen at creation: entries that were zero then stay dropped
        NUPGCM_CUDA(ctx, cudaMemcpyAsync(A->d_stage, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (A->nnz) {
            int g = (int)std::min<int64_t>((A->nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
            k_scatter_vals<<<g, 256, 0, ctx->stream>>>(A->d_vals, A->d_stage, A->d_keep, A->nnz);
            ctx->launches++;
            NUPGCM_CUDA(ctx, cudaGetLastError());
        }
    }
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_csr_combine(nupgcm_csr *out, const nupgcm_csr *M, const nupgcm_csr *Kh,
                                      const nupgcm_csr *Kv, double theta) {
    NUPGCM_REQUIRE(nullptr, out && M && Kh && Kv, "csr_combine: NULL argument");
    nupgcm_ctx *ctx = out->ctx;
    NUPGCM_REQUIRE(ctx, !out->dropped && !M->dropped && !Kh->dropped && !Kv->dropped,
                   "csr_combine: operands must be created with drop_zeros=0 (shared pattern)");
    NUPGCM_REQUIRE(ctx, out->nnz == M->nnz && out->nnz == Kh->nnz && out->nnz == Kv->nnz &&
                            out->n_rows == M->n_rows,
                   "csr_combine: operands do not share a pattern");
    if (out->nnz == 0) return NUPGCM_OK;
    int g = (int)std::min<int64_t>((out->nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_combine<<<g, 256, 0, ctx->stream>>>(out->d_vals, M->d_vals, Kh->d_vals, Kv->d_vals, theta, out->nnz);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_csr_inv_diag(const nupgcm_csr *A, nupgcm_vec *dinv) {
    NUPGCM_REQUIRE(nullptr, A && dinv, "csr_inv_diag: NULL argument");
    nupgcm_ctx *ctx = A->ctx;
    NUPGCM_REQUIRE(ctx, A->n_rows == A->n_cols && dinv->n == A->n_rows, "csr_inv_diag: size mismatch");
    if (A->n_rows == 0) return NUPGCM_OK;
    int g = (int)std::min<int64_t>((A->n_rows + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_inv_diag<<<g, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_colidx, A->d_vals, dinv->d, A->n_rows);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    return NUPGCM_OK;
}

template <int T>
static void launch_spmv(const nupgcm_csr *A, const double *x, double *y, double alpha, double beta) {
    nupgcm_ctx *ctx = A->ctx;
    const int64_t threads = A->n_rows * T;
    int64_t g = (threads + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;   // 8 CTAs of 256 threads fill an SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    k_spmv<T><<<(int)g, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_colidx, A->d_vals, x, y, A->n_rows, alpha, beta);
}

extern "C" int32_t nupgcm_spmv(const nupgcm_csr *A, const nupgcm_vec *x, nupgcm_vec *y,
                               double alpha, double beta) {
    NUPGCM_REQUIRE(nullptr, A && x && y, "spmv: NULL argument");
    nupgcm_ctx *ctx = A->ctx;
    NUPGCM_REQUIRE(ctx, x->n >= A->n_cols && y->n == A->n_rows, "spmv: vector length mismatch");
    NUPGCM_REQUIRE(ctx, x->d != y->d, "spmv: x and y must not alias");
    if (A->n_rows == 0) return NUPGCM_OK;
    switch (A->tpr) {
        case 32: launch_spmv<32>(A, x->d, y->d, alpha, beta); break;
        case 16: launch_spmv<16>(A, x->d, y->d, alpha, beta); break;
        case 8: launch_spmv<8>(A, x->d, y->d, alpha, beta); break;
        case 4: launch_spmv<4>(A, x->d, y->d, alpha, beta); break;
        default: launch_spmv<2>(A, x->d, y->d, alpha, beta); break;
    }
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    return NUPGCM_OK;
}
