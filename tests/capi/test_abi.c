/* C-language caller of libnupgcm_b200.so: drives the ABI exactly as a foreign host (Julia ccall) does —
 * 1-based int64 CSR in, one call per solve, results out — and checks them against values the CPU oracle
 * wrote (tests/test_capi_c.py generates the input file).  Compiled with gcc against include/nupgcm_b200.h
 * only: no torch, no Python, no C++.
 *
 * Input file (native-endian binary): int64 n, nnz, itmax_gmres, niter_cg_expected;  double pscale, tol;
 * int64 rowptr[n+1] (1-based), int64 colidx[nnz] (1-based), double vals[nnz], double y[n],
 * double dinv[n], double x_cg_expected[n], double hist_gmres_expected[itmax_gmres+1], double x_gmres_expected[n]. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nupgcm_b200.h"

#define CHECK(call)                                                                           \
    do {                                                                                      \
        int32_t rc_ = (call);                                                                 \
        if (rc_ != NUPGCM_OK) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, (int)rc_, nupgcm_last_error(ctx)); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

static void *read_block(FILE *f, size_t count, size_t size) {
    void *p = malloc(count * size > 0 ? count * size : 1);
    if (!p || fread(p, size, count, f) != count) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

static double rel_diff(const double *a, const double *b, int64_t n) {
    double num = 0.0, den = 0.0;
    for (int64_t i = 0; i < n; ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return sqrt(num / (den > 0.0 ? den : 1.0));
}

int main(int argc, char **argv) {
    nupgcm_ctx *ctx = NULL;
    if (argc < 2) { fprintf(stderr, "usage: test_abi <input file>\n"); return 2; }
    if (nupgcm_solve_stats_size() != (int64_t)sizeof(nupgcm_solve_stats)) { fprintf(stderr, "stats size mismatch\n"); return 1; }
    if (nupgcm_version() / 100 != NUPGCM_B200_VERSION / 100) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int64_t hdr[4];
    double par[2];
    if (fread(hdr, sizeof(int64_t), 4, f) != 4 || fread(par, sizeof(double), 2, f) != 2) return 2;
    const int64_t n = hdr[0], nnz = hdr[1], itmax = hdr[2], cg_expected = hdr[3];
    int64_t *rowptr = read_block(f, (size_t)n + 1, sizeof(int64_t)), *colidx = read_block(f, (size_t)nnz, sizeof(int64_t));
    double *vals = read_block(f, (size_t)nnz, sizeof(double)), *y = read_block(f, (size_t)n, sizeof(double));
    double *dinv = read_block(f, (size_t)n, sizeof(double)), *x_cg = read_block(f, (size_t)n, sizeof(double));
    double *hist_ref = read_block(f, (size_t)itmax + 1, sizeof(double)), *x_gm = read_block(f, (size_t)n, sizeof(double));
    fclose(f);

    CHECK(nupgcm_create(0, &ctx));
    nupgcm_csr *A = NULL;
    nupgcm_vec *vy = NULL, *vx = NULL, *vd = NULL;
    CHECK(nupgcm_csr_create(ctx, n, n, nnz, rowptr, colidx, vals, /*index_base=*/1, /*drop_zeros=*/0, &A));
    CHECK(nupgcm_vec_create(ctx, n, &vy));
    CHECK(nupgcm_vec_create(ctx, n, &vx));
    CHECK(nupgcm_vec_create(ctx, n, &vd));
    CHECK(nupgcm_vec_upload(vy, y, n));
    CHECK(nupgcm_vec_upload(vd, dinv, n));
    double *x = malloc((size_t)n * sizeof(double)), *hist = malloc(((size_t)itmax + 2) * sizeof(double));
    nupgcm_solve_stats st;

    /* CG + Jacobi (CgWorkspace, src/evolution.jl:118-126) from x = 0 */
    CHECK(nupgcm_cg_solve(A, vd, 1.0, vy, vx, par[1], par[1], 0, NULL, 0, &st));
    CHECK(nupgcm_vec_download(vx, x, n));
    if (!st.solved || llabs((long long)(st.niter - cg_expected)) > 1 || rel_diff(x, x_cg, n) > 1e-8) {
        fprintf(stderr, "cg: solved %d niter %lld (expected %lld) rel diff %.3e\n", (int)st.solved, (long long)st.niter,
                (long long)cg_expected, rel_diff(x, x_cg, n));
        return 1;
    }
    /* GMRES(20), scalar preconditioner (GmresWorkspace, src/inversion.jl:74-94): fixed iteration count */
    CHECK(nupgcm_vec_fill(vx, 0.0));
    CHECK(nupgcm_gmres_solve(A, NULL, par[0], vy, vx, 0.0, 1e-30, itmax, 20, NUPGCM_ORTH_MGS, hist, itmax + 1, &st));
    CHECK(nupgcm_vec_download(vx, x, n));
    if (st.niter != itmax || st.hist_len != itmax + 1 || rel_diff(hist, hist_ref, itmax + 1) > 1e-9 || rel_diff(x, x_gm, n) > 1e-9) {
        fprintf(stderr, "gmres: niter %lld hist_len %lld hist diff %.3e x diff %.3e\n", (long long)st.niter,
                (long long)st.hist_len, rel_diff(hist, hist_ref, itmax + 1), rel_diff(x, x_gm, n));
        return 1;
    }
    /* errors are status codes, never aborts: a size mismatch must be refused */
    nupgcm_vec *bad = NULL;
    CHECK(nupgcm_vec_create(ctx, n + 1, &bad));
    if (nupgcm_cg_solve(A, vd, 1.0, bad, vx, 1e-6, 1e-6, 0, NULL, 0, &st) != NUPGCM_ERR_INVALID) {
        fprintf(stderr, "size mismatch was not refused\n");
        return 1;
    }
    nupgcm_vec_destroy(bad);
    nupgcm_vec_destroy(vy); nupgcm_vec_destroy(vx); nupgcm_vec_destroy(vd);
    nupgcm_csr_destroy(A);
    nupgcm_destroy(ctx);
    printf("test_abi ok: n=%lld cg %lld its, gmres %lld its\n", (long long)n, (long long)cg_expected, (long long)itmax);
    return 0;
}
