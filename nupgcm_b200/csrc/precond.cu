// Operator preconditioners of the inversion: CgPreconditioner and BlockDiagonalPreconditioner
// (reference src/preconditioners.jl:5-37 and :53-125), and GMRES driven with one of them.
//
// The reference builds them nowhere by default (src/inversion.jl:60 is commented out); they are its
// author's route to fewer outer iterations (scratch/inversion_log.md:139-157).  Structure here:
//   * a block's inverse is one persistent k_cg launch (csrc/krylov.cu) — Jacobi-preconditioned CG,
//     Krylov.jl default tolerances atol = rtol = sqrt(eps), capped at `itmax`, warm-started from the
//     block's previous answer exactly like `cgp.workspace.x` (:25);
//   * the outer GMRES(m) is Krylov.jl's recurrence orchestrated from the host:
//     with an operator preconditioner every Arnoldi step contains two whole inner solves
//     (~0.1-1 ms), so the per-step launches and the host-synchronised dot products of the outer
//     loop (~0.2 ms) are not what bounds it.
// Deviation, declared: the reference's GPU set-up preconditions the friction block's inner CG with
// ILU(0) (KrylovPreconditioners.kp_ilu0, :102-107); here it is Jacobi (the variant the reference
// keeps commented out at :109-116), because sparse triangular solves do not fit a persistent kernel.
#include <cmath>
#include <vector>

#include "common.cuh"

struct nupgcm_blockprec {
    nupgcm_ctx *ctx;
    const nupgcm_csr *P, *T;
    const nupgcm_vec *Pd, *Td;
    int64_t n1, n2, itmax_p, itmax_t;
    nupgcm_vec *bp, *xp, *bt, *xt;        // right-hand sides and (persistent) solutions of the blocks
    int64_t inner_iters, applies;
};

extern "C" int32_t nupgcm_blockprec_create(nupgcm_ctx *ctx, const nupgcm_csr *P, const nupgcm_vec *P_dinv,
                                           int64_t P_itmax, const nupgcm_csr *T, const nupgcm_vec *T_dinv,
                                           int64_t T_itmax, nupgcm_blockprec **out) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, out && P && T && P_dinv && T_dinv, "blockprec_create: NULL argument");
    NUPGCM_REQUIRE(ctx, P->n_rows == P->n_cols && T->n_rows == T->n_cols, "blockprec_create: blocks must be square");
    NUPGCM_REQUIRE(ctx, P_dinv->n == P->n_rows && T_dinv->n == T->n_rows, "blockprec_create: diagonal length mismatch");
    NUPGCM_REQUIRE(ctx, P_itmax >= 0 && T_itmax >= 0, "blockprec_create: negative itmax");
    nupgcm_blockprec *b = (nupgcm_blockprec *)calloc(1, sizeof(nupgcm_blockprec));
    if (!b) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    b->ctx = ctx;
    b->P = P; b->T = T; b->Pd = P_dinv; b->Td = T_dinv;
    b->n1 = P->n_rows; b->n2 = T->n_rows;
    b->itmax_p = P_itmax; b->itmax_t = T_itmax;
    int32_t rc = nupgcm_vec_create(ctx, b->n1, &b->bp);
    if (!rc) rc = nupgcm_vec_create(ctx, b->n1, &b->xp);      // zero-filled: workspace.x .= 0 (:20)
    if (!rc) rc = nupgcm_vec_create(ctx, b->n2, &b->bt);
    if (!rc) rc = nupgcm_vec_create(ctx, b->n2, &b->xt);
    if (rc) return rc;
    *out = b;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_blockprec_destroy(nupgcm_blockprec *b) {
    if (!b) return NUPGCM_OK;
    nupgcm_vec_destroy(b->bp);
    nupgcm_vec_destroy(b->xp);
    nupgcm_vec_destroy(b->bt);
    nupgcm_vec_destroy(b->xt);
    free(b);
    return NUPGCM_OK;
}

// y = M x: y[0:n1] = CG(P)⁻¹ x[0:n1], y[n1:] = CG(T)⁻¹ x[n1:]   (src/preconditioners.jl:118-125)
static int32_t blockprec_apply_raw(nupgcm_blockprec *b, const double *x, double *y) {
    nupgcm_ctx *ctx = b->ctx;
    const double tol = 1.4901161193847656e-08;                 // sqrt(eps): Krylov.jl's default atol = rtol
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(b->bp->d, x, (size_t)b->n1 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(b->bt->d, x + b->n1, (size_t)b->n2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    nupgcm_solve_stats st;
    int32_t rc = nupgcm_cg_solve(b->P, b->Pd, 1.0, b->bp, b->xp, tol, tol, b->itmax_p, nullptr, 0, &st);
    if (rc) return rc;
    b->inner_iters += st.niter;
    rc = nupgcm_cg_solve(b->T, b->Td, 1.0, b->bt, b->xt, tol, tol, b->itmax_t, nullptr, 0, &st);
    if (rc) return rc;
    b->inner_iters += st.niter;
    b->applies++;
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(y, b->xp->d, (size_t)b->n1 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(y + b->n1, b->xt->d, (size_t)b->n2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_blockprec_apply(nupgcm_blockprec *b, const nupgcm_vec *x, nupgcm_vec *y) {
    NUPGCM_REQUIRE(nullptr, b && x && y, "blockprec_apply: NULL argument");
    NUPGCM_REQUIRE(b->ctx, x->n == b->n1 + b->n2 && y->n == x->n && x->d != y->d, "blockprec_apply: length mismatch or aliasing");
    return blockprec_apply_raw(b, x->d, y->d);
}

extern "C" int32_t nupgcm_blockprec_info(const nupgcm_blockprec *b, int64_t *applies, int64_t *inner_iters) {
    NUPGCM_REQUIRE(nullptr, b, "blockprec is NULL");
    if (applies) *applies = b->applies;
    if (inner_iters) *inner_iters = b->inner_iters;
    return NUPGCM_OK;
}

// Krylov.jl sym_givens for reals (host twin of the device routine in krylov.cu)
static void sym_givens_h(double a, double b, double &c, double &s, double &rho) {
    if (b == 0.0) { c = (a == 0.0) ? 1.0 : std::copysign(1.0, a); s = 0.0; rho = std::fabs(a); }
    else if (a == 0.0) { c = 0.0; s = std::copysign(1.0, b); rho = std::fabs(b); }
    else if (std::fabs(b) > std::fabs(a)) { const double t = a / b; s = std::copysign(1.0, b) / std::sqrt(1.0 + t * t); c = s * t; rho = b / s; }
    else { const double t = b / a; c = std::copysign(1.0, a) / std::sqrt(1.0 + t * t); s = c * t; rho = a / c; }
}

// Restarted GMRES(memory), left-preconditioned by the operator M, MGS — Krylov.jl gmres! as called at
// src/iterative_solvers.jl:58 with P = BlockDiagonalPreconditioner (src/inversion.jl:60).
extern "C" int32_t nupgcm_gmres_solve_prec(const nupgcm_csr *A, nupgcm_blockprec *M, const nupgcm_vec *y,
                                           nupgcm_vec *x, double atol, double rtol, int64_t itmax,
                                           int32_t memory, double *resid_hist, int64_t hist_cap,
                                           nupgcm_solve_stats *stats) {
    NUPGCM_REQUIRE(nullptr, A && M && y && x, "gmres_solve_prec: NULL argument");
    nupgcm_ctx *ctx = A->ctx;
    const int64_t n = A->n_rows;
    NUPGCM_REQUIRE(ctx, A->n_rows == A->n_cols && y->n == n && x->n == n && M->n1 + M->n2 == n, "gmres_solve_prec: size mismatch");
    NUPGCM_REQUIRE(ctx, memory >= 1 && memory <= kMaxMemory && itmax >= 0 && atol >= 0.0 && rtol >= 0.0, "gmres_solve_prec: bad parameter");
    NUPGCM_REQUIRE(ctx, hist_cap >= 0 && (hist_cap == 0 || resid_hist), "gmres_solve_prec: hist_cap without buffer");
    NUPGCM_REQUIRE(ctx, !A->comm, "gmres_solve_prec: sharded matrices are not supported");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int mem = memory;
    double *buf = nullptr;
    NUPGCM_CUDA(ctx, cudaMalloc(&buf, (size_t)(mem + 3) * (size_t)n * sizeof(double)));
    std::vector<nupgcm_vec> V(mem + 1);
    for (int i = 0; i <= mem; ++i) V[i] = nupgcm_vec{ctx, n, buf + (size_t)i * n};
    nupgcm_vec w{ctx, n, buf + (size_t)(mem + 1) * n}, q{ctx, n, buf + (size_t)(mem + 2) * n};
    int32_t rc = NUPGCM_OK;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    const int64_t launches0 = ctx->launches;
#define STEP(call) do { rc = (call); if (rc) goto done; } while (0)
    {
        std::vector<double> c(mem, 0.0), s(mem, 0.0), R((size_t)mem * (mem + 1) / 2, 0.0), z(mem + 1, 0.0), yv(mem, 0.0);
        int64_t nhist = 0, iter = 0;
        const int64_t cap_it = itmax == 0 ? 2 * n : itmax;
        int64_t inner_itmax = cap_it;
        const double btol = 1.8189894035458565e-12;
        // w = b − A x0 ; r0 = M w   (x0 = content of x: warm start)
        STEP(nupgcm_vec_copy(&w, y));
        STEP(nupgcm_spmv(A, x, &w, -1.0, 1.0));
        STEP(blockprec_apply_raw(M, w.d, V[0].d));
        double beta = 0.0;
        STEP(nupgcm_vec_norm2(&V[0], &beta));
        double rnorm = beta;
        const double rnorm0 = beta;
        if (nhist < hist_cap) resid_hist[nhist] = rnorm;
        nhist++;
        const double eps_tol = atol + rtol * rnorm;
        bool solved = beta == 0.0 || rnorm <= eps_tol, tired = iter >= cap_it, breakdown = false, inconsistent = false;
        int npass = 0;
        while (!(solved || tired || breakdown)) {
            std::fill(c.begin(), c.end(), 0.0);
            std::fill(s.begin(), s.end(), 0.0);
            std::fill(R.begin(), R.end(), 0.0);
            std::fill(z.begin(), z.end(), 0.0);
            if (npass >= 1) {
                STEP(nupgcm_vec_copy(&w, y));
                STEP(nupgcm_spmv(A, x, &w, -1.0, 1.0));
                STEP(blockprec_apply_raw(M, w.d, V[0].d));
                STEP(nupgcm_vec_norm2(&V[0], &beta));
            }
            z[0] = beta;
            STEP(nupgcm_vec_axpby(&V[0], 0.0, &V[0], 1.0 / beta));       // V1 = r0 / β
            npass++;
            int k = 0, nr = 0;
            bool inner_tired = false;
            while (!(solved || inner_tired || breakdown)) {
                k++;
                STEP(nupgcm_spmv(A, &V[k - 1], &w, 1.0, 0.0));
                STEP(blockprec_apply_raw(M, w.d, q.d));
                for (int i = 0; i < k; ++i) {                           // modified Gram-Schmidt
                    double h = 0.0;
                    STEP(nupgcm_vec_dot(&V[i], &q, &h));
                    R[nr + i] = h;
                    STEP(nupgcm_vec_axpby(&q, -h, &V[i], 1.0));
                }
                double Hbis = 0.0;
                STEP(nupgcm_vec_norm2(&q, &Hbis));
                for (int i = 0; i < k - 1; ++i) {
                    const double tmp = c[i] * R[nr + i] + s[i] * R[nr + i + 1];
                    R[nr + i + 1] = s[i] * R[nr + i] - c[i] * R[nr + i + 1];
                    R[nr + i] = tmp;
                }
                sym_givens_h(R[nr + k - 1], Hbis, c[k - 1], s[k - 1], R[nr + k - 1]);
                const double zeta = s[k - 1] * z[k - 1];
                z[k - 1] = c[k - 1] * z[k - 1];
                rnorm = std::fabs(zeta);
                if (nhist < hist_cap) resid_hist[nhist] = rnorm;
                nhist++;
                nr += k;
                solved = rnorm <= eps_tol || rnorm + 1.0 <= 1.0;
                breakdown = Hbis <= btol;
                inner_tired = k >= (inner_itmax < mem ? inner_itmax : (int64_t)mem);
                if (!(solved || inner_tired || breakdown)) {
                    STEP(nupgcm_vec_copy(&V[k], &q));
                    STEP(nupgcm_vec_axpby(&V[k], 0.0, &V[k], 1.0 / Hbis));
                    z[k] = zeta;
                }
            }
            for (int i = 0; i < k; ++i) yv[i] = z[i];                   // back substitution R y = z
            for (int i = k; i >= 1; --i) {
                int pos = nr + i - k - 1;
                for (int j = k; j > i; --j) { yv[i - 1] -= R[pos] * yv[j - 1]; pos = pos - j + 1; }
                if (std::fabs(R[pos]) <= btol) { yv[i - 1] = 0.0; inconsistent = true; }
                else yv[i - 1] /= R[pos];
            }
            for (int i = 0; i < k; ++i) STEP(nupgcm_vec_axpby(x, yv[i], &V[i], 1.0));   // x += Σ y_i v_i
            inner_itmax -= k;
            iter += k;
            tired = iter >= cap_it;
        }
        cudaEventRecord(e1, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            stats->niter = iter;
            stats->solved = solved;
            stats->inconsistent = inconsistent;
            stats->breakdown = breakdown;
            stats->rnorm = rnorm;
            stats->rnorm0 = rnorm0;
            stats->hist_len = hist_cap > 0 ? (nhist < hist_cap ? nhist : hist_cap) : 0;
            stats->launches = (int32_t)(ctx->launches - launches0);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            stats->device_ms = ms;
        }
    }
done:
#undef STEP
    cudaStreamSynchronize(ctx->stream);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return rc;
}
