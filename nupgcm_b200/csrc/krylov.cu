// Device-resident Krylov solvers: CG (Jacobi) and restarted GMRES(m), each ONE persistent
// cooperative kernel per solve.
//
// Replaces Krylov.krylov_solve! as called at reference src/iterative_solvers.jl:58 (GMRES
// workspace src/inversion.jl:74-94, CG workspace src/evolution.jl:118-126).  The reference's
// CUDA.jl path issues ~25 library launches and ~12 host synchronisations per GMRES iteration
// (SURVEY.md §2.1, §6); here the whole solve — SpMV, preconditioner apply, dot products, vector
// updates, Givens/QR bookkeeping, stopping test, restarts — runs inside one kernel:
//
//   * one CTA of kThreads threads per SM, launched cooperatively so that all CTAs are
//     co-resident; CTA b owns a contiguous row range (balanced on nnz) of the matrix and the same
//     index range of every vector, so all vector updates are CTA-local;
//   * the only cross-CTA traffic is (i) the gather of the multiplied vector in the SpMV, read
//     with ld.global.cg (L2-coherent), and (ii) one 8-byte partial per CTA per dot product,
//     combined after a grid barrier in a FIXED order (slot = CTA index; lane-strided sum, then
//     xor-shuffle tree).  Every CTA therefore computes bit-identical scalars and takes identical
//     branches; results are run-to-run reproducible (no floating-point atomics anywhere);
//   * scalar recurrences (Givens rotations, packed R, back substitution) are replicated per CTA.
//
// The recurrences restate Krylov.jl 0.10 `cg!` and `gmres!`: warm start
// (x on entry is the initial guess), preconditioned stopping measure against
// atol + rtol*(initial measure), itmax = 2n when 0, MGS Arnoldi (optionally CGS2).
#include <algorithm>

#include "common.cuh"

static const int kThreads = 1024;

struct KrylovArgs {
    const int32_t *rowptr;
    const int32_t *colidx;
    const double *vals;
    const int32_t *part;     // [grid+1] row ranges
    int n;
    const double *dinv;      // diagonal preconditioner or NULL
    double pscale;           // scalar preconditioner when dinv == NULL
    const double *b;
    double *x;
    double *work;            // CG: r, p, Ap (3n)   GMRES: V ((mem+1) n) then qbuf (2n)
    double atol, rtol;
    long long itmax;
    int mem;
    int orth;
    unsigned long long *barrier;
    double *partials;        // [2][kPartialSlots][grid]
    double *hist;
    long long hist_cap;
    double *result;          // niter, solved, inconsistent, breakdown, rnorm, rnorm0, hist_len, aborted
};

// ---- grid-wide reductions -------------------------------------------------------------------

struct GridReduce {
    GridBarrier bar;
    double *partials;
    int bank;
    int grid, bid;

    __device__ __forceinline__ void init(unsigned long long *counter, double *p) {
        bar.init(counter);
        partials = p;
        bank = 0;
        grid = gridDim.x;
        bid = blockIdx.x;
    }
    __device__ __forceinline__ double *slot(int j) {
        return partials + ((size_t)(bank * kPartialSlots + j)) * grid;
    }
    // plain barrier (publishes this CTA's vector rows to the other CTAs)
    __device__ __forceinline__ void barrier() { bar.sync(); }

    // Sum over all CTAs of `count` per-CTA values held in sm_in[0..count) (written by the caller
    // before a __syncthreads()).  Results land in sm_out[0..count), valid for every thread.
    __device__ __forceinline__ void sum(int count, const double *sm_in, double *sm_out) {
        if ((int)threadIdx.x < count) slot(threadIdx.x)[bid] = sm_in[threadIdx.x];
        bar.sync();
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int j = wid; j < count; j += (blockDim.x >> 5)) {
            const double *s = slot(j);
            double acc = 0.0;
            for (int i = lane; i < grid; i += 32) acc += ld_cg(s + i);
            acc = warp_sum(acc);
            if (lane == 0) sm_out[j] = acc;
        }
        bank ^= 1;
        __syncthreads();
    }
};

// Block partial of one value -> sm[0] (then __syncthreads so that GridReduce::sum can read it).
__device__ __forceinline__ void block_partial_to(double v, double *red, double *dst) {
    v = block_sum(v, red);
    if (threadIdx.x == 0) *dst = v;
    __syncthreads();
}

// ---- SpMV over the CTA's rows -----------------------------------------------------------------
// f(row, (A xin)[row]) is called by one lane per row.  xin is read with ld.cg because its rows
// are written by other CTAs earlier in the same kernel.
template <int T, class F>
__device__ __forceinline__ void spmv_rows(const KrylovArgs &a, const double *xin, int r0, int r1,
                                          F &&f) {
    const int lane = threadIdx.x & (T - 1);
    const int g = threadIdx.x / T;
    const int G = blockDim.x / T;
    for (int base = r0; base < r1; base += G) {     // uniform trip count over the CTA
        const int row = base + g;
        const bool active = row < r1;
        double acc = 0.0;
        if (active) {
            const int32_t beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
            int32_t k = beg + lane;
            // two independent accumulation chains keep more loads in flight per lane
            double acc2 = 0.0;
            for (; k + T < end; k += 2 * T) {
                const double v0 = __ldg(a.vals + k), v1 = __ldg(a.vals + k + T);
                const int32_t c0 = __ldg(a.colidx + k), c1 = __ldg(a.colidx + k + T);
                acc = fma(v0, ld_cg(xin + c0), acc);
                acc2 = fma(v1, ld_cg(xin + c1), acc2);
            }
            if (k < end) acc = fma(__ldg(a.vals + k), ld_cg(xin + __ldg(a.colidx + k)), acc);
            acc += acc2;
        }
        acc = group_sum<T>(acc);
        if (active && lane == 0) f(row, acc);
    }
}

__device__ __forceinline__ double precond(const KrylovArgs &a, int row, double v) {
    return a.dinv ? __ldg(a.dinv + row) * v : a.pscale * v;
}

// =============================================================================================
// CG  (Krylov.jl cg!)
// =============================================================================================
template <int T>
__global__ void __launch_bounds__(kThreads, 1) k_cg(KrylovArgs a) {
    __shared__ double red[32];
    __shared__ double sm_in[4], sm_out[4];
    GridReduce gr;
    gr.init(a.barrier, a.partials);
    const int r0 = a.part[blockIdx.x], r1 = a.part[blockIdx.x + 1];
    const int n = a.n;
    double *r = a.work, *p = a.work + n, *Ap = a.work + 2 * (size_t)n;
    double *x = a.x;                      // holds Δx (the warm start) until the very end
    const double eps = 2.220446049250313e-16;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    long long nhist = 0;

    // r = b − A Δx ; z = M r ; p = z ; γ = r·z          (own rows; z is never stored)
    double acc = 0.0;
    {
        double part = 0.0;
        spmv_rows<T>(a, x, r0, r1, [&](int row, double ax) {
            const double rv = a.b[row] - ax;
            const double zv = precond(a, row, rv);
            r[row] = rv;
            p[row] = zv;
            part = fma(rv, zv, part);
        });
        acc = part;
    }
    block_partial_to(acc, red, &sm_in[0]);
    gr.sum(1, sm_in, sm_out);             // also publishes p for the first SpMV
    double gamma = sm_out[0];
    double rnorm = sqrt(gamma);
    const double rnorm0 = rnorm;
    if (lead && a.hist && nhist < a.hist_cap) a.hist[nhist] = rnorm;
    nhist++;

    long long iter = 0;
    const long long itmax = a.itmax == 0 ? 2LL * n : a.itmax;
    double pnorm2 = gamma;
    const double tol = a.atol + a.rtol * rnorm;
    bool solved = (gamma == 0.0) || (rnorm <= tol);
    bool tired = iter >= itmax;
    bool zero_curv = false;
    // Krylov.jl accumulates the iterate from 0 and adds Δx at the end; here α p is accumulated
    // straight into x (which holds Δx): the same sum up to the rounding of one addition per entry.
    while (!(solved || tired || zero_curv || gr.bar.aborted())) {
        // Ap = A p ; pAp = p·Ap
        double part = 0.0;
        spmv_rows<T>(a, p, r0, r1, [&](int row, double ap) {
            Ap[row] = ap;
            part = fma(p[row], ap, part);
        });
        block_partial_to(part, red, &sm_in[0]);
        gr.sum(1, sm_in, sm_out);
        const double pAp = sm_out[0];
        if (pAp <= eps * pnorm2 && fabs(pAp) <= eps * pnorm2) {
            zero_curv = true;
            continue;
        }
        const double alpha = gamma / pAp;
        // x += α p ; r −= α Ap ; z = M r ; γ⁺ = r·z       (own rows, thread-per-row)
        part = 0.0;
        for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x) {
            x[row] = fma(alpha, p[row], x[row]);
            const double rv = fma(-alpha, Ap[row], r[row]);
            r[row] = rv;
            part = fma(rv, precond(a, row, rv), part);
        }
        block_partial_to(part, red, &sm_in[0]);
        gr.sum(1, sm_in, sm_out);
        const double gamma_next = sm_out[0];
        rnorm = sqrt(gamma_next);
        if (lead && a.hist && nhist < a.hist_cap) a.hist[nhist] = rnorm;
        nhist++;
        solved = (rnorm <= tol) || (rnorm + 1.0 <= 1.0);
        if (!solved) {
            const double beta = gamma_next / gamma;
            pnorm2 = gamma_next + beta * beta * pnorm2;
            gamma = gamma_next;
            // p = z + β p   (own rows), then publish p for the next SpMV gather
            for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x)
                p[row] = fma(beta, p[row], precond(a, row, r[row]));
            gr.barrier();
        }
        iter++;
        tired = iter >= itmax;
    }
    if (lead) {
        a.result[0] = (double)iter;
        a.result[1] = solved ? 1.0 : 0.0;
        a.result[2] = zero_curv ? 1.0 : 0.0;
        a.result[3] = 0.0;
        a.result[4] = rnorm;
        a.result[5] = rnorm0;
        a.result[6] = (double)(nhist < a.hist_cap ? nhist : a.hist_cap);
        a.result[7] = gr.bar.aborted() ? 1.0 : 0.0;
    }
}

// =============================================================================================
// GMRES(m), restarted, left preconditioned  (Krylov.jl gmres!)
// =============================================================================================

// Krylov.jl sym_givens for reals: [c s; s −c][a; b] = [ρ; 0]
__device__ __forceinline__ void sym_givens(double a, double b, double &c, double &s, double &rho) {
    if (b == 0.0) {
        c = (a == 0.0) ? 1.0 : copysign(1.0, a);
        s = 0.0;
        rho = fabs(a);
    } else if (a == 0.0) {
        c = 0.0;
        s = copysign(1.0, b);
        rho = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        const double t = a / b;
        s = copysign(1.0, b) / sqrt(1.0 + t * t);
        c = s * t;
        rho = b / s;
    } else {
        const double t = b / a;
        c = copysign(1.0, a) / sqrt(1.0 + t * t);
        s = c * t;
        rho = a / c;
    }
}

template <int T>
__global__ void __launch_bounds__(kThreads, 1) k_gmres(KrylovArgs a) {
    __shared__ double red[32];
    __shared__ double sm_in[kPartialSlots], sm_out[kPartialSlots];
    __shared__ double sc[kMaxMemory], ss[kMaxMemory], sz[kMaxMemory + 1], sy[kMaxMemory + 1];
    __shared__ double sR[kMaxMemory * (kMaxMemory + 1) / 2];
    __shared__ double s_flags[4];      // rnorm, solved, breakdown, Hbis
    __shared__ double s_seg[32];

    GridReduce gr;
    gr.init(a.barrier, a.partials);
    const int r0 = a.part[blockIdx.x], r1 = a.part[blockIdx.x + 1];
    const int n = a.n;
    const int mem = a.mem;
    double *V = a.work;                                  // V[i] = V + i*n, i = 0..mem
    double *qbuf = a.work + (size_t)(mem + 1) * n;       // two raw buffers for the SpMV gather
    double *x = a.x;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthr >> 5;
    const double btol = 1.8189894035458565e-12;          // eps^(3/4)
    long long nhist = 0;

    // ---- initial residual: w = b − A x0 ; r0 = M w (raw into qbuf[0]) ; β = ‖r0‖
    int cur = 0;
    {
        double part = 0.0;
        double *q0 = qbuf;
        spmv_rows<T>(a, x, r0, r1, [&](int row, double ax) {
            const double v = precond(a, row, a.b[row] - ax);
            q0[row] = v;
            part = fma(v, v, part);
        });
        block_partial_to(part, red, &sm_in[0]);
        gr.sum(1, sm_in, sm_out);
    }
    double beta = sqrt(sm_out[0]);
    double rnorm = beta;
    const double rnorm0 = beta;
    if (lead && a.hist && nhist < a.hist_cap) a.hist[nhist] = rnorm;
    nhist++;
    const double tol = a.atol + a.rtol * rnorm;
    long long iter = 0;
    const long long itmax = a.itmax == 0 ? 2LL * n : a.itmax;
    long long inner_itmax = itmax;
    bool breakdown = false, inconsistent = false;
    bool solved = (beta == 0.0) || (rnorm <= tol);
    bool tired = iter >= itmax;
    int npass = 0;

    while (!(solved || tired || breakdown || gr.bar.aborted())) {
        // ---- start of a pass ----
        if (tid < kMaxMemory) { sc[tid] = 0.0; ss[tid] = 0.0; }
        if (tid <= kMaxMemory) sz[tid] = 0.0;
        for (int i = tid; i < kMaxMemory * (kMaxMemory + 1) / 2; i += nthr) sR[i] = 0.0;
        if (npass >= 1) {
            // w = b − A x ; r0 = M w ; β = ‖r0‖   (x was published by the barrier ending the last pass)
            double part = 0.0;
            double *q0 = qbuf + (size_t)cur * n;
            spmv_rows<T>(a, x, r0, r1, [&](int row, double ax) {
                const double v = precond(a, row, a.b[row] - ax);
                q0[row] = v;
                part = fma(v, v, part);
            });
            block_partial_to(part, red, &sm_in[0]);
            gr.sum(1, sm_in, sm_out);
            beta = sqrt(sm_out[0]);
        }
        __syncthreads();
        if (tid == 0) sz[0] = beta;
        // V[0] = r0/β on own rows; other CTAs read the raw vector scaled by inv_h
        double inv_h = 1.0 / beta;
        {
            const double *q0 = qbuf + (size_t)cur * n;
            for (int row = r0 + tid; row < r1; row += nthr) V[row] = q0[row] * inv_h;
        }
        npass++;
        int k = 0;          // inner_iter
        int nr = 0;
        bool inner_tired = false;
        while (!(solved || inner_tired || breakdown || gr.bar.aborted())) {
            k++;
            // ---- q = M A v_k on own rows (raw v_k gathered from qbuf[cur], scaled by inv_h)
            double *q = V + (size_t)k * n;               // slot of the next basis vector
            const double *src = qbuf + (size_t)cur * n;
            double *dst = qbuf + (size_t)(cur ^ 1) * n;
            spmv_rows<T>(a, src, r0, r1, [&](int row, double av) {
                q[row] = precond(a, row, av * inv_h);
            });
            __syncthreads();
            double hsq = 0.0;                            // ‖q‖² after orthogonalisation
            if (a.orth == NUPGCM_ORTH_MGS) {
                // h_i = v_i·q ; q −= h_i v_i, sequentially (one grid reduction per i)
                double hprev = 0.0;
                for (int i = 0; i < k; ++i) {
                    const double *vi = V + (size_t)i * n;
                    const double *vp = V + (size_t)(i > 0 ? i - 1 : 0) * n;
                    double part = 0.0;
                    for (int row = r0 + tid; row < r1; row += nthr) {
                        double qv = q[row];
                        if (i > 0) { qv = fma(-hprev, vp[row], qv); q[row] = qv; }
                        part = fma(vi[row], qv, part);
                    }
                    block_partial_to(part, red, &sm_in[0]);
                    gr.sum(1, sm_in, sm_out);
                    hprev = sm_out[0];
                    if (tid == 0) sR[nr + i] = hprev;
                }
                const double *vp = V + (size_t)(k - 1) * n;
                double part = 0.0;
                for (int row = r0 + tid; row < r1; row += nthr) {
                    const double qv = fma(-hprev, vp[row], q[row]);
                    q[row] = qv;
                    dst[row] = qv;
                    part = fma(qv, qv, part);
                }
                block_partial_to(part, red, &sm_in[0]);
                gr.sum(1, sm_in, sm_out);
                hsq = sm_out[0];
            } else {
                // CGS2: all k projections at once, twice; warp w handles basis vector w % k on
                // row segment w / k.
                const int nseg = nwarps / k > 0 ? nwarps / k : 1;
                const int rows = r1 - r0;
                const int seglen = (rows + nseg - 1) / nseg;
                for (int pass = 0; pass < 2; ++pass) {
                    __syncthreads();
                    for (int u = wid; u < k * nseg; u += nwarps) {
                        const int i = u % k, sg = u / k;
                        const double *vi = V + (size_t)i * n;
                        const int lo = r0 + sg * seglen, hi = min(r1, lo + seglen);
                        double part = 0.0;
                        for (int row = lo + lane; row < hi; row += 32) part = fma(vi[row], q[row], part);
                        part = warp_sum(part);
                        if (lane == 0) s_seg[u % 32] = part;   // k*nseg <= 32 when k <= nwarps
                        __syncwarp();
                        // segments of one vector are summed in order by the first segment's warp
                    }
                    __syncthreads();
                    if (tid < k) {
                        double t = 0.0;
                        for (int sg = 0; sg < nseg; ++sg) t += s_seg[sg * k + tid];
                        sm_in[tid] = t;
                    }
                    __syncthreads();
                    gr.sum(k, sm_in, sm_out);
                    // q −= Σ h_i v_i  (own rows)
                    double part = 0.0;
                    for (int row = r0 + tid; row < r1; row += nthr) {
                        double qv = q[row];
                        for (int i = 0; i < k; ++i) qv = fma(-sm_out[i], V[(size_t)i * n + row], qv);
                        q[row] = qv;
                        if (pass == 1) { dst[row] = qv; part = fma(qv, qv, part); }
                    }
                    if (tid < k) sR[nr + tid] = (pass == 0) ? sm_out[tid] : sR[nr + tid] + sm_out[tid];
                    if (pass == 1) {
                        block_partial_to(part, red, &sm_in[0]);
                        gr.sum(1, sm_in, sm_out);
                        hsq = sm_out[0];
                    }
                }
            }
            // ---- scalar recurrences, replicated per CTA (thread 0), Krylov.jl order
            __syncthreads();
            if (tid == 0) {
                const double Hbis = sqrt(hsq);
                for (int i = 0; i < k - 1; ++i) {
                    const double tmp = sc[i] * sR[nr + i] + ss[i] * sR[nr + i + 1];
                    sR[nr + i + 1] = ss[i] * sR[nr + i] - sc[i] * sR[nr + i + 1];
                    sR[nr + i] = tmp;
                }
                double c, s, rho;
                sym_givens(sR[nr + k - 1], Hbis, c, s, rho);
                sc[k - 1] = c;
                ss[k - 1] = s;
                sR[nr + k - 1] = rho;
                const double zeta = s * sz[k - 1];
                sz[k - 1] = c * sz[k - 1];
                sz[k] = zeta;                 // only consumed if the pass continues
                s_flags[0] = fabs(zeta);
                s_flags[3] = Hbis;
            }
            __syncthreads();
            rnorm = s_flags[0];
            const double Hbis = s_flags[3];
            if (lead && a.hist && nhist < a.hist_cap) a.hist[nhist] = rnorm;
            nhist++;
            nr += k;
            solved = (rnorm <= tol) || (rnorm + 1.0 <= 1.0);
            breakdown = Hbis <= btol;
            const long long cap = inner_itmax < (long long)mem ? inner_itmax : (long long)mem;
            inner_tired = (long long)k >= cap;
            if (!(solved || inner_tired || breakdown)) {
                // v_{k+1} = q / Hbis on own rows (q already sits in V[k]); raw q was published in dst
                inv_h = 1.0 / Hbis;
                for (int row = r0 + tid; row < r1; row += nthr) q[row] *= inv_h;
                cur ^= 1;
            }
        }
        // ---- back substitution R y = z (thread 0), then x += Σ y_i v_i on own rows
        __syncthreads();
        if (tid == 0) {
            bool inc = false;
            for (int i = 0; i < k; ++i) sy[i] = sz[i];
            for (int i = k; i >= 1; --i) {
                int pos = nr + i - k - 1;
                for (int j = k; j > i; --j) {
                    sy[i - 1] -= sR[pos] * sy[j - 1];
                    pos = pos - j + 1;
                }
                if (fabs(sR[pos]) <= btol) { sy[i - 1] = 0.0; inc = true; }
                else sy[i - 1] /= sR[pos];
            }
            s_flags[1] = inc ? 1.0 : 0.0;
        }
        __syncthreads();
        if (s_flags[1] != 0.0) inconsistent = true;
        for (int row = r0 + tid; row < r1; row += nthr) {
            double xr = 0.0;
            for (int i = 0; i < k; ++i) xr = fma(sy[i], V[(size_t)i * n + row], xr);
            x[row] += xr;
        }
        inner_itmax -= k;
        iter += k;
        tired = iter >= itmax;
        // publish x (next pass gathers it) — also keeps smem scalars safe to reset
        if (!(solved || tired || breakdown)) {
            gr.barrier();
            cur = 0;
        }
    }
    if (lead) {
        a.result[0] = (double)iter;
        a.result[1] = solved ? 1.0 : 0.0;
        a.result[2] = inconsistent ? 1.0 : 0.0;
        a.result[3] = breakdown ? 1.0 : 0.0;
        a.result[4] = rnorm;
        a.result[5] = rnorm0;
        a.result[6] = (double)(nhist < a.hist_cap ? nhist : a.hist_cap);
        a.result[7] = gr.bar.aborted() ? 1.0 : 0.0;
    }
}

// =============================================================================================
// host side
// =============================================================================================

static int pow2floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// Lanes per row inside the persistent kernels: the row-length choice, narrowed when the CTA has
// fewer rows than row groups so that all rows of the CTA are processed in one sweep.
static int persistent_tpr(const nupgcm_csr *A) {
    const char *env = getenv("NUPGCM_TPR");
    if (env) {
        int v = atoi(env);
        if (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) return v;
    }
    const int grid = A->ctx->coop_grid;
    const int rows_per_cta = (int)((A->n_rows + grid - 1) / grid);
    int t = A->tpr;
    if (rows_per_cta > 0 && rows_per_cta * t > kThreads && rows_per_cta <= kThreads / 2)
        t = std::max(2, pow2floor(kThreads / rows_per_cta));
    return std::min(t, 32);
}

struct Workspace {
    double *d;
    size_t bytes;
};
static Workspace g_ws = {nullptr, 0};   // grown on demand, one per process (contexts are serialised)

static int32_t ensure_workspace(nupgcm_ctx *ctx, size_t bytes) {
    if (g_ws.bytes >= bytes) return NUPGCM_OK;
    if (g_ws.d) {
        NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(g_ws.d);
        g_ws.d = nullptr;
        g_ws.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&g_ws.d, bytes);
    if (e != cudaSuccess)
        return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "device allocation failed: %s", cudaGetErrorString(e));
    g_ws.bytes = bytes;
    return NUPGCM_OK;
}

template <int T>
static cudaError_t launch(bool gmres, KrylovArgs &args, nupgcm_ctx *ctx) {
    void *params[] = {&args};
    const void *fn = gmres ? (const void *)k_gmres<T> : (const void *)k_cg<T>;
    return cudaLaunchCooperativeKernel(fn, dim3(ctx->coop_grid), dim3(kThreads), params, 0, ctx->stream);
}

static int32_t solve_common(bool gmres, const nupgcm_csr *A, const nupgcm_vec *dinv, double pscale,
                            const nupgcm_vec *y, nupgcm_vec *x, double atol, double rtol,
                            int64_t itmax, int32_t memory, int32_t orth, double *resid_hist,
                            int64_t hist_cap, nupgcm_solve_stats *stats) {
    NUPGCM_REQUIRE(nullptr, A && y && x, "solve: NULL argument");
    nupgcm_ctx *ctx = A->ctx;
    NUPGCM_REQUIRE(ctx, A->n_rows == A->n_cols, "solve: matrix must be square");
    NUPGCM_REQUIRE(ctx, y->n == A->n_rows && x->n == A->n_rows, "solve: vector length mismatch");
    NUPGCM_REQUIRE(ctx, !dinv || dinv->n == A->n_rows, "solve: preconditioner length mismatch");
    NUPGCM_REQUIRE(ctx, itmax >= 0 && atol >= 0.0 && rtol >= 0.0, "solve: negative tolerance or itmax");
    NUPGCM_REQUIRE(ctx, x->d != y->d, "solve: x and y must not alias");
    if (gmres) {
        NUPGCM_REQUIRE(ctx, memory >= 1 && memory <= kMaxMemory, "gmres: memory must be in 1..20");
        NUPGCM_REQUIRE(ctx, orth == NUPGCM_ORTH_MGS || orth == NUPGCM_ORTH_CGS2, "gmres: unknown orth");
    }
    NUPGCM_REQUIRE(ctx, hist_cap >= 0 && (hist_cap == 0 || resid_hist), "solve: hist_cap without buffer");
    const int64_t n = A->n_rows;
    if (n == 0) {
        if (stats) { memset(stats, 0, sizeof(*stats)); stats->solved = 1; }
        return NUPGCM_OK;
    }
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t wbytes = (gmres ? (size_t)(memory + 3) : 3) * (size_t)n * sizeof(double);
    int32_t rc = ensure_workspace(ctx, wbytes);
    if (rc) return rc;
    if (hist_cap > ctx->hist_cap) {
        NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_hist);
        ctx->d_hist = nullptr;
        NUPGCM_CUDA(ctx, cudaMalloc(&ctx->d_hist, (size_t)hist_cap * sizeof(double)));
        ctx->hist_cap = hist_cap;
    }
    KrylovArgs args;
    args.rowptr = A->d_rowptr;
    args.colidx = A->d_colidx;
    args.vals = A->d_vals;
    args.part = A->d_part;
    args.n = (int)n;
    args.dinv = dinv ? dinv->d : nullptr;
    args.pscale = pscale;
    args.b = y->d;
    args.x = x->d;
    args.work = g_ws.d;
    args.atol = atol;
    args.rtol = rtol;
    args.itmax = itmax;
    args.mem = memory;
    args.orth = orth;
    args.barrier = ctx->d_barrier;
    args.partials = ctx->d_partials;
    args.hist = hist_cap > 0 ? ctx->d_hist : nullptr;
    args.hist_cap = hist_cap;
    args.result = ctx->d_scalars;

    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4 * sizeof(unsigned long long), ctx->stream));
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev0, ctx->stream));
    cudaError_t e;
    switch (persistent_tpr(A)) {
        case 32: e = launch<32>(gmres, args, ctx); break;
        case 16: e = launch<16>(gmres, args, ctx); break;
        case 8: e = launch<8>(gmres, args, ctx); break;
        case 4: e = launch<4>(gmres, args, ctx); break;
        default: e = launch<2>(gmres, args, ctx); break;
    }
    NUPGCM_CUDA(ctx, e);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev1, ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double *res = ctx->h_scalars;
    if (res[7] != 0.0)
        return nupgcm_fail(ctx, NUPGCM_ERR_CUDA, "%s",
                           "persistent solver kernel aborted: grid barrier watchdog expired");
    const int64_t hist_len = (int64_t)res[6];
    if (hist_cap > 0 && hist_len > 0)
        NUPGCM_CUDA(ctx, cudaMemcpy(resid_hist, ctx->d_hist, (size_t)hist_len * sizeof(double), cudaMemcpyDeviceToHost));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->niter = (int64_t)res[0];
        stats->solved = res[1] != 0.0;
        stats->inconsistent = res[2] != 0.0;
        stats->breakdown = res[3] != 0.0;
        stats->rnorm = res[4];
        stats->rnorm0 = res[5];
        stats->hist_len = hist_cap > 0 ? hist_len : 0;
        stats->launches = 1;
        float ms = 0.f;
        NUPGCM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->sev0, ctx->sev1));
        stats->device_ms = ms;
    }
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_cg_solve(const nupgcm_csr *A, const nupgcm_vec *dinv, double pscale,
                                   const nupgcm_vec *y, nupgcm_vec *x, double atol, double rtol,
                                   int64_t itmax, double *resid_hist, int64_t hist_cap,
                                   nupgcm_solve_stats *stats) {
    return solve_common(false, A, dinv, pscale, y, x, atol, rtol, itmax, 0, 0, resid_hist, hist_cap, stats);
}

extern "C" int32_t nupgcm_gmres_solve(const nupgcm_csr *A, const nupgcm_vec *dinv, double pscale,
                                      const nupgcm_vec *y, nupgcm_vec *x, double atol, double rtol,
                                      int64_t itmax, int32_t memory, int32_t orth,
                                      double *resid_hist, int64_t hist_cap,
                                      nupgcm_solve_stats *stats) {
    return solve_common(true, A, dinv, pscale, y, x, atol, rtol, itmax, memory, orth, resid_hist, hist_cap, stats);
}
