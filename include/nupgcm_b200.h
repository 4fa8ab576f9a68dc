/*
 * nupgcm_b200.h — C ABI of libnupgcm_b200.so: nuPGCM's per-timestep solve path on NVIDIA B200.
 *
 * This is the drop-in boundary behind nuPGCM's CPU()/GPU() architecture switch
 * (reference src/architectures.jl:4-20, extended today by ext/nuPGCMCUDAExt.jl:24-33 with
 * CUDA.jl arrays).  A host (Julia via ccall — see INTEGRATION.md; Python via ctypes in
 * nupgcm_b200/lib.py) hands over already-assembled, already-permuted operands once, and then
 * drives whole solves: no per-Krylov-iteration call crosses this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative nupgcm_status on failure;
 *     nupgcm_last_error(ctx) (ctx may be NULL for creation failures) gives the message;
 *   - handles are opaque; the library owns all device memory; host pointers are borrowed only
 *     for the duration of a call; calls on one context must be serialised by the caller;
 *   - a call returns after the results named in its out-parameters are valid on the host;
 *   - all floating point is IEEE binary64; indices cross the boundary as int64 (Julia Int) with
 *     an explicit index_base (1 for Julia, 0 for C/Python) and are stored as int32 on the device;
 *   - Krylov non-convergence is NOT an error: status 0 with *solved = 0, exactly as the
 *     reference ignores it (src/iterative_solvers.jl:58-67);
 *   - there is no CPU fallback: without a usable sm_100 device nupgcm_create fails.
 */
#ifndef NUPGCM_B200_H
#define NUPGCM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUPGCM_B200_VERSION 100 /* 0.1.0 */

typedef enum {
    NUPGCM_OK = 0,
    NUPGCM_ERR_INVALID = -1,   /* bad argument (null handle, size mismatch, bad enum ...) */
    NUPGCM_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    NUPGCM_ERR_NO_DEVICE = -3, /* no CUDA device / not an sm_100 device */
    NUPGCM_ERR_COMM = -4,      /* a sharded (multi-GPU) solve aborted: a peer did not answer within the watchdog */
    NUPGCM_ERR_ALLOC = -5      /* host or device allocation failed */
} nupgcm_status;

typedef struct nupgcm_ctx nupgcm_ctx;
typedef struct nupgcm_vec nupgcm_vec;
typedef struct nupgcm_index nupgcm_index;
typedef struct nupgcm_csr nupgcm_csr;
typedef struct nupgcm_mesh nupgcm_mesh;
typedef struct nupgcm_comm nupgcm_comm;
typedef struct nupgcm_blockprec nupgcm_blockprec;

int32_t nupgcm_version(void);
const char *nupgcm_last_error(const nupgcm_ctx *ctx);

/* ---- context ---------------------------------------------------------------------------
 * Replaces what `using CUDA` + GPU() give the reference (ext/nuPGCMCUDAExt.jl:8-16 device
 * listing, :33 print_memory_status -> CUDA.pool_status()). */
int32_t nupgcm_create(int32_t device, nupgcm_ctx **out);
int32_t nupgcm_destroy(nupgcm_ctx *ctx);
int32_t nupgcm_synchronize(nupgcm_ctx *ctx);
int32_t nupgcm_mem_status(nupgcm_ctx *ctx, size_t *free_bytes, size_t *total_bytes);
/* name: buffer of >= 64 bytes (may be NULL) */
int32_t nupgcm_device_info(nupgcm_ctx *ctx, int32_t *sm_count, int32_t *cc_major,
                           int32_t *cc_minor, char *name);
/* CUDA-event timing of the library's stream (for bench.py): *ms = elapsed between the two marks */
int32_t nupgcm_timer_start(nupgcm_ctx *ctx);
int32_t nupgcm_timer_stop(nupgcm_ctx *ctx, float *ms);
/* number of kernels this context has launched since creation */
int32_t nupgcm_launch_count(nupgcm_ctx *ctx, int64_t *count);

/* page-locked host buffers (cudaHostAlloc) for the host's copy of the state: the per-step
 * upload/download of src/model.jl:275,282,312 then runs as plain DMA */
int32_t nupgcm_host_alloc(nupgcm_ctx *ctx, int64_t bytes, void **out);
int32_t nupgcm_host_free(nupgcm_ctx *ctx, void *p);
/* Restrict the persistent solver kernels of this context to `grid` CTAs (1..SM count; default:
 * one per SM).  Lets several contexts share one device, each with a slice of the SMs. */
int32_t nupgcm_set_grid(nupgcm_ctx *ctx, int32_t grid);

/* ---- multi-GPU: row-block sharded solves -------------------------------------------------
 * The reference runs on one GPU; BASELINE.json's north_star shards the Krylov solves over the
 * GPUs of one NVLink/NVSwitch box.  One rank per GPU (normally one process per rank, launched by
 * torchrun).  Every rank builds the same operands; nupgcm_csr_shard then makes the persistent
 * solvers on that matrix collective: rank r's CTAs own the r-th contiguous block of rows of the
 * (internally RCM-ordered) system, push the rows a peer's SpMV gathers straight into the peer's
 * memory over NVLink, and combine dot products through flagged words in peer memory — no host
 * round trip and no separate collective launch per iteration.  Right-hand side and solution
 * vectors stay full-length and replicated: each rank reads its rows of y and the warm start, and
 * the solve ends with an all-gather so x is complete on every rank.
 *
 * Set-up protocol: nupgcm_comm_create on every rank; exchange the 64-byte handles from
 * nupgcm_comm_ipc_handle with any host transport (torch.distributed / MPI all-gather) and pass
 * all of them, ordered by rank, to nupgcm_comm_connect_ipc; or, for ranks living in one process,
 * pass the communicator objects to nupgcm_comm_connect_local.  Sharded solves must be called by
 * all ranks with identical arguments, in the same order. */
#define NUPGCM_MAX_RANKS 8
#define NUPGCM_IPC_HANDLE_BYTES 64
/* max_n: largest system (rows) that will be solved over this communicator */
int32_t nupgcm_comm_create(nupgcm_ctx *ctx, int32_t rank, int32_t nranks, int64_t max_n,
                           nupgcm_comm **out);
int32_t nupgcm_comm_destroy(nupgcm_comm *comm);
int32_t nupgcm_comm_ipc_handle(nupgcm_comm *comm, void *handle_out /* 64 bytes */);
int32_t nupgcm_comm_connect_ipc(nupgcm_comm *comm, const void *handles /* nranks x 64 bytes */);
int32_t nupgcm_comm_connect_local(nupgcm_comm *comm, nupgcm_comm *const *all /* nranks */);
/* make the persistent solvers on A collective over comm (NULL: back to single-GPU) */
int32_t nupgcm_csr_shard(nupgcm_csr *A, nupgcm_comm *comm);
/* rows [*row_begin, *row_end) of the internally ordered system owned by `rank`, and the number
 * of halo rows it receives per SpMV (valid after the first sharded solve or nupgcm_csr_shard) */
int32_t nupgcm_csr_shard_info(nupgcm_csr *A, int32_t rank, int64_t *row_begin, int64_t *row_end,
                              int64_t *nnz_owned, int64_t *halo_rows);

/* Host-only (no context): the layout a sharded solve uses for a square matrix over `nranks`
 * ranks of `grid_per_rank` CTAs.  perm_out[n]: internal row i = caller row perm_out[i];
 * row_begin[nranks+1]: rank r owns internal rows [row_begin[r], row_begin[r+1]);
 * halo_lo/hi[dst*nranks + src]: range of src's rows pushed to dst before each SpMV (0,0: none).
 * Output pointers may be NULL. */
int32_t nupgcm_shard_plan(int64_t n, const int64_t *rowptr, const int64_t *colidx, int32_t index_base,
                          int32_t nranks, int32_t grid_per_rank, int64_t *perm_out,
                          int64_t *row_begin, int64_t *halo_lo, int64_t *halo_hi);

/* ---- vectors (replace CuVector{Float64}: ext/nuPGCMCUDAExt.jl:24-26,32) ------------------ */
int32_t nupgcm_vec_create(nupgcm_ctx *ctx, int64_t n, nupgcm_vec **out); /* zero-filled */
int32_t nupgcm_vec_destroy(nupgcm_vec *v);
int32_t nupgcm_vec_size(const nupgcm_vec *v, int64_t *n);
int32_t nupgcm_vec_upload(nupgcm_vec *v, const double *host, int64_t n);   /* on_architecture(GPU(), a) */
int32_t nupgcm_vec_download(const nupgcm_vec *v, double *host, int64_t n); /* on_architecture(CPU(), a) */
int32_t nupgcm_vec_fill(nupgcm_vec *v, double value);
int32_t nupgcm_vec_copy(nupgcm_vec *dst, const nupgcm_vec *src);
/* y = alpha*x + beta*y (kaxpby!) */
int32_t nupgcm_vec_axpby(nupgcm_vec *y, double alpha, const nupgcm_vec *x, double beta);
int32_t nupgcm_vec_dot(const nupgcm_vec *x, const nupgcm_vec *y, double *out);
int32_t nupgcm_vec_norm2(const nupgcm_vec *x, double *out);
/* max |x_i| over the first `count` entries (count <= 0: all) and NaN flag: the blow-up check
 * of src/model.jl:149-153 (count = nu restricts the [u; p] solution vector to the velocity) */
int32_t nupgcm_vec_maxabs(const nupgcm_vec *x, int64_t count, double *maxabs, int32_t *has_nan);
/* z = d .* r — mul!(z, ::Diagonal, r) of src/inversion.jl:54, src/evolution.jl:149,167 */
int32_t nupgcm_diag_apply(nupgcm_vec *z, const nupgcm_vec *d, const nupgcm_vec *r);

/* device-resident index vector; replaces the per-step upload of `inv_perm` in
 * `solver.x[inv_perm]` (src/model.jl:282,312) */
int32_t nupgcm_index_create(nupgcm_ctx *ctx, const int64_t *idx, int64_t n, int32_t index_base,
                            nupgcm_index **out);
int32_t nupgcm_index_destroy(nupgcm_index *ix);
/* dst[i] = src[idx[i]], i < n(idx); dst may be longer than the index (tail untouched) */
int32_t nupgcm_vec_gather(nupgcm_vec *dst, const nupgcm_vec *src, const nupgcm_index *idx);

/* ---- CSR matrices (replace CuSparseMatrixCSR: ext/nuPGCMCUDAExt.jl:27) -------------------
 * rowptr has n_rows+1 entries.  drop_zeros != 0 removes explicitly stored zeros (32 % of the
 * Gridap-assembled inversion matrix) from the device copy; update_values still takes the
 * original nnz values in the original order. */
int32_t nupgcm_csr_create(nupgcm_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                          const int64_t *rowptr, const int64_t *colidx, const double *vals,
                          int32_t index_base, int32_t drop_zeros, nupgcm_csr **out);
int32_t nupgcm_csr_destroy(nupgcm_csr *A);
int32_t nupgcm_csr_info(const nupgcm_csr *A, int64_t *n_rows, int64_t *n_cols,
                        int64_t *nnz_given, int64_t *nnz_stored);
int32_t nupgcm_csr_update_values(nupgcm_csr *A, const double *vals, int64_t nnz);
/* out = M + theta*(Kh + Kv) on four matrices created from the same pattern with the same
 * drop_zeros=0 setting — collect_evolution_LHS (src/evolution.jl:143-177, src/model.jl:251-261) */
int32_t nupgcm_csr_combine(nupgcm_csr *out, const nupgcm_csr *M, const nupgcm_csr *Kh,
                           const nupgcm_csr *Kv, double theta);
/* dinv = 1 ./ diag(A) — the Jacobi preconditioner of src/evolution.jl:149,167 */
int32_t nupgcm_csr_inv_diag(const nupgcm_csr *A, nupgcm_vec *dinv);
/* y = alpha*A*x + beta*y — mul!(y, A, x) and `B*b .+ b0` of src/inversion.jl:104 */
int32_t nupgcm_spmv(const nupgcm_csr *A, const nupgcm_vec *x, nupgcm_vec *y, double alpha,
                    double beta);

/* Reverse Cuthill-McKee ordering of the symmetrised pattern of a square CSR matrix (host-only;
 * needs no context).  perm_out[i] = row placed at position i.  The persistent solvers apply this
 * ordering internally (the caller never sees it); it is exported for hosts that want the same
 * ordering the reference gets from CuthillMcKee.symrcm (src/dofs.jl:98-100). */
int32_t nupgcm_rcm_order(int64_t n, const int64_t *rowptr, const int64_t *colidx,
                         int32_t index_base, int64_t *perm_out);

/* ---- Krylov solvers (replace Krylov.krylov_solve! at src/iterative_solvers.jl:58) ---------
 * x is in/out: its content on entry is the warm start (the reference aliases x to workspace.x,
 * src/iterative_solvers.jl:26-29).  itmax == 0 means 2n (Krylov.jl default).  The stopping
 * measure is the preconditioned one, compared with atol + rtol*(initial measure).
 * Preconditioner: dinv != NULL -> M = diag(dinv); else M = pscale*I (pass 1.0 for none).
 * resid_hist (may be NULL) receives up to hist_cap measures (initial one first); *hist_len is
 * how many were written.  Out-pointers may be NULL. */
typedef struct {
    int64_t niter;
    int32_t solved;
    int32_t inconsistent;
    int32_t breakdown;     /* GMRES only */
    int32_t reserved;
    double  rnorm;         /* last value of the stopping measure */
    double  rnorm0;        /* initial value of the stopping measure */
    float   device_ms;     /* CUDA-event duration of the solve kernel(s) */
    int32_t launches;      /* kernels launched by this call */
    int64_t hist_len;
    float   phase_frac[4]; /* GMRES: share of kernel time in SpMV / local vector work / waiting for
                              grid reductions / scalar recurrences (CTA 0's clock) */
    float   sm_mhz;        /* GMRES: SM clock the kernel actually ran at (clock64 / globaltimer) */
    float   reserved2;
} nupgcm_solve_stats;

/* sizeof(nupgcm_solve_stats) as this library was built: a host binding asserts it against its own
 * mirror of the struct at load time (a shorter mirror would be overrun: the library writes the whole
 * struct).  nupgcm_version() / 100 is the ABI major version. */
int64_t nupgcm_solve_stats_size(void);

/* CG with Jacobi/scalar left preconditioner (CgWorkspace, src/evolution.jl:118-126) */
int32_t nupgcm_cg_solve(const nupgcm_csr *A, const nupgcm_vec *dinv, double pscale,
                        const nupgcm_vec *y, nupgcm_vec *x, double atol, double rtol,
                        int64_t itmax, double *resid_hist, int64_t hist_cap,
                        nupgcm_solve_stats *stats);

/* orthogonalisation variants of the Arnoldi step */
#define NUPGCM_ORTH_MGS 0  /* modified Gram-Schmidt, what Krylov.jl does: parity default */
#define NUPGCM_ORTH_CGS2 1 /* classical Gram-Schmidt twice: 3 grid reductions per iteration */
#define NUPGCM_ORTH_CGS2_FUSED 2 /* CGS2 with 2 reductions: norm by Pythagoras from the second pass,
                                    next SpMV on the exchanged q1 corrected with the Arnoldi relation */

/* restarted GMRES(memory), left preconditioned (GmresWorkspace, src/inversion.jl:74-94) */
int32_t nupgcm_gmres_solve(const nupgcm_csr *A, const nupgcm_vec *dinv, double pscale,
                           const nupgcm_vec *y, nupgcm_vec *x, double atol, double rtol,
                           int64_t itmax, int32_t memory, int32_t orth, double *resid_hist,
                           int64_t hist_cap, nupgcm_solve_stats *stats);

/* ---- operator preconditioners (src/preconditioners.jl) -------------------------------------
 * BlockDiagonalPreconditioner of the inversion system [u; p] (:53-125): y[0:n1] = P⁻¹ x[0:n1],
 * y[n1:] = T⁻¹ x[n1:], each inverse a CgPreconditioner (:5-37): CG on the block with a Jacobi
 * preconditioner, atol = rtol = sqrt(eps), at most `itmax` iterations (0: 2n), warm-started from
 * the block's previous answer.  The reference's GPU set-up uses ILU(0) for P (:102-107); Jacobi
 * here (declared deviation).  The handles passed in must outlive the preconditioner. */
int32_t nupgcm_blockprec_create(nupgcm_ctx *ctx, const nupgcm_csr *P, const nupgcm_vec *P_dinv,
                                int64_t P_itmax, const nupgcm_csr *T, const nupgcm_vec *T_dinv,
                                int64_t T_itmax, nupgcm_blockprec **out);
int32_t nupgcm_blockprec_destroy(nupgcm_blockprec *M);
int32_t nupgcm_blockprec_apply(nupgcm_blockprec *M, const nupgcm_vec *x, nupgcm_vec *y); /* mul!(y, M, x) */
int32_t nupgcm_blockprec_info(const nupgcm_blockprec *M, int64_t *applies, int64_t *inner_iters);
/* GMRES(memory), left-preconditioned by M, modified Gram-Schmidt: same contract as
 * nupgcm_gmres_solve (x in/out = warm start; non-convergence is not an error) */
int32_t nupgcm_gmres_solve_prec(const nupgcm_csr *A, nupgcm_blockprec *M, const nupgcm_vec *y,
                                nupgcm_vec *x, double atol, double rtol, int64_t itmax,
                                int32_t memory, double *resid_hist, int64_t hist_cap,
                                nupgcm_solve_stats *stats);

/* diagnostics, host only (no device needed): y = A x computed by walking the streaming-SpMV tables
 * (tiles, footprints staged in an arena of `arena` entries, per-warp streams of jagged-diagonal slices)
 * that nupgcm_csr_prepare builds for `grid` CTAs, exactly as the persistent kernels walk them; 0-based
 * CSR as given.  Fails if any row is not produced exactly once or an arena slot is reused too early.
 * gather_wavefronts / positions: shared-memory wavefronts the vector gathers need (2 per position when
 * free of bank conflicts).  Used by the CPU tests of the table builder. */
int32_t nupgcm_diag_stream_spmv_host(int64_t n, const int64_t *rowptr, const int64_t *colidx,
                                     const double *vals, const double *x, int32_t grid,
                                     int32_t arena, double *y, int64_t *n_tiles, int64_t *n_entries,
                                     int64_t *gather_wavefronts, int64_t *positions);

/* diagnostics: y = A x computed `reps` times by the STREAMING SpMV engine of the persistent solvers
 * alone (same tables, CTAs and warp roles, no grid-wide wait); reports the average device time of one
 * product.  mode 0 = the real product; 1 = pieces pulled through the rings untouched (copy pipeline
 * only), 2 = no footprint gather (timing experiments; y is then meaningless), 3 = the real product with
 * per-warp activity clocks: warp_cycles[148 CTAs][11 warps][8] receives SM cycles per product spent
 * waiting for the footprint / for ring pieces / in table loads and bookkeeping / in full positions / in
 * the jagged ends / in the row callback / at the closing CTA barrier.  Fails when the matrix is
 * SM-resident (no streaming tables). */
int32_t nupgcm_diag_stream_spmv(nupgcm_csr *A, const nupgcm_vec *x, nupgcm_vec *y, int32_t reps,
                                int32_t mode, float *us_per_spmv, double *warp_cycles);

/* diagnostics: GB/s at which cp.async.bulk alone pulls `total_bytes` of HBM into shared memory when
 * every one of `warps` warps per CTA (one CTA per SM) keeps `slots` copies of `piece` bytes in flight. */
int32_t nupgcm_diag_tma_stream(nupgcm_ctx *ctx, int64_t total_bytes, int32_t piece, int32_t slots,
                               int32_t warps, int32_t reps, float *gb_per_s);

/* diagnostics: SM cycles per dependent operation — out8 = {LDS.64 chase, DFMA chain, LDS.U16->LDS.64 pair,
 * IMAD chain, DFMA with 8 independent chains (1 warp), same with 11 warps, one block of 8 streaming-SpMV
 * slice positions (1 warp), same with 11 warps} */
int32_t nupgcm_diag_latency(nupgcm_ctx *ctx, double *out8);

/* diagnostics: average latency (µs) of the grid-wide reduction the persistent solvers use.
 * mode 0: flagged-slot exchange only; 1: + block reduction; 2: + release/acquire fences. */
int32_t nupgcm_diag_reduce_latency(nupgcm_ctx *ctx, int32_t mode, int32_t reps, int32_t grid,
                                   int32_t threads, float *us_per_reduction);

/* diagnostics: round-trip time (µs) of a flag ping-pong between CTA 0 and CTA `peer` through L2;
 * variant = 10*store_kind + load_kind (see csrc/krylov.cu). */
int32_t nupgcm_diag_pingpong(nupgcm_ctx *ctx, int32_t peer, int32_t variant, int32_t reps,
                             float *us_round_trip);

/* diagnostics: one-way latency (µs) of a flag crossing NVLink between two ranks' arenas
 * (variant 0 relaxed.sys, 1 release/acquire.sys, 2 fence.sys + relaxed).  Call on every rank. */
int32_t nupgcm_diag_xping(nupgcm_comm *comm, int32_t rank_a, int32_t rank_b, int32_t variant,
                          int32_t reps, float *us_one_way);

/* diagnostics: average latency (µs) of the sharded solvers' reduction of `count` values
 * (publish != 0: the variant that also publishes pushed rows).  Collective over all ranks. */
int32_t nupgcm_diag_xreduce(nupgcm_comm *comm, int32_t count, int32_t publish, int32_t reps,
                            float *us_per_reduction);

/* ---- per-step element right-hand side (replaces the CPU Gridap assemble_vector of
 *      src/model.jl:269-275 and the broadcast of :278) -----------------------------------
 * P2 tetrahedra (n_loc = 10) or P2 triangles (n_loc = 6).
 *   cell_b[c*n_loc + i]            index of local buoyancy DOF i of cell c in the extended vector
 *                                  [free values in solver order (nb) ; Dirichlet values (nbd)]
 *   cell_u[(c*n_loc + i)*3 + comp] same for the velocity [free (nu) ; Dirichlet (nud)]
 *   grad[(c*(d+1) + k)*3 + j]      d/dx_j of barycentric coordinate k, vol[c] the cell measure
 *   bary[q*(d+1) + k], w[q]        quadrature rule (weights sum to 1)
 * Indices are 0-based int32. */
int32_t nupgcm_mesh_create(nupgcm_ctx *ctx, int64_t n_cells, int32_t n_loc,
                           const int32_t *cell_b, const int32_t *cell_u, const double *grad,
                           const double *vol, int32_t nq, const double *bary, const double *w,
                           int64_t nb, const double *b_dirichlet, int64_t nbd, int64_t nu,
                           const double *u_dirichlet, int64_t nud, nupgcm_mesh **out);
/* Same with first-order buoyancy (reference src/spaces.jl:31-39 `Spaces(...; b_order=1)`, the production
 * set-up of scratch/run.jl:152): the velocity stays P2 (n_loc = 10 / 6 nodes in cell_u) while cell_b holds
 * n_loc_b buoyancy DOFs per cell — n_loc (P2) or the vertex count 4 / 3 (P1: cell_b[c*n_loc_b + i]).
 * All kernels of this handle (rhs_adv, rebuild_kv, rebuild_friction) follow the buoyancy order. */
int32_t nupgcm_mesh_create_orders(nupgcm_ctx *ctx, int64_t n_cells, int32_t n_loc, int32_t n_loc_b,
                                  const int32_t *cell_b, const int32_t *cell_u, const double *grad,
                                  const double *vol, int32_t nq, const double *bary, const double *w,
                                  int64_t nb, const double *b_dirichlet, int64_t nbd, int64_t nu,
                                  const double *u_dirichlet, int64_t nud, nupgcm_mesh **out);
int32_t nupgcm_mesh_destroy(nupgcm_mesh *m);
/* scheme 1: ∫(b − Δt(u·∇b + w N²)) d   (src/model.jl:292-295)
 * scheme 2: ∫(4/3 b − 1/3 b⁻ − 2/3 Δt((2u−u⁻)·∇(2b−b⁻) + (2w−w⁻) N²)) d   (:297-300)
 * b, b_prev: nb; u, u_prev: vectors whose first nu entries are the velocity in solver order
 * (the inversion solution vector [u; p] can be passed directly); out: nb. */
int32_t nupgcm_rhs_adv(nupgcm_mesh *m, int32_t scheme, double dt, double N2, const nupgcm_vec *b,
                       const nupgcm_vec *b_prev, const nupgcm_vec *u, const nupgcm_vec *u_prev,
                       nupgcm_vec *out);
/* Adaptive timestep, update_Δt! of src/timesteppers.jl:108-119 (production runs use
 * BDF1(adaptive=true), scratch/run.jl:163):
 *   *dt_out = cfl_factor * min over cells of h_cells[c] / max(max_q |u(x_q)|, u_min)
 * with the quadrature points of the mesh handle and h_cells = compute_h_cells (longest edge per
 * cell, src/meshes.jl:127-134) set once with nupgcm_mesh_set_cell_sizes.  u: vector whose first nu
 * entries are the velocity in solver order. */
int32_t nupgcm_mesh_set_cell_sizes(nupgcm_mesh *m, const double *h_cells, int64_t n_cells);
int32_t nupgcm_cfl_dt(nupgcm_mesh *m, const nupgcm_vec *u, double cfl_factor, double u_min,
                      double *dt_out);
/* Convection parameterisation (src/model.jl:229-246, src/inputs.jl:87-91): per-step rebuild of
 *   Kv = ∫ κᵥ ∂z b ∂z d,  rhs_v = ∫ κᵥ ∂z b_diri ∂z d,  rhs_diff = ∫ −N² κᵥ ∂z d
 * with κᵥ(x_q) = kv_q[c*nq+q] + kappa_c (1 + tanh(−alpha (N2 + ∂z b)(x_q) / N2min)) / 2, replacing
 * the CPU build_Kᵥ / build_rhs_diff + permutation + upload.  `pattern` / `Kv`: nb x nb matrices in
 * solver order created with drop_zeros = 0 from the evolution pattern (M, Kh, Kv share it). */
int32_t nupgcm_mesh_enable_kv_rebuild(nupgcm_mesh *m, const nupgcm_csr *pattern, const double *kv_q);
int32_t nupgcm_rebuild_kv(nupgcm_mesh *m, double alpha, double N2, double kappa_c, double N2min,
                          const nupgcm_vec *b, nupgcm_csr *Kv, nupgcm_vec *rhs_v, nupgcm_vec *rhs_diff);
/* Eddy parameterisation (src/model.jl:160-170, src/inputs.jl:130-137): rebuild of the inversion
 * matrix with ν(x_q) = LogSumExp_smoothing(nu_min, f_q² / sqrt(N2min² + (alpha (N2 + ∂z b))²)):
 *   A.vals = A0_vals + assembled ∫ 2 a2e2 ν σ(u)⊙σ(v)            (src/inversion.jl:172-182)
 * A: N x N inversion matrix in solver order, created with drop_zeros = 0; A0_vals: its
 * frictionless part (pressure gradient, divergence, Coriolis) on the same pattern; f_q[c*nq+q]. */
int32_t nupgcm_mesh_enable_nu_rebuild(nupgcm_mesh *m, const nupgcm_csr *A, const double *A0_vals,
                                      const double *f_q);
int32_t nupgcm_rebuild_friction(nupgcm_mesh *m, double a2e2, double alpha, double N2, double N2min,
                                  double smoothing, double nu_min, const nupgcm_vec *b, nupgcm_csr *A);
/* out = rhs_adv + theta*rhs_diff + dt*rhs_flux − (rhs_m + theta*(rhs_h + rhs_v))  (src/model.jl:278) */
int32_t nupgcm_rhs_combine(nupgcm_vec *out, const nupgcm_vec *rhs_adv, double theta, double dt,
                           const nupgcm_vec *rhs_diff, const nupgcm_vec *rhs_flux,
                           const nupgcm_vec *rhs_m, const nupgcm_vec *rhs_h,
                           const nupgcm_vec *rhs_v);

#ifdef __cplusplus
}
#endif
#endif /* NUPGCM_B200_H */
