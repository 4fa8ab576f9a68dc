// Shared declarations of libnupgcm_b200: handle layouts, error plumbing, device-side helpers.
// Built for sm_100a only (B200); there is deliberately no other code path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nupgcm_b200.h"

struct NupgcmTileDesc;
struct NupgcmWarpDesc;
struct NupgcmSlice;

// -------------------------------------------------------------------------------------------
// host-side handle layouts
// -------------------------------------------------------------------------------------------

struct nupgcm_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    char name[64];
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;          // nupgcm_timer_*
    cudaEvent_t sev0, sev1;        // per-solve timing
    int64_t launches;
    // persistent-kernel scratch: grid barrier word, reduction slots, result mailbox
    unsigned long long *d_barrier; // [4]
    double *d_partials;            // [kPartialSlots * grid]
    double *d_scalars;             // [64] device scalars
    double *h_scalars;             // pinned mirror
    double *d_hist;                // residual history
    int64_t hist_cap;
    int coop_grid;                 // CTAs of the persistent solver kernels
    double *d_ws;                  // workspace of the persistent solvers (grown on demand)
    size_t ws_bytes;
    char err[512];
};

// Communicator of the sharded solves: one per rank (= one GPU, normally one process).  Each rank
// owns an "arena" in its device memory that every other rank can write into (CUDA IPC mapping
// between processes, plain pointers inside one process):
//   [0, 128)                    abort word (sticky: a rank whose watchdog fired raises it everywhere)
//   [kArenaXrOffset, +24 KB)    inter-rank reduction slots  LLSlot[2 banks][kPartialSlots][kXRep][kMaxRanks]
//   [kArenaVecOffset, ...)      exchange vectors, 3 x n_pad doubles: the rows of the multiplied
//                               vector / iterate that the rank's SpMV gathers (own rows + halo),
//                               followed by 3 x n_pad flagged 16-byte words: the "mailbox" form of
//                               the same vectors, into which peers push halo rows (value + tag in
//                               one word, so no fence is needed; the reader unpacks them into the
//                               plain vectors before its SpMV)
static const int kMaxRanks = NUPGCM_MAX_RANKS;
static const int kXRep = 4;                  // replicas of the all-reader inter-rank slots
static const size_t kArenaXrOffset = 128, kArenaVecOffset = 32768;

struct nupgcm_comm {
    nupgcm_ctx *ctx;
    int rank, nranks;
    int64_t max_n, n_pad;
    char *arena;                   // this rank's arena
    size_t arena_bytes;
    char *peer[kMaxRanks];         // every rank's arena as mapped here (peer[rank] == arena)
    int ipc_mapped[kMaxRanks];
    int connected;
    int broken;                    // a sharded solve aborted: the sequence numbers are undefined
    unsigned xgen;                 // sequence number of the inter-rank reductions (never reset)
    unsigned long long ping_seq;   // nupgcm_diag_xping
};

struct nupgcm_vec {
    nupgcm_ctx *ctx;
    int64_t n;
    double *d;
};

struct nupgcm_index {
    nupgcm_ctx *ctx;
    int64_t n;
    int64_t max_value;             // largest index held (bounds check at gather time)
    int32_t *d;
};

struct nupgcm_csr {
    nupgcm_ctx *ctx;
    int64_t n_rows, n_cols;
    int64_t nnz_given;             // entries handed over by the host
    int64_t nnz;                   // entries stored on the device (after optional zero drop)
    int32_t *d_rowptr;             // [n_rows+1]
    int32_t *d_colidx;             // [nnz]
    double *d_vals;                // [nnz]
    int32_t *d_keep;               // [nnz] position in the host value array (only when dropping)
    double *d_stage;               // staging for update_values when dropping
    int32_t *d_part;               // [grid+1] row ranges of the persistent kernels
    int tpr;                       // threads per row of the SpMV variant chosen from row lengths
    int dropped;
    // SM-resident form for the persistent solvers: per CTA, the sorted list of distinct columns
    // its rows touch ("footprint") and, per stored entry, the 16-bit position of its column in
    // that list.  With it a CTA keeps its matrix slice in shared memory for the whole solve.
    uint16_t *d_loc;               // [nnz] local column index
    int32_t *d_foot_ptr;           // [coop_grid+1]
    int32_t *d_foot;               // [sum of footprints] global column ids
    int res_max_nnz, res_max_foot, res_max_rows;   // maxima over CTAs (0 when unavailable)
    int prepared_grid;             // grid the partition / resident tables were built for
    int32_t *h_rowptr, *h_col;     // host copies of the structure
    // internally reordered copy used by the persistent solvers (RCM of the whole pattern)
    int32_t *d_perm;               // [n] internal row i  <->  caller row perm[i]
    int32_t *d_prow, *d_pcol;      // reordered CSR structure (columns are internal ids)
    int32_t *d_psrc;               // [nnz] position of each reordered entry in d_vals
    double *d_pvals;               // [nnz] reordered values, refreshed when vals_version moves
    int32_t *h_prow, *h_pcol;
    long long vals_version, pvals_version;
    // streaming form (matrix slice larger than shared memory): tiled, per-warp entry streams built by
    // nupgcm_csr_prepare for the current grid — see "streaming SpMV tables" below
    double *d_svals;               // [stream entries] values in stream order (refreshed with pvals)
    uint16_t *d_scols;             // [stream entries] column = position in the tile's footprint
    int32_t *d_ssrc;               // [stream entries] position in d_pvals, -1 for alignment padding
    int64_t stream_entries;
    NupgcmTileDesc *d_tiles;       // tiles of this rank's CTAs
    int32_t *d_tile_ptr;           // [grid_per_rank+1]
    NupgcmWarpDesc *d_wdesc;       // [grid_per_rank][kMainWarps]
    NupgcmSlice *d_slices;         // slice tables of all warps
    uint8_t *d_twcnt;              // [tiles][kMainWarps] slices of each warp in each tile
    int32_t *d_srow;               // row tables: internal row ids ...
    int32_t *d_slen;               //             ... and row lengths, in slice order
    int32_t *d_sfoot;              // footprints of all tiles (internal column ids, sorted per tile)
    int str_T, str_fmax, str_max_rows;   // str_fmax: largest tile footprint of the tables (0: no tables); largest row block
    long long svals_version;
    // sharded solves: the communicator, and for every peer the range of THIS rank's rows that the
    // peer's SpMV gathers (bounding range of the peer's column footprint inside this rank's block)
    nupgcm_comm *comm;
    int prepared_ranks;
    int push_lo[kMaxRanks], push_hi[kMaxRanks];
    int32_t *d_halo_ptr;           // [grid+1] per CTA of this rank: its columns owned by other ranks
    int32_t *d_halo_idx;
    int64_t halo_total;            // distinct halo columns of the whole rank
};

int32_t nupgcm_csr_prepare(nupgcm_csr *A, int grid);
int32_t nupgcm_reserve_solver_workspace(nupgcm_ctx *ctx, int64_t max_n);

struct nupgcm_mesh {
    nupgcm_ctx *ctx;
    int64_t n_cells;
    int n_loc, n_vert, nq;         // n_loc: local velocity nodes (P2)
    int n_loc_b;                   // local buoyancy DOFs: n_loc (P2) or n_vert (P1)
    int64_t nb, nbd, nu, nud;
    int32_t *d_cell_b;             // [n_loc_b][n_cells] (transposed: coalesced over cells)
    int32_t *d_cell_u;             // [n_loc*3][n_cells]
    double *d_grad;                // [n_vert*3][n_cells]
    double *d_vol;                 // [n_cells]
    double *d_bdir, *d_udir;       // Dirichlet values
    double *d_elem;                // [n_loc_b][n_cells] elemental vectors
    int32_t *d_gptr;               // [nb+1] gather lists: per free DOF, the elemental slots
    int32_t *d_gidx;               //        sorted by cell id (deterministic summation order)
    double *d_phi;                 // [nq][n_loc] basis values
    double *d_dphi;                // [nq][n_loc][n_vert] barycentric derivatives
    double *d_w;                   // [nq]
    double *d_hcells;              // [n_cells] longest edge per cell (adaptive Δt)
    unsigned long long *d_minbits; // scratch of the CFL minimum
    // Kᵥ rebuild (convection parameterisation): per stored matrix entry, its element-matrix slots
    int32_t *d_kptr, *d_kidx;
    double *d_kvq;                 // [n_cells][nq] base κᵥ at the quadrature points
    double *d_emat, *d_evec;       // [n_loc_b*n_loc_b][n_cells], 2 x [n_loc_b][n_cells]
    int64_t kv_nnz;
    // friction-block rebuild (eddy parameterisation)
    int32_t *d_nptr, *d_nidx;
    double *d_fq;                  // [n_cells][nq] Coriolis parameter of the parameterisation
    double *d_nmat;                // [(n_loc*3)^2][n_cells] element blocks
    double *d_A0;                  // [nnz] frictionless part of the inversion matrix
    int64_t nu_nnz;
};

// ---- streaming SpMV tables ------------------------------------------------------------------
// A CTA whose matrix slice does not fit in shared memory walks its rows tile by tile.  A tile is a
// run of kTileRows consecutive rows — one slice of 32 rows for each of the kMainWarps solver warps.
// The entries of the multiplied vector a tile touches (its "footprint", the distinct columns of its
// rows) are staged in a shared-memory arena by the comm warps, several tiles ahead of the solver
// warps, and every matrix entry carries a 16-bit BYTE offset into its tile's staged footprint instead
// of a 32-bit column.  Inside a tile the rows are sorted by decreasing length and cut into slices of
// 32 rows, one row per lane, stored in jagged-diagonal order (position j of every row that has one,
// rows in slice order, then position j+1, ...): consecutive lanes read consecutive addresses, no
// padding, no cross-lane reduction, no divergence beyond the last few positions.  WHICH entry of a row
// sits at which position is free, and chosen so that the 16 lanes of a half-warp hit 16 different
// shared-memory banks when they gather the vector.  Rows longer than kLongRow are kept out of the slices
// (whole warp per row, entries contiguous).  The items of a tile — slices and long rows — are dealt to the
// warps heaviest item to lightest warp (d_twcnt says how many each warp gets), so every warp owns ONE
// contiguous entry stream per CTA — values and 16-bit offsets — which it pulls through a private
// shared-memory ring with TMA bulk copies of kPieceEntries entries.
static const int kMainWarps = 11;      // solver warps of the persistent kernels (krylov.cu)
static const int kPieceEntries = 512;  // entries per TMA piece: 4 KB of values + 1 KB of offsets (sized from
                                       // profiles/tma_piece_size_r02.txt: >= 2 KB copies reach the HBM rate)
static const int kRingPieces = 2;      // pieces per warp ring: one being consumed, one in flight.  11 warps x 5 KB in flight
                                       // saturate HBM (same profile); deeper rings only queue more bytes in front of the
                                       // comm warps' footprint gathers (3 pieces: 3.7 us per L2 round trip under load)
static const int kTileRows = 32 * kMainWarps;   // rows per tile (times a multiplier when a CTA would exceed kMaxTiles)
static const int kMaxTiles = 64;       // tiles per CTA: one pair of mbarriers each
static const int kArenaEntries = 13312; // staged footprint entries in flight (104 KB next to 110 KB of rings)

struct NupgcmTileDesc { int32_t row0, nrows, foot_off, foot_len, xs_off, dep, pad0, pad1; };   // xs_off: arena offset (entries);
                                       // dep: tile of the same CTA that must be finished before this one is staged (-1: none)
struct NupgcmWarpDesc { int32_t estart, elen, stab, rtab; };    // stream start (multiple of 8) / length, slice and row table offsets
struct NupgcmSlice { int32_t eoff, roff, nrows, lmax; };        // entry offset in the warp's stream, rows (relative to rtab), longest row;
                                       // nrows == 0: ONE row of lmax entries stored contiguously, processed by the whole warp;
                                       // nrows = 32 | nblk << 8: the first nblk blocks of 8 positions are stored "blocked" (csr.cu)
static const int kLongRow = 96;        // rows longer than this are such whole-warp items (they would give a slice a long jagged end)

static const int kPartialSlots = 24;   // >= memory+2 of GMRES
static const int kMaxMemory = 20;

// -------------------------------------------------------------------------------------------
// error plumbing
// -------------------------------------------------------------------------------------------

extern char g_nupgcm_err[512];

static inline int nupgcm_fail(nupgcm_ctx *ctx, int code, const char *fmt, const char *a = "",
                              const char *b = "") {
    char *dst = ctx ? ctx->err : g_nupgcm_err;
    snprintf(dst, 512, fmt, a, b);
    if (ctx) snprintf(g_nupgcm_err, 512, "%s", dst);
    return code;
}

#define NUPGCM_CUDA(ctx, call)                                                                  \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return nupgcm_fail((ctx), NUPGCM_ERR_CUDA, "CUDA error: %s at %s",                  \
                               cudaGetErrorString(e_), #call);                                  \
    } while (0)

#define NUPGCM_REQUIRE(ctx, cond, msg)                                                          \
    do {                                                                                        \
        if (!(cond)) return nupgcm_fail((ctx), NUPGCM_ERR_INVALID, "invalid argument: %s", msg); \
    } while (0)

// -------------------------------------------------------------------------------------------
// device helpers
// -------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// L2-coherent load: data written by other CTAs of the same (persistent) kernel must not be
// served from this SM's L1.
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int T>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum with a fixed reduction tree (deterministic).  `red` needs 32 doubles.
// Result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

#endif  // __CUDACC__
