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
//   * when the CTA's matrix slice fits in shared memory (the reference's shipped meshes: 14-29 MB
//     over 148 SMs) it is bulk-copied there once (TMA, cp.async.bulk) and stays for the whole
//     solve, together with the CTA's rows of the Krylov basis; otherwise the matrix streams
//     from L2/HBM;
//   * the only cross-CTA traffic is (i) the gather of the multiplied vector in the SpMV, read
//     with ld.global.cg (L2-coherent), and (ii) one flagged 16-byte word per CTA per dot product,
//     combined in a FIXED order (slot = CTA index; lane-strided sum, then xor-shuffle tree).
//     Every CTA therefore computes bit-identical scalars and takes identical branches; results
//     are run-to-run reproducible (no floating-point atomics anywhere);
//   * scalar recurrences (Givens rotations, packed R, back substitution) are replicated per CTA.
//
// The recurrences restate Krylov.jl 0.10 `cg!` and `gmres!`: warm start
// (x on entry is the initial guess), preconditioned stopping measure against
// atol + rtol*(initial measure), itmax = 2n when 0, MGS Arnoldi — or CGS2 with three, or (fused)
// two, grid reductions per iteration (see the orthogonalisation branches of k_gmres).
//
// Multi-GPU (template parameter MR): the same kernels run one rank per GPU on a row-block sharded
// system.  Rank r's CTA b is part r*gridDim.x + b of the partition; rows that a peer's SpMV gathers
// are pushed into the peer's memory over NVLink as flagged 16-byte words (HaloPush) and unpacked by
// the reader before its SpMV; reductions take two hops (CTA -> reducer CTA on the same GPU ->
// every rank's arena); a solve ends with an all-gather of the solution.  See comm.cu and DESIGN.md §6.
#include <algorithm>
#include <vector>

#include "common.cuh"

static const int kThreads = 512;       // 16 warps: up to 128 registers per thread, cheap CTA barriers

struct ResidentLayout {       // byte offsets into dynamic shared memory (all multiples of 16)
    int vals, cols, xs, foot, rp, bar, vec, total;
    int vec_rows;             // row capacity of one vector slice (0: vectors stay in global memory)
    // streaming form (matrix does not fit on chip, common.cuh "streaming SpMV tables"): per-warp rings of
    // kRingPieces x kPieceEntries entries, the mbarriers, and the arena of staged footprints
    int st_ring_v, st_ring_c, st_bars, st_xs;
    int streaming;            // 1: SpmvEngine<T,false> runs the tiled TMA streams described by st_*
};

static const int kRingEntries = kPieceEntries * kRingPieces;
static const uint32_t kPieceBytes = kPieceEntries * (sizeof(double) + sizeof(uint16_t));
static_assert((kPieceEntries & (kPieceEntries - 1)) == 0, "piece size must be a power of two");
static_assert(8 * 32 + kPieceEntries <= kRingEntries, "a block of 8 slice positions must fit in the ring next to one piece");
static_assert(kRingPieces >= 2, "double buffering at least");

struct KrylovArgs {
    const int32_t *rowptr;
    const int32_t *colidx;
    const double *vals;
    const int32_t *part;     // [grid+1] row ranges
    const uint16_t *loc;     // SM-resident tables (see SpmvEngine)
    const int32_t *foot_ptr; // [grid][2]: start (multiple of 4) and length of each CTA's footprint
    const int32_t *foot;
    const int32_t *perm;     // [n] internal row -> caller row (vectors cross the ABI in caller order)
    // streaming form (common.cuh "streaming SpMV tables"; indexed by blockIdx.x: this rank's CTAs only)
    const double *svals;
    const uint16_t *scols;
    const NupgcmTileDesc *tiles;
    const int32_t *tile_ptr;
    const NupgcmWarpDesc *wdesc;
    const NupgcmSlice *slices;
    const uint8_t *twcnt;
    const int32_t *srow;
    const int32_t *slen;
    const int32_t *sfoot;
    ResidentLayout lay;
    int n;
    const double *dinv;      // diagonal preconditioner or NULL
    double pscale;           // scalar preconditioner when dinv == NULL
    const double *b;
    double *x;
    double *work;            // CG: p, then r, Ap, p, x, d slices (6n)   GMRES: V ((mem+1) n), qbuf (2n)
    double atol, rtol;
    long long itmax;
    int mem;
    int orth;
    int poll_depth;          // replicas of the reduction slots (GridReduce)
    int xmode;               // sharded solves: 1 = gather / multi-rank broadcast reductions, 0 = two-level
    int xfence;              // 1: release.sys / acquire.sys on the inter-rank flags of publishing reductions
    unsigned long long *trace;   // debug: arrival/completion stamps of a window of reductions
    unsigned long long *barrier;
    double *partials;        // LLSlot [2][kPartialSlots][grid]
    double *hist;
    long long hist_cap;
    double *result;          // niter, solved, inconsistent, breakdown, rnorm, rnorm0, hist_len, aborted
    // sharded solves (nranks > 1): this rank's CTA b is part `rank * gridDim.x + b` of the partition
    int rank, nranks;
    unsigned xgen_base;      // sequence number of the communicator's inter-rank reductions so far
    char *arena[kMaxRanks];  // exchange arenas of all ranks as mapped here (common.cuh: nupgcm_comm)
    int push_lo[kMaxRanks], push_hi[kMaxRanks];   // rows of this rank that peer p's SpMV gathers
    const int32_t *halo_ptr;     // [grid+1] per CTA: columns of its rows owned by other ranks
    const int32_t *halo_idx;
    long long ll_off;            // byte offset of the flagged form of the exchange vectors in an arena
};

// Watchdog of every cross-CTA wait: a CTA that waits longer than this raises the abort word and
// all CTAs leave the solve, so a lost CTA becomes a reported error instead of a hung GPU.
static const unsigned long long kWaitTimeoutNs = 10000000000ULL;   // 10 s: also covers the launch skew between ranks (processes)

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- mbarrier / TMA bulk-copy primitives ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waits for the phase with the given parity.  Watchdog: a copy or partner that never arrives raises
// the abort word (if given) after kWaitTimeoutNs and the wait returns false instead of hanging the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, unsigned long long *abort_word = nullptr) {
    uint32_t done;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kWaitTimeoutNs) {
                if (abort_word) atomicExch(abort_word, 1ULL);
                return false;
            }
        }
    }
}
// 1-D TMA bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


// ---- grid-wide reductions ---------------------------------------------------------------------
// All-to-all exchange of per-CTA partial sums without atomics and without a separate barrier
// (NCCL "LL"-style flagged words): a value is stored as two 8-byte words {lo32|flag, hi32|flag};
// 8-byte accesses are single-copy atomic, so a reader that sees the expected flag in both words
// has the complete value.  flag = sequence number of the reduction inside this launch (slots are
// zeroed before the launch); two banks alternate so that a fast CTA cannot overwrite a value a
// slow CTA has not read yet.  Every CTA reads all grid slots of a value in the same fixed order:
// bit-identical results in all CTAs.
//
// publish = true additionally makes the vector rows this CTA wrote before the call visible to
// all CTAs after the call (release fence before the slot store, acquire fence after the poll):
// that is the grid barrier of the SpMV gather.  Reductions that only carry scalars skip the fences.
struct LLSlot {
    unsigned long long a, b;
};

// SYS = true: system scope — the word lives in (or is read by) another GPU's memory over NVLink.
template <bool RELEASE, bool SYS = false>
__device__ __forceinline__ void ll_store(LLSlot *p, double v, unsigned flag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long a = (bits & 0xffffffffULL) | ((unsigned long long)flag << 32);
    const unsigned long long b = (bits >> 32) | ((unsigned long long)flag << 32);
    if (SYS) {
        if (RELEASE)
            asm volatile("st.release.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
        else
            asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    } else {
        if (RELEASE)
            asm volatile("st.release.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
        else
            asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    }
}

template <bool ACQUIRE, bool SYS = false>
__device__ __forceinline__ void ll_load_raw(const LLSlot *p, unsigned long long &a, unsigned long long &b) {
    if (SYS) {
        if (ACQUIRE)
            asm volatile("ld.acquire.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
        else
            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    } else {
        if (ACQUIRE)
            asm volatile("ld.acquire.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
        else
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    }
}

static const unsigned kTraceStart = 2000, kTraceWindow = 256, kTraceStamps = 6;   // reductions recorded by the debug trace
static const int kPollWarps = 5;           // 5 x 32 lanes cover grids up to 160 CTAs (B200: 148 SMs)
static const int kMaxReplicas = 8;
static const int kMaxV = 12;       // values one polling lane keeps in flight (largest without spills)

// ---- warp specialisation ------------------------------------------------------------------------
// Warps 0..kMainWarps-1 ("main") run the solver; warps kMainWarps.. ("comm") do nothing but the
// cross-CTA exchange.  The two roles never share registers at a program point, so the polling
// code (many 16-byte loads in flight) does not spill the solver's state and vice versa — an
// earlier single-role version spilled 0.6-2 KB per thread, and every acquire (which invalidates
// L1) turned the spill reloads into L2 round trips.  Main and comm warps meet at two named
// barriers (request / response); main warps synchronise among themselves on a third.
static const int kCommWarps = 5;            // kMainWarps = 11 (common.cuh)
static const int kMainThreads = 32 * kMainWarps, kCommThreads = 32 * kCommWarps;
static_assert(kMainThreads + kCommThreads == kThreads, "role split must cover the CTA");
#define NUPGCM_BAR_MAIN 1
#define NUPGCM_BAR_REQ 2
#define NUPGCM_BAR_RSP 3
#define NUPGCM_BAR_COMM 4

__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void main_sync() { named_sync(NUPGCM_BAR_MAIN, kMainThreads); }

static const int kReqExit = -1, kReqGather = -2;
struct CommMailbox {
    int count;                       // > 0: values to reduce; 0: barrier only; kReqExit; kReqGather: stage the
                                     // footprints of `xin` for the streaming SpMV (no response)
    const double *xin;
    int from_warps;                  // 1: the single value is the sum of wpart[0..kMainWarps)
    int publish;                     // release/acquire: also publishes the CTA's global rows
    int dead;                        // set by the comm warps when the watchdog fired
    double wpart[kMainWarps];
    double in[kPartialSlots];
    double out[kPartialSlots];
    double part[kPartialSlots][kPollWarps];
    double xpart[kPartialSlots][kMaxRanks];   // sharded solves: per-rank totals of each value
};

// Scratch layout (LLSlot units) of one bank of the exchange area.
static const int kBcastCopies = 8, kBcastStride = 32;
__host__ __device__ inline size_t exch_bank_slots(int gpad) {
    return (size_t)(kMaxReplicas + kPartialSlots) * gpad + (size_t)kBcastCopies * kBcastStride;
}

// Back-off between two polls of a flagged word (ns; 0 = spin).  Every CTA polls the same few cache lines, so
// the polls themselves load the L2 slices that hold them (NUPGCM_POLL_SLEEP, tools/reduce_sweep.py).
__constant__ int c_poll_sleep_ns = 0;

// Wait (with watchdog) until the flagged word at p carries `gen`; returns its value.
template <bool ACQUIRE, bool SYS = false>
__device__ __forceinline__ double wait_flagged(const LLSlot *p, unsigned gen, unsigned long long *abort_word, bool &bad) {
    unsigned long long a, b;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        ll_load_raw<ACQUIRE, SYS>(p, a, b);
        if ((unsigned)(a >> 32) == gen && (unsigned)(b >> 32) == gen)
            return __longlong_as_double((long long)((a & 0xffffffffULL) | (b << 32)));
        if (c_poll_sleep_ns > 0) __nanosleep((unsigned)c_poll_sleep_ns);
        if ((++spins & 255u) == 0) {
            unsigned long long fl;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(fl) : "l"(abort_word) : "memory");
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (fl != 0 || now - t0 > kWaitTimeoutNs) {
                atomicExch(abort_word, 1ULL);
                bad = true;
                return 0.0;
            }
        }
    }
}

// One value per CTA, all-to-all: every CTA stores its partial (nrep replicas) and reads all grid
// slots, one per comm lane; lanes are combined by an xor-shuffle tree and warps in index order —
// a fixed summation order, identical in every CTA.  PUBLISH uses release stores / acquire loads so
// that the global-memory rows written by the CTA before the call are visible to all CTAs after it.
// Inter-rank slot of the exchange arena (common.cuh): value `val`, replica `rep`, written by `rank`.
__device__ __forceinline__ LLSlot *xr_slot(char *arena, unsigned bank, int val, int rep, int rank) {
    return reinterpret_cast<LLSlot *>(arena + kArenaXrOffset) +
           (((size_t)bank * kPartialSlots + val) * kXRep + rep) * kMaxRanks + rank;
}

// MR (multi-rank, sharded solves): the all-to-all below yields the sum over THIS rank's CTAs in
// every CTA; CTA 0 then stores it into every rank's arena (kXRep replicas, one remote 16-byte store
// per comm lane — over NVLink for the peers) and every CTA adds the nranks words of its own arena in
// rank order.  All CTAs of all ranks obtain bit-identical totals, so the ranks' scalar recurrences
// and branches stay in lockstep without any further synchronisation.  Flags of the inter-rank
// words are xgen = (communicator sequence number) + gen: these words are never zeroed between
// solves (a peer may already be writing the next solve's first word), and their bank is xgen & 1.
template <bool PUBLISH, bool MR>
__device__ __forceinline__ bool comm_sum1(CommMailbox *mb, const KrylovArgs &a, LLSlot *bank_base, unsigned long long *abort_word,
                                          unsigned long long *trow, bool tr, unsigned gen, int grid, int gpad,
                                          int nrep) {
    const int ct = threadIdx.x - kMainThreads, cw = ct >> 5, lane = ct & 31;
    const int bid = blockIdx.x;
    if (cw == 0) {
        double val;
        if (mb->from_warps) {
            val = lane < kMainWarps ? mb->wpart[lane] : 0.0;
            val = warp_sum(val);                              // fixed tree over the main warps' partials
        } else {
            val = mb->in[0];
        }
        if (lane < nrep) ll_store<PUBLISH>(bank_base + (size_t)lane * gpad + bid, val, gen);
    }
    if (tr) trow[2 * (size_t)grid] = global_timer_ns();            // own partial stored
    bool bad = false;
    const int c = cw * 32 + lane;
    double v = 0.0;
    if (c < grid) v = wait_flagged<PUBLISH>(bank_base + (size_t)(bid % nrep) * gpad + c, gen, abort_word, bad);
    v = warp_sum(v);
    if (lane == 0) mb->part[0][cw] = v;
    if (bad) mb->dead = 1;
    if (tr) trow[3 * (size_t)grid] = global_timer_ns();
    named_sync(NUPGCM_BAR_COMM, kCommThreads);
    if (tr) trow[4 * (size_t)grid] = global_timer_ns();
    const int npw = (grid + 31) >> 5;
    if constexpr (MR) {
        const unsigned xgen = a.xgen_base + gen, xbank = xgen & 1u;
        if (bid == 0 && ct < a.nranks * kXRep) {
            double total = 0.0;
            for (int g = 0; g < npw; ++g) total += mb->part[0][g];
            ll_store<PUBLISH, true>(xr_slot(a.arena[ct % a.nranks], xbank, kPartialSlots - 1, ct / a.nranks, a.rank), total, xgen);
        }
        if (cw == 0) {
            bool xbad = false;
            double xv = 0.0;
            if (lane < a.nranks)
                xv = wait_flagged<PUBLISH, true>(xr_slot(a.arena[a.rank], xbank, kPartialSlots - 1, bid % kXRep, lane), xgen, abort_word, xbad);
            double total = 0.0;
            for (int r = 0; r < a.nranks; ++r) total += __shfl_sync(0xffffffffu, xv, r);
            if (lane == 0) mb->out[0] = total;
            if (xbad) mb->dead = 1;
            bad = bad || xbad;
        }
    } else {
        if (ct == 0) {
            double total = 0.0;
            for (int g = 0; g < npw; ++g) total += mb->part[0][g];
            mb->out[0] = total;
        }
    }
    return bad;
}

// `count` values per CTA (CGS2 projections), gather-broadcast: CTA j is the reducer of value j —
// it waits for the grid's partials of that value, adds them in the same fixed order as comm_sum1
// and stores the total into kBcastCopies replicated result blocks; every CTA then waits for the
// `count` totals of one replica.  An all-to-all of count values would issue grid² x count sector
// reads per poll round (438 k at 148 CTAs x 20 values) against a few hundred cache lines and was
// measured at 6-12 us; this form issues (grid + copies) x count.
// MR: the reducer CTA of value j exchanges its rank's total with the reducer CTAs of value j on the
// other ranks (one remote store per peer), adds the nranks totals in rank order and broadcasts.
template <bool MR, bool PUBLISH>
__device__ __forceinline__ bool comm_sumN(CommMailbox *mb, const KrylovArgs &a, LLSlot *bank_base, unsigned long long *abort_word,
                                          unsigned long long *trow, bool tr, unsigned gen, int grid, int gpad,
                                          int count) {
    const int ct = threadIdx.x - kMainThreads, cw = ct >> 5, lane = ct & 31;
    const int bid = blockIdx.x;
    LLSlot *vals = bank_base + (size_t)kMaxReplicas * gpad;                 // [kPartialSlots][gpad]
    LLSlot *res = vals + (size_t)kPartialSlots * gpad;                      // [kBcastCopies][kBcastStride]
    // PUBLISH: value 0 carries the publication (release by its store, acquire by its reducer, release by
    // the reducer's result, acquire by every CTA's wait for result 0); the other values stay relaxed.
    if (ct < count) {
        if (PUBLISH && ct == 0) ll_store<true>(vals + bid, mb->in[0], gen);
        else ll_store<false>(vals + (size_t)ct * gpad + bid, mb->in[ct], gen);
    }
    if (tr) trow[2 * (size_t)grid] = global_timer_ns();
    bool bad = false;
    if (bid < count) {                                                      // reducer of value `bid`
        const int c = cw * 32 + lane;
        double v = 0.0;
        if (c < grid) v = (PUBLISH && bid == 0) ? wait_flagged<true>(vals + c, gen, abort_word, bad)
                                                : wait_flagged<false>(vals + (size_t)bid * gpad + c, gen, abort_word, bad);
        v = warp_sum(v);
        if (lane == 0) mb->part[0][cw] = v;
        named_sync(NUPGCM_BAR_COMM, kCommThreads);
        const int npw = (grid + 31) >> 5;
        if constexpr (MR) {
            const unsigned xgen = a.xgen_base + gen, xbank = xgen & 1u;
            if (ct < a.nranks) {
                double mine = 0.0;
                for (int g = 0; g < npw; ++g) mine += mb->part[0][g];
                ll_store<false, true>(xr_slot(a.arena[ct], xbank, bid, 0, a.rank), mine, xgen);
            }
            if (ct < kBcastCopies) {
                double total = 0.0;
                for (int r = 0; r < a.nranks; ++r)
                    total += wait_flagged<false, true>(xr_slot(a.arena[a.rank], xbank, bid, 0, r), xgen, abort_word, bad);
                if (PUBLISH && bid == 0) ll_store<true>(res + (size_t)ct * kBcastStride, total, gen);
                else ll_store<false>(res + (size_t)ct * kBcastStride + bid, total, gen);
            }
        } else if (ct < kBcastCopies) {
            double total = 0.0;
            for (int g = 0; g < npw; ++g) total += mb->part[0][g];
            if (PUBLISH && bid == 0) ll_store<true>(res + (size_t)ct * kBcastStride, total, gen);
                else ll_store<false>(res + (size_t)ct * kBcastStride + bid, total, gen);
        }
    }
    if (tr) trow[3 * (size_t)grid] = global_timer_ns();
    if (ct < count)
        mb->out[ct] = (PUBLISH && ct == 0) ? wait_flagged<true>(res + (size_t)(bid % kBcastCopies) * kBcastStride, gen, abort_word, bad)
                                           : wait_flagged<false>(res + (size_t)(bid % kBcastCopies) * kBcastStride + ct, gen, abort_word, bad);
    if (bad) mb->dead = 1;
    if (tr) trow[4 * (size_t)grid] = global_timer_ns();
    return bad;
}

// Sharded solves, xmode 1: every reduction (1..20 values, or a bare barrier) takes two hops.
//   hop 1 (inside the GPU)  every CTA stores its partials into flagged slots; CTA j — the reducer
//                           of value j on its rank — adds the rank's partials in CTA order;
//   hop 2 (over NVLink)     the reducer stores the rank's total into EVERY rank's arena (kXRep
//                           replicas each: nranks x kXRep 16-byte stores, one per comm lane), and
//                           every CTA of every rank polls the nranks words of each value in its own
//                           arena and adds them in rank order.
// All CTAs of all ranks obtain bit-identical totals.  PUBLISH carries release/acquire along the
// chain (gpu scope inside the GPU, sys scope across), after the main threads' system fence.
template <bool PUBLISH>
__device__ __forceinline__ bool comm_mr(CommMailbox *mb, const KrylovArgs &a, LLSlot *bank_base,
                                        unsigned long long *abort_word, unsigned gen, int grid, int gpad, int count) {
    const int ct = threadIdx.x - kMainThreads, cw = ct >> 5, lane = ct & 31;
    const int bid = blockIdx.x, P = a.nranks;
    LLSlot *vals = bank_base + (size_t)kMaxReplicas * gpad;                 // [kPartialSlots][gpad]
    const unsigned xgen = a.xgen_base + gen, xbank = xgen & 1u;
    if (count == 1 && mb->from_warps) {
        if (cw == 0) {
            double val = lane < kMainWarps ? mb->wpart[lane] : 0.0;
            val = warp_sum(val);
            if (lane == 0) ll_store<PUBLISH>(vals + bid, val, gen);
        }
    } else if (ct < count) {
        // value 0 carries the publication of this CTA's rows, the other values stay relaxed
        if (PUBLISH && ct == 0) ll_store<true>(vals + bid, mb->in[0], gen);
        else ll_store<false>(vals + (size_t)ct * gpad + bid, mb->in[ct], gen);
    }
    bool bad = false;
    if (bid < count) {                                                      // reducer of value `bid` on this rank
        const int c = cw * 32 + lane;
        double v = 0.0;
        if (c < grid) v = (PUBLISH && bid == 0) ? wait_flagged<true>(vals + c, gen, abort_word, bad)
                                                : wait_flagged<false>(vals + (size_t)bid * gpad + c, gen, abort_word, bad);
        v = warp_sum(v);
        if (lane == 0) mb->part[0][cw] = v;
        named_sync(NUPGCM_BAR_COMM, kCommThreads);
        if (ct < P * kXRep) {
            const int npw = (grid + 31) >> 5;
            double total = 0.0;
            for (int g = 0; g < npw; ++g) total += mb->part[0][g];
            // Publishing reductions.  Rows of this rank's own CTAs follow a gpu-scope release/acquire
            // chain (value 0: slot -> this reducer -> its word in the own arena -> every CTA).  Halo rows
            // of the Krylov vectors need nothing: they travel as self-validating flagged words.  The
            // plain remote stores of the closing all-gather were fenced at system scope by their CTA
            // (GridReduce::fence_remote) before this reducer could acquire that CTA's slot, so they are
            // performed at the peers before this flag is even issued — a relaxed store is enough (the
            // NCCL pattern: writers fence, then the flag).  xfence = 1 uses formal release.sys /
            // acquire.sys flags instead, at +1.6 us per hop (tools/xrank_latency.py).
            LLSlot *dst = xr_slot(a.arena[ct % P], xbank, bid, ct / P, a.rank);
            if (PUBLISH && bid == 0 && a.xfence) ll_store<true, true>(dst, total, xgen);
            else if (PUBLISH && bid == 0 && ct % P == a.rank) ll_store<true, false>(dst, total, xgen);   // rows of this rank's own CTAs
            else ll_store<false, true>(dst, total, xgen);
        }
    }
    // count * nranks words (up to 21 x 8 = 168) over the 160 comm threads
    for (int w = ct; w < count * P; w += kCommThreads) {
        const int j = w / P, r = w % P;
        const LLSlot *src = xr_slot(a.arena[a.rank], xbank, j, bid % kXRep, r);
        if (PUBLISH && j == 0 && a.xfence) mb->xpart[j][r] = wait_flagged<true, true>(src, xgen, abort_word, bad);
        else if (PUBLISH && j == 0 && r == a.rank) mb->xpart[j][r] = wait_flagged<true, false>(src, xgen, abort_word, bad);
        else mb->xpart[j][r] = wait_flagged<false, true>(src, xgen, abort_word, bad);
    }
    if (bad) mb->dead = 1;
    named_sync(NUPGCM_BAR_COMM, kCommThreads);
    if (ct < count) {
        double t = 0.0;
        for (int r = 0; r < P; ++r) t += mb->xpart[ct][r];
        mb->out[ct] = t;
    }
    return bad;
}

// Service loop of the comm warps: one request per grid-wide reduction, until the main warps post
// the exit request.
__device__ __forceinline__ bool stream_gather_service(const KrylovArgs &a, unsigned char *smem, const double *xin,
                                                      unsigned &calls, unsigned long long *abort_word);

template <bool MR = false>
__device__ __forceinline__ void comm_warp_loop(CommMailbox *mb, const KrylovArgs &a, int nrep, unsigned char *smem = nullptr) {
    LLSlot *slots = reinterpret_cast<LLSlot *>(a.partials);
    // sharded solves watch the arena's abort word, which every rank can raise
    unsigned long long *abort_word = MR ? reinterpret_cast<unsigned long long *>(a.arena[a.rank]) : a.barrier + 1;
    const int grid = gridDim.x, gpad = (grid + 7) & ~7;
    const int ct = threadIdx.x - kMainThreads;
    unsigned gen = 0, spmv_calls = 0;
    bool dead = false;
    for (;;) {
        named_sync(NUPGCM_BAR_REQ, kThreads);
        const int count = mb->count;
        if (count == kReqGather) {                               // streaming SpMV: stage the footprints, no response
            if (!stream_gather_service(a, smem, mb->xin, spmv_calls, abort_word)) mb->dead = 1;
            continue;
        }
        if (count < 0) break;
        gen += 1;
        if (!dead) {
            const bool tr = a.trace && gen >= kTraceStart && gen < kTraceStart + kTraceWindow && ct == 0;
            unsigned long long *trow = a.trace + ((size_t)(gen - kTraceStart) * kTraceStamps) * grid + blockIdx.x;
            if (tr) trow[1 * (size_t)grid] = global_timer_ns();        // request seen by the comm warps
            LLSlot *bank_base = slots + (size_t)(gen & 1) * exch_bank_slots(gpad);
            bool bad;
            if (MR && a.xmode == 1) {
                const int cnt = count > 0 ? count : 1;           // a bare barrier sums one zero
                bad = mb->publish ? comm_mr<true>(mb, a, bank_base, abort_word, gen, grid, gpad, cnt)
                                  : comm_mr<false>(mb, a, bank_base, abort_word, gen, grid, gpad, cnt);
            } else if (count > 1) bad = mb->publish ? comm_sumN<MR, true>(mb, a, bank_base, abort_word, trow, tr, gen, grid, gpad, count)
                                                    : comm_sumN<MR, false>(mb, a, bank_base, abort_word, trow, tr, gen, grid, gpad, count);
            else if (mb->publish) bad = comm_sum1<true, MR>(mb, a, bank_base, abort_word, trow, tr, gen, grid, gpad, nrep);
            else bad = comm_sum1<false, MR>(mb, a, bank_base, abort_word, trow, tr, gen, grid, gpad, nrep);
            dead = __any_sync(0xffffffffu, bad) || mb->dead != 0;
            if (MR && dead && (ct & 31) == 0) {
                // take the other ranks down too: their waits poll their own arena's abort word
                for (int r = 0; r < a.nranks; ++r)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.arena[r]), "l"(1ULL) : "memory");
            }
        }
        named_arrive(NUPGCM_BAR_RSP, kThreads);
    }
}

// Main-warp side of the exchange.
struct GridReduce {
    CommMailbox *mb;
    unsigned long long *trace;
    unsigned gen;
    bool dead;
    bool remote;                     // sharded solve: this CTA also stores rows into peer GPUs' memory
    __device__ __forceinline__ void init(CommMailbox *m, unsigned long long *tr = nullptr, bool pushes_remote = false) {
        mb = m;
        trace = tr;
        gen = 0;
        dead = false;
        remote = pushes_remote;
    }
    // Rows stored over NVLink must be performed at the peers before the flag chain of a publishing
    // reduction starts.  One cumulative system fence per CTA, by thread 0 after a CTA barrier (the
    // NCCL "barrier, one thread fences, then the flag" pattern): a fence in each of the 352 main
    // threads of each CTA was measured at 8.7 us per reduction, this form at ~1.6 us and only in
    // CTAs that actually push.
    __device__ __forceinline__ void fence_remote() {
        if (remote) {
            main_sync();
            if (threadIdx.x == 0) __threadfence_system();
        }
    }
    __device__ __forceinline__ bool aborted() const { return dead; }
    __device__ __forceinline__ void round_trip() {
        gen += 1;
        const bool tr = trace && gen >= kTraceStart && gen < kTraceStart + kTraceWindow && threadIdx.x == 0;
        unsigned long long *trow = trace + ((size_t)(gen - kTraceStart) * kTraceStamps) * gridDim.x + blockIdx.x;
        if (tr) trow[0] = global_timer_ns();                          // main warps post the request
        named_arrive(NUPGCM_BAR_REQ, kThreads);
        named_sync(NUPGCM_BAR_RSP, kThreads);
        if (tr) trow[5 * (size_t)gridDim.x] = global_timer_ns();      // main warps resume
        dead = mb->dead != 0;
    }
    // Sum over the whole grid of a per-thread value (main warps only).  Every main thread gets
    // the result.  The block-level part is one shuffle tree per warp; the comm warps add the
    // kMainWarps warp partials.
    template <bool PUBLISH>
    __device__ __forceinline__ double sum_threads(double v) {
        if (PUBLISH) fence_remote();
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) mb->wpart[threadIdx.x >> 5] = v;
        if (threadIdx.x == 0) { mb->count = 1; mb->from_warps = 1; mb->publish = PUBLISH ? 1 : 0; }
        round_trip();
        return mb->out[0];
    }
    // Sum of `count` values per CTA that main threads put into mb->in[0..count) beforehand
    // (followed by main_sync()).  Results in mb->out[0..count).
    __device__ __forceinline__ void sumN(int count, bool publish = false) {
        if (publish) fence_remote();
        if (threadIdx.x == 0) { mb->count = count; mb->from_warps = 0; mb->publish = publish ? 1 : 0; }
        round_trip();
    }
    // Grid barrier that publishes this CTA's global-memory rows.
    __device__ __forceinline__ void barrier() {
        fence_remote();
        if (threadIdx.x == 0) { mb->count = 0; mb->from_warps = 0; mb->in[0] = 0.0; mb->publish = 1; }
        round_trip();
    }
    // Tell the comm warps to leave (call once, at the end, by all main threads).
    __device__ __forceinline__ void finish() {
        if (threadIdx.x == 0) mb->count = kReqExit;
        named_arrive(NUPGCM_BAR_REQ, kThreads);
    }
};

static size_t reduce_scratch_bytes(int grid) { return 2 * exch_bank_slots((grid + 7) & ~7) * sizeof(LLSlot); }

// ---- SpMV over the CTA's rows -----------------------------------------------------------------
// Two forms, chosen at launch:
//  * streaming (RES = false): entries are read from global memory (L2/HBM) every time; the
//    multiplied vector is gathered with ld.cg.  Used when the CTA's slice does not fit on chip.
//  * SM-resident (RES = true): at kernel start the CTA bulk-copies (TMA, cp.async.bulk) its slice
//    of values and 16-bit local column indices into shared memory, where it stays for the whole
//    solve; each SpMV first stages the CTA's column footprint of the multiplied vector into
//    shared memory (one ld.cg sweep) and then runs entirely out of shared memory.
// f(row, (A xin)[row]) is called by one lane per row.

// Shared-memory layout of the SM-resident form.  n_vec vector slices of vec_rows rows follow the
// matrix when they fit as well.
static ResidentLayout resident_layout(int max_nnz, int max_foot, int max_rows, int n_vec, int limit) {
    ResidentLayout L;
    int off = 0;
    L.vals = off; off += ((max_nnz + 2 + 1) & ~1) * 8;
    L.cols = off; off += (((max_nnz + 8 + 7) & ~7) * 2 + 15) & ~15;
    L.xs = off;   off += ((max_foot + 1) & ~1) * 8;
    L.foot = off; off += ((max_foot + 3) & ~3) * 4;
    L.rp = off;   off += ((max_rows + 1 + 4 + 3) & ~3) * 4;
    L.bar = off;  off += 16;
    L.vec = off;
    L.vec_rows = 0;
    const long long vec_bytes = (long long)n_vec * ((max_rows + 1) & ~1) * 8;
    if (off + vec_bytes <= limit) {
        L.vec_rows = (max_rows + 1) & ~1;
        off += (int)vec_bytes;
    }
    L.total = off;
    return L;
}

template <int T, bool RES>
struct SpmvEngine {
    const KrylovArgs &a;
    int r0, r1;
    // resident state (pointers are biased so that global positions index them directly)
    const double *vs;
    const uint16_t *cs;
    const int32_t *rp;
    const int32_t *foot;
    double *xs;
    int nfoot;

    __device__ __forceinline__ void finish() {}

    __device__ __forceinline__ SpmvEngine(const KrylovArgs &args, int r0_, int r1_, unsigned char *smem, CommMailbox *)
        : a(args), r0(r0_), r1(r1_) {
        if constexpr (RES) {
            const ResidentLayout &L = a.lay;
            uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);
            const int k0 = a.rowptr[r0], k1 = a.rowptr[r1];
            const int lb = a.rank * gridDim.x + blockIdx.x;   // position in the partition (all ranks)
            const int f0 = a.foot_ptr[2 * lb];
            nfoot = a.foot_ptr[2 * lb + 1];
            const int ka = k0 & ~1, kc = k0 & ~7, ra = r0 & ~3;
            const uint32_t bf = (uint32_t)(((nfoot + 3) & ~3) * 4);
            const uint32_t bv = (uint32_t)(((k1 - ka + 1) & ~1) * 8);
            const uint32_t bc = (uint32_t)(((k1 - kc + 7) & ~7) * 2);
            const uint32_t br = (uint32_t)(((r1 + 1 - ra + 3) & ~3) * 4);
            if (threadIdx.x == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            // All pieces ride on ONE phase of the barrier: a single arrive.expect_tx with the total byte
            // count, then the copies (values in pieces of 32 KB).  Round 1 issued one arrive.expect_tx
            // PER copy on a barrier of count 1 — the second arrival could land in the next phase and
            // the waits went out of step (the "fault with several copies outstanding"); it was never
            // the hardware.  Thread 0 alone waits (with the watchdog); the CTA barrier publishes.
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, bv + bc + br + bf);
                for (uint32_t done = 0; done < bv; done += 32768u)
                    bulk_g2s(smem + L.vals + done, reinterpret_cast<const unsigned char *>(a.vals + ka) + done,
                             min(bv - done, 32768u), bar);
                if (bc) bulk_g2s(smem + L.cols, a.loc + kc, bc, bar);
                if (br) bulk_g2s(smem + L.rp, a.rowptr + ra, br, bar);
                if (bf) bulk_g2s(smem + L.foot, a.foot + f0, bf, bar);
                mbar_wait(bar, 0, a.barrier + 1);
            }
            __syncthreads();
            vs = reinterpret_cast<const double *>(smem + L.vals) - ka;
            cs = reinterpret_cast<const uint16_t *>(smem + L.cols) - kc;
            rp = reinterpret_cast<const int32_t *>(smem + L.rp) - ra;
            xs = reinterpret_cast<double *>(smem + L.xs);
            foot = reinterpret_cast<const int32_t *>(smem + L.foot);
        }
    }

    template <class F>
    __device__ __forceinline__ void run(const double *xin, F &&f) {
        const int lane = threadIdx.x & (T - 1);
        const int g = threadIdx.x / T;
        const int G = kMainThreads / T;
        if constexpr (RES) {
            main_sync();                                     // previous readers of xs are done
            {   // stage the footprint of xin.  The gathers are L2 round trips (~0.7 us under load) and a
                // thread owns up to ~14 footprint entries: issue up to 16 of them before the first use
                // so that the whole footprint costs one round trip instead of one per group of four.
                const int nt = kMainThreads;
                constexpr int U = 16;
                for (int base = threadIdx.x; base < nfoot; base += U * nt) {
                    double xv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int i = base + u * nt;
                        xv[u] = i < nfoot ? ld_cg(xin + foot[i]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int i = base + u * nt;
                        if (i < nfoot) xs[i] = xv[u];
                    }
                }
            }
            main_sync();
            for (int base = r0; base < r1; base += G) {      // uniform trip count over the CTA
                const int row = base + g;
                const bool active = row < r1;
                double acc = 0.0, acc2 = 0.0;
                if (active) {
                    const int beg = rp[row], end = rp[row + 1];
                    int k = beg + lane;
                    for (; k + T < end; k += 2 * T) {
                        acc = fma(vs[k], xs[cs[k]], acc);
                        acc2 = fma(vs[k + T], xs[cs[k + T]], acc2);
                    }
                    if (k < end) acc = fma(vs[k], xs[cs[k]], acc);
                    acc += acc2;
                }
                acc = group_sum<T>(acc);
                if (active && lane == 0) f(row, acc);
            }
        } else {
            for (int base = r0; base < r1; base += G) {      // uniform trip count over the CTA
                const int row = base + g;
                const bool active = row < r1;
                double acc = 0.0;
                if (active) {
                    const int32_t beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
                    // issue the loads of up to 4 entries per lane before the dependent gathers
                    int32_t k = beg + lane;
                    double acc2 = 0.0;
                    for (; k + 3 * T < end; k += 4 * T) {
                        const double v0 = __ldg(a.vals + k), v1 = __ldg(a.vals + k + T);
                        const double v2 = __ldg(a.vals + k + 2 * T), v3 = __ldg(a.vals + k + 3 * T);
                        const int32_t c0 = __ldg(a.colidx + k), c1 = __ldg(a.colidx + k + T);
                        const int32_t c2 = __ldg(a.colidx + k + 2 * T), c3 = __ldg(a.colidx + k + 3 * T);
                        const double x0 = ld_cg(xin + c0), x1 = ld_cg(xin + c1);
                        const double x2 = ld_cg(xin + c2), x3 = ld_cg(xin + c3);
                        acc = fma(v0, x0, acc);
                        acc2 = fma(v1, x1, acc2);
                        acc = fma(v2, x2, acc);
                        acc2 = fma(v3, x3, acc2);
                    }
                    for (; k < end; k += T) acc = fma(__ldg(a.vals + k), ld_cg(xin + __ldg(a.colidx + k)), acc);
                    acc += acc2;
                }
                acc = group_sum<T>(acc);
                if (active && lane == 0) f(row, acc);
            }
        }
    }
};


// Streaming form (common.cuh "streaming SpMV tables"): the CTA's slice of the reordered matrix does not
// fit on chip and is pulled from HBM every SpMV — by the copy engine, not by the threads:
//   * every solver warp owns one contiguous stream of entries (values + 16-bit offsets, no padding)
//     and a private ring of kRingPieces x kPieceEntries entries in shared memory; lane 0 keeps the
//     ring full with TMA bulk copies (one mbarrier phase per piece: a single arrive.expect_tx, then
//     the value and the offset copy).  Warps never synchronise with each other, and the first pieces
//     of the NEXT SpMV are already in flight when one ends (the matrix is immutable);
//   * the multiplied vector is read from shared memory: the comm warps, idle during an SpMV, stage
//     the footprint of each tile (one ld.cg gather per distinct column) into an arena, as many tiles
//     ahead of the solver warps as the arena holds (one full / one empty mbarrier per tile, each
//     completing one phase per SpMV; the host plans the arena slots and which tile must be finished
//     before a slot is overwritten);
//   * a warp works through slices of 32 length-sorted rows in jagged-diagonal order, one row per lane:
//     consecutive lanes read consecutive ring entries, no cross-lane reduction, lanes only drop out
//     over the last few positions of a slice, and the table builder places entries so that the gather
//     of a position is (nearly) free of bank conflicts.  (The first version of this engine processed
//     CSR rows with groups of T lanes and spent 52 warp instructions per 32 entries on divergent
//     unrolled row loops, ring-index arithmetic and shuffles; the second staged footprints in two
//     buffers and lost 55 % of the solver warps' time to footprint waits, per-tile imbalance and bank
//     conflicts — profiles/stream_spmv_history_r02.txt.)
// f(row, (A xin)[row]) is called once per row, by one lane, in no particular order.
struct StreamShared {                        // pointers into dynamic shared memory, same for all threads
    double *ring_v;                          // [kMainWarps][kRingEntries]
    uint16_t *ring_c;                        // [kMainWarps][kRingEntries]
    uint64_t *full;                          // [kMainWarps][kRingPieces]
    uint64_t *xs_full, *xs_empty;            // [kMaxTiles] each
    double *xs;                              // arena, kArenaEntries
    __device__ __forceinline__ StreamShared(const ResidentLayout &L, unsigned char *smem) {
        ring_v = reinterpret_cast<double *>(smem + L.st_ring_v);
        ring_c = reinterpret_cast<uint16_t *>(smem + L.st_ring_c);
        full = reinterpret_cast<uint64_t *>(smem + L.st_bars);
        xs_full = full + kMainWarps * kRingPieces;
        xs_empty = xs_full + kMaxTiles;
        xs = reinterpret_cast<double *>(smem + L.st_xs);
    }
};

// Comm-warp side of one streaming SpMV: stage the footprint of every tile of this CTA.  `calls` counts
// the SpMVs so far (every per-tile mbarrier completes one phase per SpMV: parity = calls & 1).  The
// gathers are L2 round trips under full HBM load (2-4 us each): the work is cut into rounds of
// kGatherU entries per thread, all of a round's loads are in flight together, and the column ids of
// the NEXT round (of this or the next tile) are fetched while the current round's vector entries are
// on their way, so a round costs one round trip instead of two dependent ones.
static const int kGatherU = 12;
__device__ __forceinline__ bool stream_gather_service(const KrylovArgs &a, unsigned char *smem, const double *xin,
                                                      unsigned &calls, unsigned long long *abort_word) {
    const StreamShared sh(a.lay, smem);
    const int ct = threadIdx.x - kMainThreads;
    const int t0 = a.tile_ptr[blockIdx.x], t1 = a.tile_ptr[blockIdx.x + 1];
    const unsigned par = calls & 1u;
    bool ok = true;
    constexpr int U = kGatherU, RND = kGatherU * kCommThreads;
    // A round = entries [base, base + RND) of the footprint of tile t.  Three rounds overlap: the column
    // ids of round r+2 and the vector entries of round r+1 are in flight while round r is stored.
    struct Cursor { int t, base; NupgcmTileDesc td; };
    auto advance = [&](Cursor &c) {
        c.base += RND;
        if (c.base >= c.td.foot_len) {
            c.base = 0;
            if (++c.t < t1) c.td = a.tiles[c.t];
        }
    };
    auto load_idx = [&](int32_t *dst, const Cursor &c) {
        const int32_t *foot = a.sfoot + c.td.foot_off;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = c.base + ct + u * kCommThreads;
            dst[u] = (c.t < t1 && i < c.td.foot_len) ? __ldg(foot + i) : 0;
        }
    };
    auto load_x = [&](double *dst, const int32_t *idx) {
#pragma unroll
        for (int u = 0; u < U; ++u) dst[u] = ld_cg(xin + idx[u]);
    };
    Cursor c0{t0, 0, a.tiles[t0]}, c1 = c0, c2;
    int32_t idx[U];
    double xa[U], xb[U];
    load_idx(idx, c0);
    load_x(xa, idx);                         // round 0
    advance(c1);
    load_idx(idx, c1);                       // ids of round 1
    c2 = c1;
    while (c0.t < t1) {
        if (c1.t < t1) load_x(xb, idx);      // round r+1
        advance(c2);
        load_idx(idx, c2);                   // ids of round r+2
        if (c0.base == 0 && c0.td.dep >= 0) ok = mbar_wait(sh.xs_empty + c0.td.dep, par, abort_word) && ok;
        double *dst = sh.xs + c0.td.xs_off;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = c0.base + ct + u * kCommThreads;
            if (i < c0.td.foot_len) dst[i] = xa[u];
        }
        if (c0.base + RND >= c0.td.foot_len) mbar_arrive(sh.xs_full + (c0.t - t0));   // tile complete (release: the stores above)
        c0 = c1;
        c1 = c2;
#pragma unroll
        for (int u = 0; u < U; ++u) xa[u] = xb[u];
    }
    ++calls;
    return ok;
}

template <int T>
struct SpmvEngine<T, false> {
    const KrylovArgs &a;
    int r0, r1;
    bool legacy;
    CommMailbox *mb;
    // this warp's stream
    double *rv;
    uint16_t *rc;
    uint64_t *full;
    const double *gv;
    const uint16_t *gc;
    const NupgcmSlice *slices;
    const int32_t *srow, *slen;
    uint64_t *xs_full, *xs_empty;
    const unsigned char *xs;
    int npieces, t0, t1;
    // pieces go through the ring slots round-robin, across run() calls: slot / parity of the next wait
    // (all lanes) and slot of the next issue (lane 0)
    int wslot, islot;
    unsigned wpar;
    unsigned calls;           // run() calls so far (parity of the per-tile mbarriers)
    bool failed;

    __device__ __forceinline__ void issue(int p) {                // lane 0 of the warp
        uint64_t *bar = full + islot;
        mbar_expect_tx(bar, kPieceBytes);
        bulk_g2s(rv + islot * kPieceEntries, gv + (size_t)p * kPieceEntries, kPieceEntries * 8, bar);
        bulk_g2s(rc + islot * kPieceEntries, gc + (size_t)p * kPieceEntries, kPieceEntries * 2, bar);
        islot = islot + 1 == kRingPieces ? 0 : islot + 1;
    }
    __device__ __forceinline__ void wait_piece(unsigned long long *abort_word) {
        if (!mbar_wait(full + wslot, wpar, abort_word)) failed = true;
        if (++wslot == kRingPieces) { wslot = 0; wpar ^= 1u; }
    }

    __device__ __forceinline__ SpmvEngine(const KrylovArgs &args, int r0_, int r1_, unsigned char *smem, CommMailbox *mailbox)
        : a(args), r0(r0_), r1(r1_), mb(mailbox) {
        legacy = a.lay.streaming == 0;
        wslot = islot = 0;
        wpar = 0;
        calls = 0;
        failed = false;
        if (legacy) return;
        const StreamShared sh(a.lay, smem);
        for (int i = threadIdx.x; i < kMainWarps * kRingPieces + 2 * kMaxTiles; i += blockDim.x)
            mbar_init(sh.full + i, i < kMainWarps * kRingPieces ? 1 : (i < kMainWarps * kRingPieces + kMaxTiles ? kCommThreads : kMainWarps));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        t0 = a.tile_ptr[blockIdx.x];
        t1 = a.tile_ptr[blockIdx.x + 1];
        xs_full = sh.xs_full;
        xs_empty = sh.xs_empty;
        xs = reinterpret_cast<const unsigned char *>(sh.xs);
        const int wid = threadIdx.x >> 5;
        if (wid >= kMainWarps) return;
        const NupgcmWarpDesc wd = a.wdesc[(size_t)blockIdx.x * kMainWarps + wid];
        rv = sh.ring_v + (size_t)wid * kRingEntries;
        rc = sh.ring_c + (size_t)wid * kRingEntries;
        full = sh.full + wid * kRingPieces;
        gv = a.svals + wd.estart;
        gc = a.scols + wd.estart;
        slices = a.slices + wd.stab;
        srow = a.srow + wd.rtab;
        slen = a.slen + wd.rtab;
        npieces = (wd.elen + kPieceEntries - 1) / kPieceEntries;
        if ((threadIdx.x & 31) == 0)
            for (int p = 0; p < npieces && p < kRingPieces; ++p) issue(p);      // prime the ring
    }

    // Before the CTA exits: the pieces prefetched for an SpMV that never came must have landed.
    __device__ __forceinline__ void finish() {
        if (legacy) return;
        for (int p = 0; p < npieces && p < kRingPieces; ++p) wait_piece(a.barrier + 1);
    }

    // DBG (diagnostics only, nupgcm_diag_stream_spmv): 1 = pull the pieces through the ring without
    // touching them, 2 = multiply by x[lane] instead of the gathered entry (no bank conflicts),
    // 3 = per-warp activity clocks
    template <int DBG = 0, class F>
    __device__ __forceinline__ void run(const double *xin, F &&f) {
        if (legacy) {
            const int lane = threadIdx.x & (T - 1);
            const int g = threadIdx.x / T;
            const int G = kMainThreads / T;
            for (int base = r0; base < r1; base += G) {      // uniform trip count over the CTA
                const int row = base + g;
                const bool active = row < r1;
                double acc = 0.0;
                if (active) {
                    const int32_t beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
                    for (int32_t k = beg + lane; k < end; k += T)
                        acc = fma(__ldg(a.vals + k), ld_cg(xin + __ldg(a.colidx + k)), acc);
                }
                acc = group_sum<T>(acc);
                if (active && lane == 0) f(row, acc);
            }
            return;
        }
        if (t0 == t1) return;                                    // a CTA without rows posts nothing (both roles know)
        long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = DBG == 3 ? clock64() : 0;   // DBG 3: cycles per activity
        auto tick = [&](int k) { if (DBG == 3) { const long long now = clock64(); tk[k] += now - tlast; tlast = now; } };
        // ask the comm warps to stage the footprints of xin (no response: the tiles' mbarriers pace us)
        if (threadIdx.x == 0) { mb->count = kReqGather; mb->xin = xin; }
        named_arrive(NUPGCM_BAR_REQ, kThreads);
        unsigned long long *abort_word = a.barrier + 1;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const unsigned par = calls & 1u;
        int rpos = wslot * kPieceEntries;                        // ring position of the next entry to consume
        int landed = 0;                                          // pieces of this call known to have arrived
        int issued = npieces < kRingPieces ? npieces : kRingPieces;
        auto ensure = [&](int need) {                            // entries [0, need) of the stream are in the ring
            if (DBG == 4) return;                                // (timing experiment: no copies at all, stale ring)
            if (landed * kPieceEntries < need) {
                tick(3);
                do { wait_piece(abort_word); ++landed; } while (landed * kPieceEntries < need);
                tick(1);
            }
        };
        auto release = [&](int cons) {                           // entries [0, cons) are consumed: refill freed slots
            if (DBG == 4) return;
            const int lim = min(npieces, kRingPieces + cons / kPieceEntries);
            if (issued < lim) {
                tick(3);
                // Every lane's reads of the freed slot have returned (their values were consumed above):
                // the copy engine may overwrite it.  No proxy fence: this is write-after-READ (the consumer
                // release of any TMA pipeline); fence.proxy.async here cost ~1.9 k cycles per piece under load.
                __syncwarp();
                if (lane == 0)
                    for (int p = issued; p < lim; ++p) issue(p);
                issued = lim;
                if (DBG == 3) { __syncwarp(); tick(6); tk[7] += 1; }
            }
        };
        const unsigned char *xt = xs;                            // staged footprint of the current tile
        auto entry = [&](int idx) -> double {                    // product of ring entry idx with its vector entry
            if (idx >= kRingEntries) idx -= kRingEntries;
            const double xv = DBG == 2 ? *reinterpret_cast<const double *>(xt + 8 * lane)
                                       : *reinterpret_cast<const double *>(xt + rc[idx]);
            return rv[idx] * xv;
        };
        // 8 positions of 32 entries each that do not go round the end of the ring: every address is the
        // lane's base plus a compile-time constant (no index arithmetic per position)
        auto dense8 = [&](int base, double &s0, double &s1) {
            const double *pv = rv + base + lane;
            const uint16_t *pc = rc + base + lane;
#pragma unroll
            for (int p = 0; p < 8; p += 2) {
                const double x0 = DBG == 2 ? *reinterpret_cast<const double *>(xt + 8 * lane)
                                           : *reinterpret_cast<const double *>(xt + pc[32 * p]);
                const double x1 = DBG == 2 ? *reinterpret_cast<const double *>(xt + 8 * lane)
                                           : *reinterpret_cast<const double *>(xt + pc[32 * (p + 1)]);
                s0 = fma(pv[32 * p], x0, s0);
                s1 = fma(pv[32 * (p + 1)], x1, s1);
            }
        };
        // the first slice's tables are fetched before anything is waited for; then always one slice ahead
        int cur = 0;                                             // stream entries consumed so far (slices may be preceded by alignment padding)
        int si = 0;                                              // next slice of this warp's list
        // (two-stage: the header of item i+2 is requested while item i is processed, so that the row table of
        // item i+1 — whose address needs that item's header — can be requested without waiting for a header)
        NupgcmSlice sl = slices[0], sl2 = slices[1];
        int mylen = 0, myrow = -1;
        if (lane < max(sl.nrows & 0xff, 1)) { mylen = __ldg(slen + sl.roff + lane); myrow = __ldg(srow + sl.roff + lane); }
        NupgcmTileDesc td = a.tiles[t0];
        int mine_next = __ldg(a.twcnt + (size_t)t0 * kMainWarps + wid);       // my items in the first tile
        // results of whole-warp (long) rows are parked one per lane and handed to f together: f's global
        // accesses then cost one round trip per 32 rows instead of one per row
        int nparked = 0, prow = -1;
        double pval = 0.0;
        auto flush_parked = [&]() {
            if (lane < nparked) f(prow, pval);
            nparked = 0;
        };
        for (int tix = 0; tix < t1 - t0; ++tix) {
            const int xs_off = td.xs_off;
            const int mine = mine_next;                          // my items (slices, long rows) in this tile
            if (tix + 1 < t1 - t0) {                             // next tile's tables: in flight during this one
                td = a.tiles[t0 + tix + 1];
                mine_next = __ldg(a.twcnt + (size_t)(t0 + tix + 1) * kMainWarps + wid);
            }
            tick(2);
            if (!mbar_wait(xs_full + tix, par, abort_word)) failed = true;
            tick(0);
            xt = xs + (size_t)xs_off * 8;
            for (int s = 0; s < mine; ++s) {
                const NupgcmSlice cs = sl;
                const int clen = mylen, crow = myrow;
                ++si;
                sl = sl2;                                        // header of the next item: requested one item ago
                sl2 = slices[si + 1];                            // (the table ends with two null slices)
                mylen = 0;
                myrow = -1;
                if (lane < max(sl.nrows & 0xff, 1)) { mylen = __ldg(slen + sl.roff + lane); myrow = __ldg(srow + sl.roff + lane); }
                const int nr = cs.nrows & 0xff, nblk = cs.nrows >> 8;
                int off = cs.eoff;
                rpos += off - cur;                               // skip the alignment padding in front of a blocked slice
                if (rpos >= kRingEntries) rpos -= kRingEntries;
                double acc0 = 0.0, acc1 = 0.0;
                if (nr == 0) {
                    // a long row: the whole warp strides over its (contiguous) entries, 8 x 32 at a time
                    const int row = __shfl_sync(0xffffffffu, crow, 0);
                    tick(2);
                    for (int k0 = 0; k0 < cs.lmax; k0 += 256) {
                        const int m = min(256, cs.lmax - k0);
                        ensure(off + m);
                        if (DBG != 1) {
                            if (m == 256 && rpos + 256 <= kRingEntries) {
                                dense8(rpos, acc0, acc1);
                            } else {
#pragma unroll
                                for (int p = 0; p < 8; p += 2) {
                                    if (32 * p + lane < m) acc0 += entry(rpos + 32 * p + lane);
                                    if (32 * (p + 1) + lane < m) acc1 += entry(rpos + 32 * (p + 1) + lane);
                                }
                            }
                        }
                        off += m;
                        rpos += m;
                        if (rpos >= kRingEntries) rpos -= kRingEntries;
                        release(off);
                    }
                    tick(4);
                    acc0 = warp_sum(acc0 + acc1);
                    if (lane == nparked) { prow = row; pval = acc0; }
                    if (++nparked == 32) flush_parked();
                    cur = off;
                    tick(5);
                    continue;
                }
                const bool act = lane < nr;
                const int lmin = __shfl_sync(0xffffffffu, clen, nr - 1);
                tick(2);
                int j = 8 * nblk;
                // blocked part of a full slice: the lane's 8 offsets in one 16-byte load, its values in pairs
                for (int bq = 0; bq < nblk; ++bq) {
                    ensure(off + 256);
                    if (DBG != 1) {
                        int ic = rpos + lane * 8;
                        if (ic >= kRingEntries) ic -= kRingEntries;
                        const uint4 cv = *reinterpret_cast<const uint4 *>(rc + ic);
                        const unsigned cw[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp) {
                            int iv = rpos + pp * 64 + lane * 2;
                            if (iv >= kRingEntries) iv -= kRingEntries;
                            const double2 vv = *reinterpret_cast<const double2 *>(rv + iv);
                            const double x0 = *reinterpret_cast<const double *>(xt + (DBG == 2 ? 8u * lane : (cw[pp] & 0xffffu)));
                            const double x1 = *reinterpret_cast<const double *>(xt + (DBG == 2 ? 8u * lane : (cw[pp] >> 16)));
                            acc0 = fma(vv.x, x0, acc0);
                            acc1 = fma(vv.y, x1, acc1);
                        }
                    }
                    off += 256;
                    rpos += 256;
                    if (rpos >= kRingEntries) rpos -= kRingEntries;
                    release(off);
                }
                // positions every row of the slice has: nr entries each, lane q reads entry q
                for (; j + 8 <= lmin; j += 8) {
                    ensure(off + 8 * nr);
                    if (DBG != 1) {
                        if (nr == 32 && rpos + 256 <= kRingEntries) {
                            dense8(rpos, acc0, acc1);
                        } else if (act) {
#pragma unroll
                            for (int p = 0; p < 8; p += 2) {
                                acc0 += entry(rpos + p * nr + lane);
                                acc1 += entry(rpos + (p + 1) * nr + lane);
                            }
                        }
                    }
                    off += 8 * nr;
                    rpos += 8 * nr;
                    if (rpos >= kRingEntries) rpos -= kRingEntries;
                    release(off);
                }
                if (j < lmin) {
                    const int m = lmin - j;
                    ensure(off + m * nr);
                    if (DBG != 1 && act)
                        for (int p = 0; p < m; ++p) acc0 += entry(rpos + p * nr + lane);
                    off += m * nr;
                    rpos += m * nr;
                    if (rpos >= kRingEntries) rpos -= kRingEntries;
                    j = lmin;
                }
                tick(3);
                // the jagged end, also 8 positions at a time: rows are sorted by length, so position j is
                // held by lanes 0 .. cnt_j-1 and starts cnt_0 + .. + cnt_{j-1} entries further on
                while (j < cs.lmax) {
                    int start[9];
                    start[0] = 0;
#pragma unroll
                    for (int p = 0; p < 8; ++p)
                        start[p + 1] = start[p] + __popc(__ballot_sync(0xffffffffu, clen > j + p));
                    ensure(off + start[8]);
                    if (DBG != 1) {
#pragma unroll
                        for (int p = 0; p < 8; p += 2) {
                            if (clen > j + p) acc0 += entry(rpos + start[p] + lane);
                            if (clen > j + p + 1) acc1 += entry(rpos + start[p + 1] + lane);
                        }
                    }
                    off += start[8];
                    rpos += start[8];
                    if (rpos >= kRingEntries) rpos -= kRingEntries;
                    j += 8;
                    release(off);
                }
                release(off);
                cur = off;
                tick(4);
                if (act) f(crow, acc0 + acc1);
                tick(5);
            }
            flush_parked();
            __syncwarp();
            if (lane == 0) mbar_arrive(xs_empty + tix);
        }
        ++calls;
        // this SpMV is done: start pulling the first pieces of the next one
        __syncwarp();
        if (lane == 0 && DBG != 4)
            for (int p = 0; p < npieces && p < kRingPieces; ++p) issue(p);
        if (failed) mb->dead = 1;
        if (DBG == 3 && a.trace && lane == 0) {
            tick(2);
            unsigned long long *o = a.trace + ((size_t)blockIdx.x * kMainWarps + wid) * 8;
            for (int k = 0; k < 6; ++k) o[k] += (unsigned long long)tk[k];
            o[7] += (unsigned long long)tk[6];                   // ring refills (inside the position rows' code)
        }
    }
};

// CTA-local storage of a family of vectors: rows [r0, r1) of vector i are at at(i)[row].
// In shared memory when the launch reserved room for it, else in the global workspace.
struct VecSlices {
    double *base;
    size_t stride;
    __device__ __forceinline__ VecSlices(const KrylovArgs &a, unsigned char *smem, double *global, int r0) {
        if (a.lay.vec_rows > 0) {
            base = reinterpret_cast<double *>(smem + a.lay.vec) - r0;
            stride = (size_t)a.lay.vec_rows;
        } else {
            base = global;
            stride = (size_t)a.n;
        }
    }
    __device__ __forceinline__ double *at(int i) const { return base + (size_t)i * stride; }
};

// ---- halo push (sharded solves) ----------------------------------------------------------------
// Vectors that other CTAs gather from (the raw Krylov vector, CG's p, the iterate) live in the
// rank's exchange arena.  A rank's SpMV also gathers rows owned by other ranks (its halo); those are
// PUSHED by their owner: whenever a CTA stores row i of such a vector it also stores it, over
// NVLink, at the same arena offset in every peer whose halo range contains i.  The data is in the
// reader's own memory before the flag of the next publishing reduction arrives, so no gather ever
// crosses the link (remote loads would put ~2 us of latency into every SpMV).
template <bool MR>
struct HaloPush;
template <>
struct HaloPush<false> {
    __device__ __forceinline__ void init(const KrylovArgs &, int, int, int *, char **) {}
    __device__ __forceinline__ void put_all(const double *, double) const {}
};
template <>
struct HaloPush<true> {
    const int *range;       // shared: [2*j], [2*j+1] row range of active peer j (intersected with the CTA's rows)
    char *const *dest;      // shared: arena of active peer j
    char *local;
    char *const *all;       // arenas of all ranks
    const double *vec0;     // first exchange vector in the local arena
    const int32_t *hlist;   // columns of this CTA's rows owned by other ranks
    long long ll_off;
    int np, nh, rank, nranks;
    unsigned tag;           // sequence number of the publishing reduction that follows the stores
    // `sh_range` (2*kMaxRanks ints) and `sh_dest` (kMaxRanks pointers) are shared-memory scratch;
    // call from all threads, followed by a CTA barrier before the first put().
    __device__ __forceinline__ void init(const KrylovArgs &a, int r0, int r1, int *sh_range, char **sh_dest) {
        int cnt = 0;
        for (int p = 0; p < a.nranks; ++p) {
            if (p == a.rank) continue;
            const int lo = max(r0, a.push_lo[p]), hi = min(r1, a.push_hi[p]);
            if (lo < hi) {
                if (threadIdx.x == 0) { sh_range[2 * cnt] = lo; sh_range[2 * cnt + 1] = hi; sh_dest[cnt] = a.arena[p]; }
                cnt++;
            }
        }
        np = cnt;
        range = sh_range;
        dest = sh_dest;
        local = a.arena[a.rank];
        all = a.arena;
        rank = a.rank;
        nranks = a.nranks;
        vec0 = reinterpret_cast<const double *>(local + kArenaVecOffset);
        ll_off = a.ll_off;
        hlist = a.halo_idx + a.halo_ptr[blockIdx.x];
        nh = a.halo_ptr[blockIdx.x + 1] - a.halo_ptr[blockIdx.x];
        tag = 0;
    }
    // Row `row` of the exchange vector whose local entry is p_local: value + tag as ONE flagged
    // 16-byte word into the mailbox form of that vector in every peer that gathers the row.  No
    // fence: the reader accepts the word only when both halves carry the tag.
    __device__ __forceinline__ void put_row(const double *p_local, int row, double v) const {
        const size_t idx = (size_t)(p_local - vec0);
        for (int j = 0; j < np; ++j)
            if (row >= range[2 * j] && row < range[2 * j + 1])
                ll_store<false, true>(reinterpret_cast<LLSlot *>(dest[j] + ll_off) + idx, v, tag);
    }
    // plain store to every peer (the all-gather that ends a solve; followed by a system fence)
    __device__ __forceinline__ void put_all(const double *p_local, double v) const {
        for (int p = 0; p < nranks; ++p)
            if (p != rank)
                *reinterpret_cast<double *>(all[p] + (reinterpret_cast<const char *>(p_local) - local)) = v;
    }
    // Before an SpMV that gathers `xin` (an exchange vector of the local arena): take this CTA's halo
    // entries out of the mailbox form (waiting for the tag of the last publishing reduction) and
    // store them into the plain vector, where the SpMV engines read them like any other entry.
    // Several CTAs may unpack the same column: they store the same value.
    __device__ __forceinline__ bool unpack(const double *xin, unsigned want, unsigned long long *abort_word) const {
        bool bad = false;
        if (nh > 0) {
            const LLSlot *box = reinterpret_cast<const LLSlot *>(local + ll_off) + (size_t)(xin - vec0);
            double *dst = const_cast<double *>(xin);
            for (int i = threadIdx.x; i < nh; i += kMainThreads) {
                const int c = hlist[i];
                dst[c] = wait_flagged<false, true>(box + c, want, abort_word, bad);
            }
        }
        return bad;
    }
};
template <bool MR>
__device__ __forceinline__ bool pushes_remote(const HaloPush<MR> &hp) {
    (void)hp;              // halo rows travel as flagged words; only the closing all-gather fences
    return false;
}
// Arm the pushes that follow with the tag of the NEXT reduction (the one that publishes them), and
// the SpMV-side counterpart: unpack the halo of `xin`, published by the LAST reduction.
template <bool MR>
__device__ __forceinline__ void halo_arm(HaloPush<MR> &hp, const KrylovArgs &a, unsigned gen) {
    if constexpr (MR) hp.tag = a.xgen_base + gen + 1u;
}
template <bool MR>
__device__ __forceinline__ void halo_unpack(const HaloPush<MR> &hp, const KrylovArgs &a, unsigned gen, CommMailbox *mb,
                                            const double *xin) {
    if constexpr (MR) {
        if (hp.nh > 0) {
            if (hp.unpack(xin, a.xgen_base + gen, reinterpret_cast<unsigned long long *>(a.arena[a.rank]))) mb->dead = 1;
            main_sync();
        }
    }
}
template <bool MR>
__device__ __forceinline__ void halo_put(const HaloPush<MR> &hp, const double *p_local, int row, double v) {
    if constexpr (MR) hp.put_row(p_local, row, v);
}

// =============================================================================================
// CG  (Krylov.jl cg!)
// =============================================================================================
template <int T, bool RES, bool MR>
__global__ void __launch_bounds__(kThreads, 1) k_cg(const __grid_constant__ KrylovArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ CommMailbox mailbox;
    __shared__ int s_prange[2 * kMaxRanks];
    __shared__ char *s_pdest[kMaxRanks];
    if (threadIdx.x == 0) mailbox.dead = 0;
    const int lb = a.rank * gridDim.x + blockIdx.x;       // position in the partition (all ranks)
    const int r0 = a.part[lb], r1 = a.part[lb + 1];
    SpmvEngine<T, RES> eng(a, r0, r1, dyn_smem, &mailbox);
    HaloPush<MR> hp;
    hp.init(a, r0, r1, s_prange, s_pdest);
    __syncthreads();
    if (threadIdx.x >= kMainThreads) {                     // comm warps: exchange service only
        comm_warp_loop<MR>(&mailbox, a, a.poll_depth, dyn_smem);
        return;
    }
    GridReduce gr;
    gr.init(&mailbox, a.trace, pushes_remote(hp));
    const int n = a.n;
    // CTA-local vectors: r, Ap, local copy of p, the iterate, the Jacobi diagonal
    VecSlices loc(a, dyn_smem, a.work + n, r0);       // global fallback: work[n .. 6n)
    double *r = loc.at(0), *Ap = loc.at(1), *pl = loc.at(2), *xl = loc.at(3), *dl = loc.at(4);
    // vectors the other CTAs gather: in the exchange arena when the solve is sharded
    double *xbase = MR ? reinterpret_cast<double *>(a.arena[a.rank] + kArenaVecOffset) : nullptr;
    double *pg = MR ? xbase : a.work;                  // p as the other CTAs gather it
    const double eps = 2.220446049250313e-16;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const int tid = threadIdx.x, nthr = kMainThreads;
    long long nhist = 0;

    // vectors cross the ABI in caller order; internally rows follow the reordered matrix
    double *xi = MR ? xbase + n : a.work + 6 * (size_t)n;   // Δx in internal order, gathered by all CTAs
    halo_arm(hp, a, gr.gen);
    for (int row = r0 + tid; row < r1; row += nthr) {
        const int src = a.perm[row];
        const double x0 = a.x[src];                    // Δx, the warm start
        xl[row] = x0;
        xi[row] = x0;
        halo_put(hp, xi + row, row, x0);
        dl[row] = a.dinv ? a.dinv[src] : a.pscale;
    }
    gr.barrier();
    // r = b − A Δx ; z = M r ; p = z ; γ = r·z          (own rows; z is never stored)
    double part = 0.0;
    halo_unpack(hp, a, gr.gen, &mailbox, xi);
    halo_arm(hp, a, gr.gen);
    eng.run(xi, [&](int row, double ax) {
        const double rv = a.b[a.perm[row]] - ax;
        const double zv = dl[row] * rv;
        r[row] = rv;
        pl[row] = zv;
        pg[row] = zv;
        halo_put(hp, pg + row, row, zv);
        part = fma(rv, zv, part);
    });
    double gamma = gr.sum_threads<true>(part);      // also publishes p for the first SpMV
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
    // straight onto Δx: the same sum up to the rounding of one addition per entry.
    while (!(solved || tired || zero_curv || gr.aborted())) {
        // Ap = A p ; pAp = p·Ap
        part = 0.0;
        halo_unpack(hp, a, gr.gen, &mailbox, pg);
        eng.run(pg, [&](int row, double ap) {
            Ap[row] = ap;
            part = fma(pl[row], ap, part);
        });
        const double pAp = gr.sum_threads<false>(part);
        if (pAp <= eps * pnorm2 && fabs(pAp) <= eps * pnorm2) {
            zero_curv = true;
            continue;
        }
        const double alpha = gamma / pAp;
        // x += α p ; r −= α Ap ; z = M r ; γ⁺ = r·z       (own rows, thread-per-row)
        part = 0.0;
        for (int row = r0 + tid; row < r1; row += nthr) {
            xl[row] = fma(alpha, pl[row], xl[row]);
            const double rv = fma(-alpha, Ap[row], r[row]);
            r[row] = rv;
            part = fma(rv, dl[row] * rv, part);
        }
        const double gamma_next = gr.sum_threads<false>(part);
        rnorm = sqrt(gamma_next);
        if (lead && a.hist && nhist < a.hist_cap) a.hist[nhist] = rnorm;
        nhist++;
        solved = (rnorm <= tol) || (rnorm + 1.0 <= 1.0);
        if (!solved) {
            const double beta = gamma_next / gamma;
            pnorm2 = gamma_next + beta * beta * pnorm2;
            gamma = gamma_next;
            // p = z + β p   (own rows), then publish p for the next SpMV gather
            halo_arm(hp, a, gr.gen);
            for (int row = r0 + tid; row < r1; row += nthr) {
                const double pv = fma(beta, pl[row], dl[row] * r[row]);
                pl[row] = pv;
                pg[row] = pv;
                halo_put(hp, pg + row, row, pv);
            }
            gr.barrier();
        }
        iter++;
        tired = iter >= itmax;
    }
    if constexpr (MR) {
        // all-gather of the solution: every rank receives every row, then scatters all rows to
        // caller order, so x is complete (replicated) on every rank when the solve returns
        for (int row = r0 + tid; row < r1; row += nthr) {
            const double xv = xl[row];
            xi[row] = xv;
            hp.put_all(xi + row, xv);
        }
        gr.remote = true;
        gr.barrier();
        for (int row = blockIdx.x * nthr + tid; row < n; row += gridDim.x * nthr) a.x[a.perm[row]] = ld_cg(xi + row);
        gr.barrier();      // no rank leaves (and lets the next solve reuse the arena) before all have read it
    } else {
        for (int row = r0 + tid; row < r1; row += nthr) a.x[a.perm[row]] = xl[row];
    }
    if (lead) a.result[13] = (double)gr.gen;
    gr.finish();
    eng.finish();
    if (lead) {
        a.result[0] = (double)iter;
        a.result[1] = solved ? 1.0 : 0.0;
        a.result[2] = zero_curv ? 1.0 : 0.0;
        a.result[3] = 0.0;
        a.result[4] = rnorm;
        a.result[5] = rnorm0;
        a.result[6] = (double)(nhist < a.hist_cap ? nhist : a.hist_cap);
        a.result[7] = gr.aborted() ? 1.0 : 0.0;
    }
}

// =============================================================================================
// GMRES(m), restarted, left preconditioned  (Krylov.jl gmres!)
// =============================================================================================

// Coarse phase timer (CTA 0, thread 0 only): cycles spent in SpMV / local vector work / waiting
// for grid reductions / scalar recurrences, reported per solve for the roofline analysis.
template <bool ENABLED>
struct PhaseClock;
template <>
struct PhaseClock<false> {
    __device__ __forceinline__ void start(bool) {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void dump(double *) const {}
};
template <>
struct PhaseClock<true> {
    long long last, acc[4], c0;
    unsigned long long ns0;
    bool on;
    __device__ __forceinline__ void dump(double *result) const {
        if (!on) return;
        for (int i = 0; i < 4; ++i) result[8 + i] = (double)acc[i];
        result[12] = mhz();
    }
    __device__ __forceinline__ void start(bool enable) {
        on = enable;
        acc[0] = acc[1] = acc[2] = acc[3] = 0;
        last = c0 = on ? clock64() : 0;
        ns0 = on ? global_timer_ns() : 0;
    }
    // SM clock actually seen by this kernel (cycles per nanosecond -> MHz)
    __device__ __forceinline__ double mhz() const {
        const unsigned long long ns = global_timer_ns() - ns0;
        return ns ? 1e3 * (double)(clock64() - c0) / (double)ns : 0.0;
    }
    __device__ __forceinline__ void mark(int phase) {
        if (on) {
            const long long now = clock64();
            acc[phase] += now - last;
            last = now;
        }
    }
};

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

__device__ __forceinline__ double precond(const KrylovArgs &a, int row, double v) {
    return a.dinv ? __ldg(a.dinv + __ldg(a.perm + row)) * v : a.pscale * v;
}

template <int T, bool RES, bool PROF, bool MR>
__global__ void __launch_bounds__(kThreads, 1) k_gmres(const __grid_constant__ KrylovArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ CommMailbox mailbox;
    __shared__ double sc[kMaxMemory], ss[kMaxMemory], sz[kMaxMemory + 1], sy[kMaxMemory + 1];
    __shared__ double sR[kMaxMemory * (kMaxMemory + 1) / 2];
    __shared__ double s_flags[4];      // rnorm, inconsistent, -, Hbis
    __shared__ double s_seg[32];
    __shared__ double s_dot[kMaxMemory + 1][kMainWarps];       // per-warp partial dots of the register-resident CGS2 form
    // fused CGS2: unrotated Hessenberg columns (packed: column j holds rows 0..j+1 at j(j+3)/2),
    // the second-pass coefficients h₂, the correction H̄ h₂ / H and per-warp partial norms
    __shared__ double sHbar[kMaxMemory * (kMaxMemory + 3) / 2], s_h2[kMaxMemory], s_corr[kMaxMemory + 1], s_wsum[kMainWarps];
    __shared__ int s_prange[2 * kMaxRanks];
    __shared__ char *s_pdest[kMaxRanks];

    if (threadIdx.x == 0) mailbox.dead = 0;
    const int lb = a.rank * gridDim.x + blockIdx.x;       // position in the partition (all ranks)
    const int r0 = a.part[lb], r1 = a.part[lb + 1];
    SpmvEngine<T, RES> eng(a, r0, r1, dyn_smem, &mailbox);
    HaloPush<MR> hp;
    hp.init(a, r0, r1, s_prange, s_pdest);
    __syncthreads();
    if (threadIdx.x >= kMainThreads) {                     // comm warps: exchange service only
        comm_warp_loop<MR>(&mailbox, a, a.poll_depth, dyn_smem);
        return;
    }
    GridReduce gr;
    gr.init(&mailbox, a.trace, pushes_remote(hp));
    double *sm_in = mailbox.in, *sm_out = mailbox.out;
    const int n = a.n;
    const int mem = a.mem;
    VecSlices V(a, dyn_smem, a.work, r0);                // V.at(i), i = 0..mem: CTA-local basis rows
    // vectors the other CTAs gather: in the exchange arena when the solve is sharded
    double *xbase = MR ? reinterpret_cast<double *>(a.arena[a.rank] + kArenaVecOffset) : a.work + (size_t)(mem + 1) * n;
    double *qbuf = xbase;                                // two raw buffers for the SpMV gather (global)
    double *x = xbase + 2 * (size_t)n;                   // iterate in internal order (gathered by all CTAs)
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const int tid = threadIdx.x, nthr = kMainThreads;
    const int lane = tid & 31, wid = tid >> 5, nwarps = kMainWarps;
    // vectors cross the ABI in caller order; internally rows follow the reordered matrix
    halo_arm(hp, a, gr.gen);
    for (int row = r0 + tid; row < r1; row += nthr) {
        const double xv = a.x[a.perm[row]];
        x[row] = xv;
        halo_put(hp, x + row, row, xv);
    }
    gr.barrier();
    const double btol = 1.8189894035458565e-12;          // eps^(3/4)
    long long nhist = 0;
    PhaseClock<PROF> pc;
    pc.start(lead);

    // ---- initial residual: w = b − A x0 ; r0 = M w (raw into qbuf[0]) ; β = ‖r0‖
    int cur = 0;
    double beta;
    {
        double part = 0.0;
        double *q0 = qbuf;
        double *v0 = V.at(0);
        halo_unpack(hp, a, gr.gen, &mailbox, x);
        halo_arm(hp, a, gr.gen);
        eng.run(x, [&](int row, double ax) {
            const double v = precond(a, row, a.b[a.perm[row]] - ax);
            q0[row] = v;
            halo_put(hp, q0 + row, row, v);
            v0[row] = v;
            part = fma(v, v, part);
        });
        beta = sqrt(gr.sum_threads<true>(part));
    }
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

    const bool fused = a.orth == NUPGCM_ORTH_CGS2_FUSED;
    constexpr int QCAP = RES ? 0 : kArenaEntries;      // rows of the new vector the idle footprint arena can hold
    // Givens / least-squares update of one Arnoldi column (Krylov.jl order), by ONE thread
    auto scalar_step = [&](int k, int nr, double hsq) {
        const double Hbis = sqrt(hsq);
        // previous rotations applied to the new column; the running entry stays in a register so that
        // the loop-carried dependence is two arithmetic ops per rotation, not a shared-memory round trip
        double ri = sR[nr];
        for (int i = 0; i < k - 1; ++i) {
            const double ci = sc[i], si = ss[i], rn = sR[nr + i + 1];
            sR[nr + i] = ci * ri + si * rn;
            ri = si * ri - ci * rn;
        }
        double c, s, rho;
        sym_givens(ri, Hbis, c, s, rho);
        sc[k - 1] = c;
        ss[k - 1] = s;
        sR[nr + k - 1] = rho;
        const double zeta = s * sz[k - 1];
        sz[k - 1] = c * sz[k - 1];
        sz[k] = zeta;                 // only consumed if the pass continues
        s_flags[0] = fabs(zeta);
        s_flags[3] = Hbis;
    };
    while (!(solved || tired || breakdown || gr.aborted())) {
        // ---- start of a pass ----
        if (tid < kMaxMemory) { sc[tid] = 0.0; ss[tid] = 0.0; }
        if (tid <= kMaxMemory) sz[tid] = 0.0;
        for (int i = tid; i < kMaxMemory * (kMaxMemory + 1) / 2; i += nthr) sR[i] = 0.0;
        if (npass >= 1) {
            // w = b − A x ; r0 = M w ; β = ‖r0‖   (x was published by the barrier ending the last pass)
            double part = 0.0;
            double *q0 = qbuf + (size_t)cur * n;
            double *v0 = V.at(0);
            halo_unpack(hp, a, gr.gen, &mailbox, x);
            halo_arm(hp, a, gr.gen);
            eng.run(x, [&](int row, double ax) {
                const double v = precond(a, row, a.b[a.perm[row]] - ax);
                q0[row] = v;
                halo_put(hp, q0 + row, row, v);
                v0[row] = v;
                part = fma(v, v, part);
            });
            beta = sqrt(gr.sum_threads<true>(part));
        }
        main_sync();
        if (tid == 0) sz[0] = beta;
        // V[0] = r0/β on own rows; other CTAs read the raw vector scaled by inv_h
        double inv_h = 1.0 / beta;
        {
            double *v0 = V.at(0);
            for (int row = r0 + tid; row < r1; row += nthr) v0[row] *= inv_h;
        }
        npass++;
        int k = 0;          // inner_iter
        int nr = 0;
        bool inner_tired = false;
        bool have_corr = false;     // fused CGS2: the exchanged vector is q₁, corrected after the SpMV
        while (!(solved || inner_tired || breakdown || gr.aborted())) {
            k++;
            // ---- q = M A v_k on own rows (raw v_k gathered from qbuf[cur], scaled by inv_h)
            double *q = V.at(k);                          // slot of the next basis vector
            const double *src = qbuf + (size_t)cur * n;
            double *dst = qbuf + (size_t)(cur ^ 1) * n;
            pc.mark(3);
            halo_unpack(hp, a, gr.gen, &mailbox, src);
            eng.run(src, [&](int row, double av) { q[row] = precond(a, row, av * inv_h); });
            main_sync();
            pc.mark(0);
            double hsq = 0.0;                            // ‖q‖² after orthogonalisation
            bool normalised = false;                     // V[k] already holds q / H
            if (a.orth == NUPGCM_ORTH_MGS) {
                // h_i = v_i·q ; q −= h_i v_i, sequentially (one grid reduction per i)
                constexpr int RMAX = MR ? 4 : 7;
                if (r1 - r0 <= RMAX * nthr) {
                    // Register-resident form (up to RMAX rows per thread): the thread keeps its rows of q and of
                    // the current basis vector in registers and fetches its rows of v_{i+1} while the
                    // reduction of h_i is in flight, so that between two reductions there is only register
                    // arithmetic (one fma per row, then the next partial dot) — k+1 reductions in a row are what
                    // modified Gram-Schmidt costs, nothing should be added to them.  Same operations on the
                    // same rows in the same order as the streaming form below: bit-identical results.
                    double qr[RMAX], vc[RMAX], vn[RMAX];
                    const double *v0p = V.at(0);
#pragma unroll
                    for (int j = 0; j < RMAX; ++j) {
                        const int row = r0 + tid + j * nthr;
                        qr[j] = row < r1 ? q[row] : 0.0;
                        vc[j] = row < r1 ? v0p[row] : 0.0;
                    }
                    for (int i = 0; i < k; ++i) {
                        double part = 0.0;
#pragma unroll
                        for (int j = 0; j < RMAX; ++j) part = fma(vc[j], qr[j], part);
                        if (i + 1 < k) {
                            const double *vnp = V.at(i + 1);
#pragma unroll
                            for (int j = 0; j < RMAX; ++j) {
                                const int row = r0 + tid + j * nthr;
                                vn[j] = row < r1 ? vnp[row] : 0.0;
                            }
                        }
                        pc.mark(1);
                        const double h = gr.sum_threads<false>(part);
                        pc.mark(2);
                        if (tid == 0) sR[nr + i] = h;
#pragma unroll
                        for (int j = 0; j < RMAX; ++j) {
                            qr[j] = fma(-h, vc[j], qr[j]);
                            vc[j] = vn[j];
                        }
                    }
                    double part = 0.0;
                    halo_arm(hp, a, gr.gen);
#pragma unroll
                    for (int j = 0; j < RMAX; ++j) {
                        const int row = r0 + tid + j * nthr;
                        if (row < r1) {
                            dst[row] = qr[j];
                            halo_put(hp, dst + row, row, qr[j]);
                            part = fma(qr[j], qr[j], part);
                        }
                    }
                    pc.mark(1);
                    hsq = gr.sum_threads<true>(part);     // publishes dst for the next gather
                    pc.mark(2);
                    // v_{k+1} = q / H straight from the registers (H is known to every thread now): saves the
                    // store of the raw q and the normalisation sweep over it further down
                    const double ih = 1.0 / sqrt(hsq);
#pragma unroll
                    for (int j = 0; j < RMAX; ++j) {
                        const int row = r0 + tid + j * nthr;
                        if (row < r1) q[row] = qr[j] * ih;
                    }
                    normalised = true;
                } else {
                // Streaming form (more rows per thread than registers hold).  In the streamed-matrix kernels the
                // footprint arena is idle between two SpMVs: the CTA's first `qcap` rows of the new vector live
                // there through the k+1 steps (read and written once per step where the global form does two
                // accesses each), and v_{k+1} = q / H is written to the basis straight from it.
                double *qs = QCAP ? reinterpret_cast<double *>(dyn_smem + a.lay.st_xs) : nullptr;
                const int qcap = QCAP && a.lay.streaming ? QCAP : 0;
                double hprev = 0.0;
                for (int i = 0; i < k; ++i) {
                    const double *vi = V.at(i);
                    const double *vp = V.at(i > 0 ? i - 1 : 0);
                    double part = 0.0;
#pragma unroll 8
                    for (int row = r0 + tid; row < r1; row += nthr) {
                        const bool in_s = row - r0 < qcap;
                        double qv = (i > 0 && in_s) ? qs[row - r0] : q[row];
                        if (i > 0) qv = fma(-hprev, vp[row], qv);
                        if (in_s) qs[row - r0] = qv;
                        else if (i > 0) q[row] = qv;
                        part = fma(vi[row], qv, part);
                    }
                    // the next step reads v_{i+1} (and v_i again): have the former on its way into L2 while the
                    // reduction is in flight — the basis does not fit in L2 at the sizes that take this form
                    if (i + 1 < k && (lane & 15) == 0) {
                        const double *vn = V.at(i + 1);
                        for (int row = r0 + tid; row < r1; row += nthr)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(vn + row));
                    }
                    pc.mark(1);
                    hprev = gr.sum_threads<false>(part);
                    pc.mark(2);
                    if (tid == 0) sR[nr + i] = hprev;
                }
                const double *vp = V.at(k - 1);
                double part = 0.0;
                halo_arm(hp, a, gr.gen);
                for (int row = r0 + tid; row < r1; row += nthr) {
                    const bool in_s = row - r0 < qcap;
                    const double qv = fma(-hprev, vp[row], in_s ? qs[row - r0] : q[row]);
                    if (in_s) qs[row - r0] = qv;
                    else q[row] = qv;
                    dst[row] = qv;
                    halo_put(hp, dst + row, row, qv);
                    part = fma(qv, qv, part);
                }
                pc.mark(1);
                hsq = gr.sum_threads<true>(part);     // publishes dst for the next gather
                pc.mark(2);
                if (qcap > 0) {
                    const double ih = 1.0 / sqrt(hsq);
                    for (int row = r0 + tid; row < r1; row += nthr) q[row] = (row - r0 < qcap ? qs[row - r0] : q[row]) * ih;
                    normalised = true;
                }
                }
            } else if (a.orth == NUPGCM_ORTH_CGS2_FUSED) {
                // CGS2 with two grid reductions per iteration instead of three:
                //   1. h₁ = Vᵀq ;            q₁ = q − V h₁ — q₁ is what the other CTAs will gather
                //   2. [h₂ = Vᵀq₁, ‖q₁‖²] in ONE reduction that also publishes q₁;
                //      ‖q₂‖² = ‖q₁‖² − ‖h₂‖² (Pythagoras; h₂ = O(ε)‖q₁‖), q₂ = q₁ − V h₂ locally.
                // The next SpMV runs on the published q₁ instead of q₂ = q₁ − V h₂.  By the Arnoldi relation
                // (Â V_k = V_{k+1} H̄_k) that adds V_{k+1} c, c = H̄_k h₂ / H, to the operator's output: a
                // vector INSIDE the span the next projection removes.  So q₁ of the next iteration is
                // unchanged and only its coefficients need the correction h = Vᵀq̂ − c (VᵀV = I up to
                // O(ε), c = O(ε): the neglected term is O(ε²)).  Equal to CGS2 in exact arithmetic.
                constexpr int RF = MR ? 4 : 8;
                // (streamed-matrix kernels only: with the basis in shared memory the segment form below is faster —
                // 18.7 against 21.6 us per iteration at h = 0.08 — because its k projections run on different warps
                // while this form walks them in sequence, two warp reductions per pair of basis vectors)
                if (!RES && r1 - r0 <= RF * (nthr - 32)) {
                    // Register-resident form (up to RF rows per thread of warps 1..10; warp 0 owns no rows and does
                    // the scalar recurrences while the others finish the vector): the thread keeps its rows of q
                    // in registers through both projections and both updates, so each of the four passes reads
                    // the basis rows once, with all loads of a pass independent, and q is read and written once
                    // per iteration instead of eight times.
                    const int t1 = tid - 32, nt = nthr - 32;
                    double qr[RF];
#pragma unroll
                    for (int j = 0; j < RF; ++j) {
                        const int row = r0 + t1 + j * nt;
                        qr[j] = (t1 >= 0 && row < r1) ? q[row] : 0.0;
                    }
                    // (the basis vectors are taken two at a time so that 2 x RF loads are in flight per thread)
                    auto dots = [&](int extra, double extra_val) {       // sm_in[i] <- this CTA's part of v_i . q (i < k)
                        for (int i = 0; i < k; i += 2) {
                            const double *va = V.at(i), *vb = V.at(i + 1 < k ? i + 1 : i);
                            double xa[RF], xb[RF];
#pragma unroll
                            for (int j = 0; j < RF; ++j) {
                                const int row = r0 + t1 + j * nt;
                                const bool ok = t1 >= 0 && row < r1;
                                xa[j] = ok ? va[row] : 0.0;
                                xb[j] = ok ? vb[row] : 0.0;
                            }
                            double pa = 0.0, pb = 0.0;
#pragma unroll
                            for (int j = 0; j < RF; ++j) {
                                pa = fma(xa[j], qr[j], pa);
                                pb = fma(xb[j], qr[j], pb);
                            }
                            pa = warp_sum(pa);
                            pb = warp_sum(pb);
                            if (lane == 0) {
                                s_dot[i][wid] = pa;
                                if (i + 1 < k) s_dot[i + 1][wid] = pb;
                            }
                        }
                        if (extra) {
                            const double e = warp_sum(extra_val);
                            if (lane == 0) s_dot[k][wid] = e;
                        }
                        main_sync();
                        if (tid < k + extra) {
                            double t = 0.0;
                            for (int wv = 1; wv < nwarps; ++wv) t += s_dot[tid][wv];
                            sm_in[tid] = t;
                        }
                        main_sync();
                    };
                    auto update = [&]() {                                // q <- q - V h, h = sm_out[0..k)
                        for (int i = 0; i < k; i += 2) {
                            const double *va = V.at(i), *vb = V.at(i + 1 < k ? i + 1 : i);
                            const double ha = sm_out[i], hb = i + 1 < k ? sm_out[i + 1] : 0.0;
                            double xa[RF], xb[RF];
#pragma unroll
                            for (int j = 0; j < RF; ++j) {
                                const int row = r0 + t1 + j * nt;
                                const bool ok = t1 >= 0 && row < r1;
                                xa[j] = ok ? va[row] : 0.0;
                                xb[j] = ok ? vb[row] : 0.0;
                            }
#pragma unroll
                            for (int j = 0; j < RF; ++j) qr[j] = fma(-hb, xb[j], fma(-ha, xa[j], qr[j]));
                        }
                    };
                    dots(0, 0.0);
                    pc.mark(1);
                    gr.sumN(k);
                    pc.mark(2);
                    update();
                    double part = 0.0;
                    halo_arm(hp, a, gr.gen);
#pragma unroll
                    for (int j = 0; j < RF; ++j) {
                        const int row = r0 + t1 + j * nt;
                        if (t1 >= 0 && row < r1) {
                            dst[row] = qr[j];
                            halo_put(hp, dst + row, row, qr[j]);
                            part = fma(qr[j], qr[j], part);
                        }
                    }
                    if (tid < k) sR[nr + tid] = have_corr ? sm_out[tid] - s_corr[tid] : sm_out[tid];
                    dots(1, part);
                    pc.mark(1);
                    gr.sumN(k + 1, true);                          // also publishes dst (= q₁)
                    pc.mark(2);
                    double h2sq = 0.0;
                    for (int i = 0; i < k; ++i) h2sq = fma(sm_out[i], sm_out[i], h2sq);
                    hsq = fmax(sm_out[k] - h2sq, 0.0);
                    const double Hb = sqrt(hsq);
                    if (wid == 0) {
                        const int hc = (k - 1) * (k + 2) / 2;      // packed offset of H̄ column k-1
                        if (lane < k) {
                            const double h = sR[nr + lane] + sm_out[lane];
                            sR[nr + lane] = h;
                            s_h2[lane] = sm_out[lane];
                            sHbar[hc + lane] = h;
                        } else if (lane == k) {
                            sHbar[hc + k] = Hb;
                        }
                        __syncwarp();
                        if (lane <= k) {
                            double c = 0.0;
                            for (int j = lane > 0 ? lane - 1 : 0; j < k; ++j) c = fma(sHbar[j * (j + 3) / 2 + lane], s_h2[j], c);
                            s_corr[lane] = Hb > 0.0 ? c / Hb : 0.0;
                        }
                        __syncwarp();
                        if (lane == 0) scalar_step(k, nr, hsq);
                    } else {
                        update();                                  // q₂ = q₁ − V h₂ ; v_{k+1} = q₂ / H
                        const double ih = 1.0 / Hb;
#pragma unroll
                        for (int j = 0; j < RF; ++j) {
                            const int row = r0 + t1 + j * nt;
                            if (row < r1) q[row] = qr[j] * ih;
                        }
                    }
                    have_corr = true;
                } else {
                const int nseg = nwarps / k > 0 ? nwarps / k : 1;
                const int rows = r1 - r0;
                const int seglen = (rows + nseg - 1) / nseg;
                auto project = [&]() {                         // s_seg <- per-segment partial dots
                    for (int u = wid; u < k * nseg; u += nwarps) {
                        const int i = u % k, sg = u / k;
                        const double *vi = V.at(i);
                        const int lo = r0 + sg * seglen, hi = min(r1, lo + seglen);
                        double part = 0.0;
                        for (int row = lo + lane; row < hi; row += 32) part = fma(vi[row], q[row], part);
                        part = warp_sum(part);
                        if (lane == 0) s_seg[u] = part;
                    }
                };
                main_sync();
                project();
                main_sync();
                if (tid < k) {
                    double t = 0.0;
                    for (int sg = 0; sg < nseg; ++sg) t += s_seg[sg * k + tid];
                    sm_in[tid] = t;
                }
                main_sync();
                pc.mark(1);
                gr.sumN(k);
                pc.mark(2);
                double part = 0.0;
                halo_arm(hp, a, gr.gen);
                for (int row = r0 + tid; row < r1; row += nthr) {
                    double qv = q[row];
                    for (int i = 0; i < k; ++i) qv = fma(-sm_out[i], V.at(i)[row], qv);
                    q[row] = qv;
                    dst[row] = qv;
                    halo_put(hp, dst + row, row, qv);
                    part = fma(qv, qv, part);
                }
                if (tid < k) sR[nr + tid] = have_corr ? sm_out[tid] - s_corr[tid] : sm_out[tid];
                part = warp_sum(part);
                if (lane == 0) s_wsum[wid] = part;
                main_sync();
                project();
                main_sync();
                if (tid < k) {
                    double t = 0.0;
                    for (int sg = 0; sg < nseg; ++sg) t += s_seg[sg * k + tid];
                    sm_in[tid] = t;
                } else if (tid == k) {
                    double t = 0.0;
                    for (int wv = 0; wv < nwarps; ++wv) t += s_wsum[wv];
                    sm_in[k] = t;
                }
                main_sync();
                pc.mark(1);
                gr.sumN(k + 1, true);                          // also publishes dst (= q₁)
                pc.mark(2);
                double h2sq = 0.0;
                for (int i = 0; i < k; ++i) h2sq = fma(sm_out[i], sm_out[i], h2sq);
                hsq = fmax(sm_out[k] - h2sq, 0.0);
                const double Hb = sqrt(hsq);
                if (wid == 0) {
                    // warp 0: Hessenberg column, correction for the next iteration, Givens update —
                    // while the other warps update and normalise the vector
                    const int hc = (k - 1) * (k + 2) / 2;      // packed offset of H̄ column k-1
                    if (lane < k) {
                        const double h = sR[nr + lane] + sm_out[lane];
                        sR[nr + lane] = h;
                        s_h2[lane] = sm_out[lane];
                        sHbar[hc + lane] = h;
                    } else if (lane == k) {
                        sHbar[hc + k] = Hb;
                    }
                    __syncwarp();
                    if (lane <= k) {
                        double c = 0.0;
                        for (int j = lane > 0 ? lane - 1 : 0; j < k; ++j) c = fma(sHbar[j * (j + 3) / 2 + lane], s_h2[j], c);
                        s_corr[lane] = Hb > 0.0 ? c / Hb : 0.0;
                    }
                    __syncwarp();
                    if (lane == 0) scalar_step(k, nr, hsq);
                } else {
                    // q₂ = q₁ − V h₂ and v_{k+1} = q₂ / H in one sweep (H is known to every thread)
                    const double ih = 1.0 / Hb;
                    for (int row = r0 + tid - 32; row < r1; row += nthr - 32) {
                        double qv = q[row];
                        for (int i = 0; i < k; ++i) qv = fma(-sm_out[i], V.at(i)[row], qv);
                        q[row] = qv * ih;
                    }
                }
                have_corr = true;
                }
            } else {
                // CGS2: all k projections at once, twice; warp w handles basis vector w % k on
                // row segment w / k.
                const int nseg = nwarps / k > 0 ? nwarps / k : 1;
                const int rows = r1 - r0;
                const int seglen = (rows + nseg - 1) / nseg;
                for (int pass = 0; pass < 2; ++pass) {
                    main_sync();
                    for (int u = wid; u < k * nseg; u += nwarps) {
                        const int i = u % k, sg = u / k;
                        const double *vi = V.at(i);
                        const int lo = r0 + sg * seglen, hi = min(r1, lo + seglen);
                        double part = 0.0;
                        for (int row = lo + lane; row < hi; row += 32) part = fma(vi[row], q[row], part);
                        part = warp_sum(part);
                        if (lane == 0) s_seg[u] = part;        // k*nseg <= nwarps <= 32
                    }
                    main_sync();
                    if (tid < k) {
                        double t = 0.0;
                        for (int sg = 0; sg < nseg; ++sg) t += s_seg[sg * k + tid];
                        sm_in[tid] = t;
                    }
                    main_sync();
                    pc.mark(1);
                    gr.sumN(k);
                    pc.mark(2);
                    // q −= Σ h_i v_i  (own rows)
                    double part = 0.0;
                    halo_arm(hp, a, gr.gen);
                    for (int row = r0 + tid; row < r1; row += nthr) {
                        double qv = q[row];
                        for (int i = 0; i < k; ++i) qv = fma(-sm_out[i], V.at(i)[row], qv);
                        q[row] = qv;
                        if (pass == 1) { dst[row] = qv; halo_put(hp, dst + row, row, qv); part = fma(qv, qv, part); }
                    }
                    if (tid < k) sR[nr + tid] = (pass == 0) ? sm_out[tid] : sR[nr + tid] + sm_out[tid];
                    if (pass == 1) {
                        pc.mark(1);
                        hsq = gr.sum_threads<true>(part);
                        pc.mark(2);
                    }
                }
            }
            // ---- scalar recurrences, replicated per CTA (one thread), Krylov.jl order
            if (!fused) {
                main_sync();
                if (tid == 0) scalar_step(k, nr, hsq);
            }
            main_sync();
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
                if (!fused && !normalised)
                    for (int row = r0 + tid; row < r1; row += nthr) q[row] *= inv_h;
                cur ^= 1;
            }
        }
        // ---- back substitution R y = z (thread 0), then x += Σ y_i v_i on own rows
        main_sync();
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
        main_sync();
        if (s_flags[1] != 0.0) inconsistent = true;
        halo_arm(hp, a, gr.gen);
        for (int row = r0 + tid; row < r1; row += nthr) {
            double xr = 0.0;
            for (int i = 0; i < k; ++i) xr = fma(sy[i], V.at(i)[row], xr);
            const double xv = x[row] + xr;
            x[row] = xv;
            halo_put(hp, x + row, row, xv);
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
    if constexpr (MR) {
        // all-gather of the solution (see k_cg)
        for (int row = r0 + tid; row < r1; row += nthr) hp.put_all(x + row, x[row]);
        gr.remote = true;
        gr.barrier();
        for (int row = blockIdx.x * nthr + tid; row < n; row += gridDim.x * nthr) a.x[a.perm[row]] = ld_cg(x + row);
        gr.barrier();
    } else {
        for (int row = r0 + tid; row < r1; row += nthr) a.x[a.perm[row]] = x[row];
    }
    if (lead) a.result[13] = (double)gr.gen;
    gr.finish();
    eng.finish();
    if (lead) {
        a.result[0] = (double)iter;
        a.result[1] = solved ? 1.0 : 0.0;
        a.result[2] = inconsistent ? 1.0 : 0.0;
        a.result[3] = breakdown ? 1.0 : 0.0;
        a.result[4] = rnorm;
        a.result[5] = rnorm0;
        a.result[6] = (double)(nhist < a.hist_cap ? nhist : a.hist_cap);
        a.result[7] = gr.aborted() ? 1.0 : 0.0;
        pc.mark(3);
        pc.dump(a.result);
    }
}

// =============================================================================================
// host side
// =============================================================================================

// replicas of the reduction slots (NUPGCM_REPLICAS overrides); measured on B200 with
// tools/reduce_cliff.py
static int poll_config(int grid) {
    // all 148 SMs spinning on the same ~19 cache lines sit past a knee of the L2 slices that hold them: at
    // h = 0.04 (tools/reduce_sweep.py, profiles/reduce_sweep_r02.txt) MGS runs at 123 / 107 / 101 us per iteration
    // with 1 / 2 / 4 replicas on 148 CTAs (8: 100; h = 0.08: 54 / 49 / 48 with 2 / 4 / 8), and at 103 / 100 / 102 on 140 CTAs
    int rep = grid > 100 ? kMaxReplicas : 1;
    if (const char *er = getenv("NUPGCM_REPLICAS")) {
        const int v = atoi(er);
        if (v >= 1 && v <= kMaxReplicas) rep = v;
    }
    return rep;
}

static int pow2floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// Lanes per row inside the persistent kernels: the row-length choice, narrowed when the CTA has
// fewer rows than row groups so that all rows of the CTA are processed in one sweep.
static int persistent_tpr(const nupgcm_csr *A, bool resident, int grid) {
    const char *env = getenv("NUPGCM_TPR");
    if (env) {
        int v = atoi(env);
        if (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) return v;
    }
    const int rows_per_cta = (int)((A->n_rows + grid - 1) / grid);
    int t = A->tpr;
    if (resident) return std::min(t, 8);       // measured on B200: 4-8 lanes per row are fastest
    if (rows_per_cta > 0 && rows_per_cta * t > kMainThreads && rows_per_cta <= kMainThreads / 2)
        t = std::max(2, pow2floor(kMainThreads / rows_per_cta));
    return std::min(t, 32);
}

static int32_t ensure_workspace(nupgcm_ctx *ctx, size_t bytes);

// Sharded solves must not meet a device-wide synchronisation (cudaFree of a grown workspace)
// between the launches of two ranks: reserve the largest workspace any solve over the
// communicator can need when a matrix is sharded.
int32_t nupgcm_reserve_solver_workspace(nupgcm_ctx *ctx, int64_t max_n) {
    return ensure_workspace(ctx, (size_t)(kMaxMemory + 4) * (size_t)max_n * sizeof(double));
}

static int32_t ensure_workspace(nupgcm_ctx *ctx, size_t bytes) {
    if (ctx->ws_bytes >= bytes) return NUPGCM_OK;
    if (ctx->d_ws) {
        NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_ws);
        ctx->d_ws = nullptr;
        ctx->ws_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->d_ws, bytes);
    if (e != cudaSuccess)
        return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "device allocation failed: %s", cudaGetErrorString(e));
    ctx->ws_bytes = bytes;
    return NUPGCM_OK;
}

template <int T, bool RES>
static cudaError_t launch2(bool gmres, KrylovArgs &args, nupgcm_ctx *ctx, size_t smem, int grid) {
    void *params[] = {&args};
    const char *ep = getenv("NUPGCM_PROFILE");
    const bool prof = ep && atoi(ep) != 0;
    const bool mr = args.nranks > 1;
    const void *fn;
    if (gmres) {
        if (mr) fn = (const void *)k_gmres<T, RES, false, true>;
        else fn = prof ? (const void *)k_gmres<T, RES, true, false> : (const void *)k_gmres<T, RES, false, false>;
    } else {
        fn = mr ? (const void *)k_cg<T, RES, true> : (const void *)k_cg<T, RES, false>;
    }
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // cooperative launch = co-residency guarantee (the kernels never call grid.sync()); NUPGCM_COOP=0
    // uses a plain launch, for experiments with several contexts sharing a device
    const char *ec = getenv("NUPGCM_COOP");
    if (ec && atoi(ec) == 0) return cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), params, smem, ctx->stream);
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), params, smem, ctx->stream);
}

template <int T>
static cudaError_t launch(bool gmres, KrylovArgs &args, nupgcm_ctx *ctx, size_t smem, int grid, bool resident) {
    return resident ? launch2<T, true>(gmres, args, ctx, smem, grid) : launch2<T, false>(gmres, args, ctx, smem, grid);
}

// Shared-memory plan of the SM-resident form; total == 0 when the matrix slice does not fit.
static ResidentLayout plan_resident(const nupgcm_csr *A, bool gmres, int memory) {
    ResidentLayout none;
    memset(&none, 0, sizeof(none));
    const char *env = getenv("NUPGCM_RESIDENT");
    if (env && atoi(env) == 0) return none;
    if (!A->d_loc || A->res_max_nnz == 0) return none;
    const int static_smem = gmres ? 8704 : 4096;          // the kernels' static shared memory
    const int limit = 227 * 1024 - static_smem;
    const char *envv = getenv("NUPGCM_VEC_SMEM");
    const int n_vec = (envv && atoi(envv) == 0) ? 1 << 20 : (gmres ? memory + 1 : 5);
    ResidentLayout L = resident_layout(A->res_max_nnz, A->res_max_foot, A->res_max_rows, n_vec, limit);
    return L.total <= limit ? L : none;
}

// Shared-memory plan of the streaming form (tiled TMA streams); total == 0 when the matrix has no
// streaming tables (then the legacy direct-load loop runs).
static ResidentLayout plan_streaming(const nupgcm_csr *A, bool gmres, int memory) {
    ResidentLayout L;
    memset(&L, 0, sizeof(L));
    const char *env = getenv("NUPGCM_STREAM_TMA");
    if (env && atoi(env) == 0) return L;
    if (!A->d_svals || A->str_fmax <= 0) return L;
    const int static_smem = gmres ? 8704 : 4096;
    const int limit = 227 * 1024 - static_smem;
    int off = 0;
    L.st_ring_v = off; off += kMainWarps * kRingEntries * 8;
    L.st_ring_c = off; off += kMainWarps * kRingEntries * 2;
    L.st_bars = off;   off += ((kMainWarps * kRingPieces + 2 * kMaxTiles) * 8 + 15) & ~15;
    L.st_xs = off;     off += kArenaEntries * 8;
    L.vec = off;
    const long long vec_bytes = (long long)(gmres ? memory + 1 : 5) * ((A->str_max_rows + 1) & ~1) * 8;
    if (off + vec_bytes <= limit) {
        L.vec_rows = (A->str_max_rows + 1) & ~1;
        off += (int)vec_bytes;
    }
    L.total = off;
    L.streaming = 1;
    if (L.total > limit) memset(&L, 0, sizeof(L));
    return L;
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
    // CTAs of the persistent kernel: all SMs by default (NUPGCM_GRID overrides, for experiments)
    int grid = ctx->coop_grid;
    if (const char *eg = getenv("NUPGCM_GRID")) {
        const int v = atoi(eg);
        if (v >= 1 && v <= ctx->coop_grid) grid = v;
    }
    NUPGCM_REQUIRE(ctx, grid <= 32 * kPollWarps, "solve: grid larger than the reduction supports");
    if (gmres) {
        NUPGCM_REQUIRE(ctx, memory >= 1 && memory <= kMaxMemory, "gmres: memory must be in 1..20");
        NUPGCM_REQUIRE(ctx, orth == NUPGCM_ORTH_MGS || orth == NUPGCM_ORTH_CGS2 || orth == NUPGCM_ORTH_CGS2_FUSED, "gmres: unknown orth");
    }
    NUPGCM_REQUIRE(ctx, hist_cap >= 0 && (hist_cap == 0 || resid_hist), "solve: hist_cap without buffer");
    const int64_t n = A->n_rows;
    if (n == 0) {
        if (stats) { memset(stats, 0, sizeof(*stats)); stats->solved = 1; }
        return NUPGCM_OK;
    }
    nupgcm_comm *comm = A->comm;
    const int nranks = comm ? comm->nranks : 1;
    if (comm) {
        NUPGCM_REQUIRE(ctx, comm->connected, "solve: communicator is not connected");
        NUPGCM_REQUIRE(ctx, !comm->broken, "solve: communicator is unusable after an aborted sharded solve");
        NUPGCM_REQUIRE(ctx, n <= comm->max_n, "solve: system larger than the communicator's max_n");
    }
    if (gmres && orth != NUPGCM_ORTH_MGS)
        NUPGCM_REQUIRE(ctx, grid >= memory + 1, "gmres: CGS2 needs at least `memory` CTAs (one reducer per projection)");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t rcp = nupgcm_csr_prepare(const_cast<nupgcm_csr *>(A), grid);
    if (rcp) return rcp;
    const size_t wbytes = (gmres ? (size_t)(memory + 4) : 7) * (size_t)n * sizeof(double);
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
    memset(&args, 0, sizeof(args));
    args.rowptr = A->d_prow;
    args.colidx = A->d_pcol;
    args.vals = A->d_pvals;
    args.perm = A->d_perm;
    args.part = A->d_part;
    args.loc = A->d_loc;
    args.foot_ptr = A->d_foot_ptr;
    args.foot = A->d_foot;
    args.lay = plan_resident(A, gmres, memory);
    const bool resident = args.lay.total > 0;
    if (!resident) args.lay = plan_streaming(A, gmres, memory);
    args.svals = A->d_svals;
    args.scols = A->d_scols;
    args.tiles = A->d_tiles;
    args.tile_ptr = A->d_tile_ptr;
    args.wdesc = A->d_wdesc;
    args.slices = A->d_slices;
    args.twcnt = A->d_twcnt;
    args.srow = A->d_srow;
    args.slen = A->d_slen;
    args.sfoot = A->d_sfoot;
    args.n = (int)n;
    args.dinv = dinv ? dinv->d : nullptr;
    args.pscale = pscale;
    args.b = y->d;
    args.x = x->d;
    args.work = ctx->d_ws;
    args.rank = comm ? comm->rank : 0;
    args.nranks = nranks;
    args.xgen_base = comm ? comm->xgen : 0;
    args.xmode = 1;
    if (const char *ex = getenv("NUPGCM_XMODE")) args.xmode = atoi(ex) != 0;
    if (const char *ex = getenv("NUPGCM_XFENCE")) args.xfence = atoi(ex) != 0;
    args.halo_ptr = A->d_halo_ptr;
    args.halo_idx = A->d_halo_idx;
    args.ll_off = comm ? (long long)(kArenaVecOffset + 3 * (size_t)comm->n_pad * sizeof(double)) : 0;
    for (int p = 0; p < nranks && comm; ++p) {
        args.arena[p] = comm->peer[p];
        args.push_lo[p] = A->push_lo[p];
        args.push_hi[p] = A->push_hi[p];
    }
    args.atol = atol;
    args.rtol = rtol;
    args.itmax = itmax;
    args.mem = memory;
    args.orth = orth;
    args.poll_depth = poll_config(grid);
    args.barrier = ctx->d_barrier;
    args.partials = ctx->d_partials;
    args.hist = hist_cap > 0 ? ctx->d_hist : nullptr;
    args.hist_cap = hist_cap;
    args.result = ctx->d_scalars;

    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4 * sizeof(unsigned long long), ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, reduce_scratch_bytes(ctx->coop_grid), ctx->stream));
    {
        // experiment switch: the constant is per device, so it is written on every solve while the switch is (or
        // has just been) in use, and never touched otherwise
        static bool sleep_in_use = false;
        int want = 0;
        if (const char *es = getenv("NUPGCM_POLL_SLEEP")) want = std::max(0, atoi(es));
        if (want != 0 || sleep_in_use) {
            NUPGCM_CUDA(ctx, cudaMemcpyToSymbolAsync(c_poll_sleep_ns, &want, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream));
            sleep_in_use = true;
        }
    }
    const char *trace_path = getenv("NUPGCM_TRACE_FILE");       // debug only
    unsigned long long *d_trace = nullptr;
    const size_t trace_words = (size_t)kTraceWindow * kTraceStamps * grid;
    if (trace_path && gmres) {
        NUPGCM_CUDA(ctx, cudaMalloc(&d_trace, trace_words * sizeof(unsigned long long)));
        NUPGCM_CUDA(ctx, cudaMemsetAsync(d_trace, 0, trace_words * sizeof(unsigned long long), ctx->stream));
    }
    args.trace = d_trace;
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev0, ctx->stream));
    cudaError_t e;
    const size_t smem = (size_t)args.lay.total;
    switch (args.lay.streaming ? A->str_T : persistent_tpr(A, resident, grid * nranks)) {
        case 32: e = launch<32>(gmres, args, ctx, smem, grid, resident); break;
        case 16: e = launch<16>(gmres, args, ctx, smem, grid, resident); break;
        case 8: e = launch<8>(gmres, args, ctx, smem, grid, resident); break;
        case 4: e = launch<4>(gmres, args, ctx, smem, grid, resident); break;
        default: e = launch<2>(gmres, args, ctx, smem, grid, resident); break;
    }
    NUPGCM_CUDA(ctx, e);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev1, ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 14 * sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    {
        cudaError_t es = cudaStreamSynchronize(ctx->stream);
        if (es != cudaSuccess) {
            if (comm) comm->broken = 1;
            NUPGCM_CUDA(ctx, es);
        }
    }
    const double *res = ctx->h_scalars;
    if (comm) {
        // all ranks executed the same reductions: the communicator's sequence number advances in step
        if (res[7] != 0.0) comm->broken = 1;
        else comm->xgen += (unsigned)res[13];
    }
    if (d_trace) {
        unsigned long long *h = (unsigned long long *)malloc(trace_words * sizeof(unsigned long long));
        cudaMemcpy(h, d_trace, trace_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(trace_path, "wb")) {
            const int hdr[4] = {grid, (int)kTraceWindow, (int)kTraceStamps, 0};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h, sizeof(unsigned long long), trace_words, f);
            fclose(f);
        }
        free(h);
        cudaFree(d_trace);
    }
    if (res[7] != 0.0)
        return nupgcm_fail(ctx, comm ? NUPGCM_ERR_COMM : NUPGCM_ERR_CUDA, "%s",
                           "persistent solver kernel aborted: cross-CTA wait watchdog expired");
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
        const char *ep = getenv("NUPGCM_PROFILE");
        if (gmres && ep && atoi(ep) != 0) {
            const double tot = res[8] + res[9] + res[10] + res[11];
            for (int i = 0; i < 4; ++i) stats->phase_frac[i] = tot > 0 ? (float)(res[8 + i] / tot) : 0.f;
            stats->sm_mhz = (float)res[12];
        }
    }
    return NUPGCM_OK;
}


// ---- diagnostics: the streaming SpMV engine alone ------------------------------------------------
// y = A x computed `reps` times by the persistent kernels' streaming engine (one CTA per SM, same
// tables, same warp roles) without any grid-wide wait: a plain kernel that ncu can replay, the unit
// test of the engine, and the SpMV leg of the roofline report.  x / y are in caller order; the kernel
// gathers x through the internal ordering into `xi` first (a separate launch).
__global__ void k_diag_permute_in(double *xi, const double *x, const int32_t *perm, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xi[i] = x[perm[i]];
}

template <int T, int DBG>
__global__ void __launch_bounds__(kThreads, 1) k_diag_stream_spmv(const __grid_constant__ KrylovArgs a, const double *xi,
                                                                  double *y, int reps) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ CommMailbox mailbox;
    if (threadIdx.x == 0) mailbox.dead = 0;
    const int r0 = a.part[blockIdx.x], r1 = a.part[blockIdx.x + 1];
    SpmvEngine<T, false> eng(a, r0, r1, dyn_smem, &mailbox);
    __syncthreads();
    if (threadIdx.x >= kMainThreads) {
        comm_warp_loop<false>(&mailbox, a, 1, dyn_smem);
        return;
    }
    for (int rep = 0; rep < reps; ++rep) {
        eng.template run<DBG>(xi, [&](int row, double v) { y[a.perm[row]] = v; });
        const long long tsync = DBG == 3 ? clock64() : 0;
        main_sync();
        if (DBG == 3 && a.trace && (threadIdx.x & 31) == 0)
            a.trace[((size_t)blockIdx.x * kMainWarps + (threadIdx.x >> 5)) * 8 + 6] += (unsigned long long)(clock64() - tsync);
    }
    if (threadIdx.x == 0) mailbox.count = kReqExit;
    named_arrive(NUPGCM_BAR_REQ, kThreads);
    eng.finish();
    if (blockIdx.x == 0 && threadIdx.x == 0) a.result[7] = mailbox.dead ? 1.0 : 0.0;
}

extern "C" int32_t nupgcm_diag_stream_spmv(nupgcm_csr *A, const nupgcm_vec *x, nupgcm_vec *y, int32_t reps,
                                           int32_t mode, float *us_per_spmv, double *warp_cycles) {
    NUPGCM_REQUIRE(nullptr, A && x && y, "diag_stream_spmv: NULL argument");
    nupgcm_ctx *ctx = A->ctx;
    NUPGCM_REQUIRE(ctx, A->n_rows == A->n_cols && x->n == A->n_rows && y->n == A->n_rows && x->d != y->d,
                   "diag_stream_spmv: square matrix and distinct vectors of its size");
    NUPGCM_REQUIRE(ctx, reps >= 1 && mode >= 0 && mode <= 4 && !A->comm && (mode != 3 || warp_cycles), "diag_stream_spmv: bad argument");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int grid = ctx->coop_grid;
    int32_t rc = nupgcm_csr_prepare(A, grid);
    if (rc) return rc;
    KrylovArgs args;
    memset(&args, 0, sizeof(args));
    args.lay = plan_streaming(A, false, 0);
    NUPGCM_REQUIRE(ctx, args.lay.streaming == 1, "diag_stream_spmv: the matrix has no streaming tables (it is SM-resident; set NUPGCM_RESIDENT=0)");
    rc = ensure_workspace(ctx, (size_t)A->n_rows * sizeof(double));
    if (rc) return rc;
    args.rowptr = A->d_prow; args.colidx = A->d_pcol; args.vals = A->d_pvals; args.perm = A->d_perm; args.part = A->d_part;
    args.svals = A->d_svals; args.scols = A->d_scols; args.tiles = A->d_tiles; args.tile_ptr = A->d_tile_ptr;
    args.wdesc = A->d_wdesc; args.slices = A->d_slices; args.twcnt = A->d_twcnt; args.srow = A->d_srow; args.slen = A->d_slen; args.sfoot = A->d_sfoot;
    args.n = (int)A->n_rows;
    args.nranks = 1;
    args.barrier = ctx->d_barrier;
    args.result = ctx->d_scalars;
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4 * sizeof(unsigned long long), ctx->stream));
    k_diag_permute_in<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_ws, x->d, A->d_perm, (int)A->n_rows);
    const void *fn;
    fn = mode == 0 ? (const void *)k_diag_stream_spmv<8, 0> : mode == 1 ? (const void *)k_diag_stream_spmv<8, 1>
       : mode == 2 ? (const void *)k_diag_stream_spmv<8, 2> : mode == 3 ? (const void *)k_diag_stream_spmv<8, 3>
       : (const void *)k_diag_stream_spmv<8, 4>;
    unsigned long long *d_trace = nullptr;
    const size_t trace_words = (size_t)grid * kMainWarps * 8;
    if (mode == 3) {
        NUPGCM_CUDA(ctx, cudaMalloc(&d_trace, trace_words * sizeof(unsigned long long)));
        NUPGCM_CUDA(ctx, cudaMemsetAsync(d_trace, 0, trace_words * sizeof(unsigned long long), ctx->stream));
    }
    args.trace = d_trace;
    NUPGCM_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, args.lay.total));
    const double *xi = ctx->d_ws;
    double *yd = y->d;
    int r = reps;
    void *params[] = {&args, &xi, &yd, &r};
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev0, ctx->stream));
    NUPGCM_CUDA(ctx, cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), params, (size_t)args.lay.total, ctx->stream));
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev1, ctx->stream));
    ctx->launches += 2;
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (d_trace) {
        std::vector<unsigned long long> h(trace_words);
        cudaMemcpy(h.data(), d_trace, trace_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(d_trace);
        for (size_t i = 0; i < trace_words; ++i) warp_cycles[i] = (double)h[i] / reps;
    }
    if (ctx->h_scalars[7] != 0.0) return nupgcm_fail(ctx, NUPGCM_ERR_CUDA, "%s", "diag_stream_spmv: a shared-memory pipeline wait timed out");
    float ms = 0.f;
    NUPGCM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->sev0, ctx->sev1));
    if (us_per_spmv) *us_per_spmv = 1e3f * ms / reps;
    return NUPGCM_OK;
}


// ---- diagnostics: throughput of TMA bulk copies as a function of their size ------------------------
// Every warp of every CTA streams its own contiguous region of `src` through `slots` shared-memory
// buffers of `piece` bytes with cp.async.bulk (one mbarrier phase per piece), touching nothing: the
// rate at which the copy engine alone can pull HBM into shared memory.  The streaming SpMV sizes its
// pieces from this measurement (profiles/tma_piece_size_r02.txt).
__global__ void __launch_bounds__(512, 1) k_diag_tma_stream(const unsigned char *src, long long per_warp, int piece, int slots, int reps) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int nw = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(dyn_smem);
    unsigned char *buf = dyn_smem + 1024 + (size_t)wid * slots * piece;
    uint64_t *bar = bars + wid * slots;
    if (lane == 0) {
        for (int i = 0; i < slots; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane != 0) return;
    const unsigned char *mine = src + ((size_t)blockIdx.x * nw + wid) * (size_t)per_warp;
    const int np = (int)(per_warp / piece);
    unsigned g = 0;                                              // pieces waited so far
    for (int rep = 0; rep < reps; ++rep) {
        int issued = 0;
        for (; issued < np && issued < slots; ++issued) {
            const unsigned sl = (g + issued) % slots;
            mbar_expect_tx(bar + sl, piece);
            bulk_g2s(buf + (size_t)sl * piece, mine + (size_t)issued * piece, piece, bar + sl);
        }
        for (int p = 0; p < np; ++p, ++g) {
            const unsigned sl = g % slots;
            mbar_wait(bar + sl, (g / slots) & 1u);
            if (issued < np) {
                mbar_expect_tx(bar + sl, piece);
                bulk_g2s(buf + (size_t)sl * piece, mine + (size_t)issued * piece, piece, bar + sl);
                ++issued;
            }
        }
    }
}

extern "C" int32_t nupgcm_diag_tma_stream(nupgcm_ctx *ctx, int64_t total_bytes, int32_t piece, int32_t slots,
                                          int32_t warps, int32_t reps, float *gb_per_s) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, piece >= 16 && piece % 16 == 0 && slots >= 1 && slots <= 32 && warps >= 1 && warps <= 16 &&
                            reps >= 1 && gb_per_s && total_bytes > 0, "diag_tma_stream: bad argument");
    const size_t smem = 1024 + (size_t)warps * slots * piece;
    NUPGCM_REQUIRE(ctx, smem <= 227 * 1024 && (size_t)warps * slots * 8 <= 1024, "diag_tma_stream: does not fit in shared memory");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int grid = ctx->coop_grid;
    long long per_warp = total_bytes / ((long long)grid * warps) / piece * piece;
    NUPGCM_REQUIRE(ctx, per_warp >= piece, "diag_tma_stream: total_bytes too small");
    unsigned char *src = nullptr;
    NUPGCM_CUDA(ctx, cudaMalloc(&src, (size_t)per_warp * grid * warps));
    NUPGCM_CUDA(ctx, cudaMemsetAsync(src, 1, (size_t)per_warp * grid * warps, ctx->stream));
    NUPGCM_CUDA(ctx, cudaFuncSetAttribute((const void *)k_diag_tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_diag_tma_stream<<<grid, 32 * warps, smem, ctx->stream>>>(src, per_warp, piece, slots, 1);     // warm-up
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev0, ctx->stream));
    k_diag_tma_stream<<<grid, 32 * warps, smem, ctx->stream>>>(src, per_warp, piece, slots, reps);
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev1, ctx->stream));
    ctx->launches += 2;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(src);
    NUPGCM_CUDA(ctx, e);
    float ms = 0.f;
    NUPGCM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->sev0, ctx->sev1));
    *gb_per_s = (float)((double)per_warp * grid * warps * reps / (ms * 1e-3) / 1e9);
    return NUPGCM_OK;
}

// ---- diagnostics: latency of the grid-wide reduction primitive ---------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_diag_reduce(const __grid_constant__ KrylovArgs a, int mode, int reps, double *out) {
    __shared__ CommMailbox mailbox;
    if (threadIdx.x == 0) mailbox.dead = 0;
    __syncthreads();
    if (threadIdx.x >= kMainThreads) {
        comm_warp_loop(&mailbox, a, a.poll_depth);
        return;
    }
    GridReduce gr;
    gr.init(&mailbox);
    double v = 1.0 + blockIdx.x, acc = 0.0;
    for (int i = 0; i < reps && !gr.aborted(); ++i) {
        const double mine = threadIdx.x == 0 ? v : 0.0;
        const double s = mode == 2 ? gr.sum_threads<true>(mine) : gr.sum_threads<false>(mine);
        acc += s;
        v = s * 1e-6 + blockIdx.x;
    }
    gr.finish();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = acc; out[7] = gr.aborted() ? 1.0 : 0.0; }
}

// Ping-pong between CTA 0 and CTA `peer` through one global word: one-way visibility latency of
// different store/load flavours (variant = 10*store + load).
//   store: 0 st.relaxed.gpu  1 st.volatile  2 atomicExch  3 st.release.gpu  4 plain st + __threadfence
//   load : 0 ld.relaxed.gpu  1 ld.volatile  2 ld.global.cg  3 atomicAdd(p,0)  4 ld.acquire.gpu
__device__ __forceinline__ void pp_store(unsigned long long *p, unsigned long long v, int kind) {
    switch (kind) {
        case 0: asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); break;
        case 1: asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); break;
        case 2: atomicExch(p, v); break;
        case 3: asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); break;
        default: *p = v; __threadfence(); break;
    }
}
__device__ __forceinline__ unsigned long long pp_load(unsigned long long *p, int kind) {
    unsigned long long v;
    switch (kind) {
        case 0: asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); break;
        case 1: asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); break;
        case 2: asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); break;
        case 3: v = atomicAdd(p, 0ULL); break;
        default: asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); break;
    }
    return v;
}
__global__ void k_diag_pingpong(unsigned long long *words, int peer, int variant, int reps, double *out) {
    if (threadIdx.x != 0) return;
    const int sk = variant / 10, lk = variant % 10;
    unsigned long long *mine, *theirs;
    if (blockIdx.x == 0) { mine = words; theirs = words + 16; }
    else if ((int)blockIdx.x == peer) { mine = words + 16; theirs = words; }
    else return;
    const unsigned long long t_start = global_timer_ns();
    bool ok = true;
    for (int i = 1; i <= reps && ok; ++i) {
        if (blockIdx.x == 0) pp_store(mine, (unsigned long long)i, sk);
        unsigned long long spins = 0;
        while (pp_load(theirs, lk) < (unsigned long long)i)
            if (++spins > 20000000ULL) { ok = false; break; }
        if (blockIdx.x != 0) pp_store(mine, (unsigned long long)i, sk);
    }
    if (blockIdx.x == 0) { out[0] = (double)(global_timer_ns() - t_start); out[7] = ok ? 0.0 : 1.0; }
}

extern "C" int32_t nupgcm_diag_pingpong(nupgcm_ctx *ctx, int32_t peer, int32_t variant, int32_t reps,
                                        float *us_round_trip) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, peer >= 1 && peer < ctx->coop_grid && reps > 0 && us_round_trip, "diag: bad argument");
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, 64 * sizeof(double), ctx->stream));
    unsigned long long *w = reinterpret_cast<unsigned long long *>(ctx->d_partials);
    double *out = ctx->d_scalars;
    int p = peer, v = variant, r = reps;
    void *params[] = {&w, &p, &v, &r, &out};
    NUPGCM_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_diag_pingpong, dim3(ctx->coop_grid), dim3(32), params, 0, ctx->stream));
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_scalars[7] != 0.0) return nupgcm_fail(ctx, NUPGCM_ERR_CUDA, "%s", "ping-pong timed out");
    *us_round_trip = (float)(ctx->h_scalars[0] * 1e-3 / reps);
    return NUPGCM_OK;
}

// Ping-pong between two RANKS through one word of each other's arena: the one-way latency of a
// flag that crosses NVLink (variant 0: relaxed.sys, 1: release.sys / acquire.sys, 2: fence.sys +
// relaxed).  Collective over the two ranks involved; the others return immediately.
__global__ void k_diag_xping(unsigned long long *mine, unsigned long long *theirs, int initiator, int variant,
                             unsigned long long seq0, int reps, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long t_start = global_timer_ns();
    bool ok = true;
    for (int i = 1; i <= reps && ok; ++i) {
        const unsigned long long tag = seq0 + (unsigned long long)i;
        for (int phase = 0; phase < 2; ++phase) {
            if ((phase == 0) == (initiator != 0)) {
                if (variant == 2) __threadfence_system();
                if (variant == 1) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(tag) : "memory");
                else asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(tag) : "memory");
            } else {
                unsigned long long v, spins = 0;
                for (;;) {
                    if (variant == 1) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
                    else asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
                    if (v >= tag) break;
                    if ((++spins & 1023ULL) == 0 && global_timer_ns() - t_start > kWaitTimeoutNs) { ok = false; break; }
                }
                if (!ok) break;
            }
        }
    }
    out[0] = (double)(global_timer_ns() - t_start);
    out[7] = ok ? 0.0 : 1.0;
}

extern "C" int32_t nupgcm_diag_xping(nupgcm_comm *comm, int32_t rank_a, int32_t rank_b, int32_t variant,
                                     int32_t reps, float *us_one_way) {
    NUPGCM_REQUIRE(nullptr, comm, "comm is NULL");
    nupgcm_ctx *ctx = comm->ctx;
    NUPGCM_REQUIRE(ctx, comm->connected && rank_a != rank_b && rank_a >= 0 && rank_b >= 0 &&
                            rank_a < comm->nranks && rank_b < comm->nranks && reps > 0 && us_one_way,
                   "diag_xping: bad argument");
    const unsigned long long seq0 = comm->ping_seq;
    comm->ping_seq += (unsigned long long)reps;
    *us_one_way = 0.f;
    if (comm->rank != rank_a && comm->rank != rank_b) return NUPGCM_OK;
    const int other = comm->rank == rank_a ? rank_b : rank_a;
    unsigned long long *mine = reinterpret_cast<unsigned long long *>(comm->arena + 64);
    unsigned long long *theirs = reinterpret_cast<unsigned long long *>(comm->peer[other] + 64);
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    k_diag_xping<<<1, 32, 0, ctx->stream>>>(mine, theirs, comm->rank == rank_a, variant, seq0, reps, ctx->d_scalars);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_scalars[7] != 0.0) return nupgcm_fail(ctx, NUPGCM_ERR_CUDA, "%s", "cross-rank ping-pong timed out");
    *us_one_way = (float)(ctx->h_scalars[0] * 1e-3 / reps / 2.0);
    return NUPGCM_OK;
}

// Latency of the sharded solvers' reduction: `count` values per reduction (1: sum_threads, >1: sumN),
// publish = 1 adds the system fence + release/acquire chain.  Collective over all ranks.
__global__ void __launch_bounds__(kThreads, 1) k_diag_xreduce(const __grid_constant__ KrylovArgs a, int count, int publish,
                                                               int reps, double *out) {
    __shared__ CommMailbox mailbox;
    if (threadIdx.x == 0) mailbox.dead = 0;
    __syncthreads();
    if (threadIdx.x >= kMainThreads) {
        comm_warp_loop<true>(&mailbox, a, a.poll_depth);
        return;
    }
    GridReduce gr;
    gr.init(&mailbox, nullptr, publish == 1 || (publish == 2 && (blockIdx.x % 8) == 0));   // 1: every CTA pushes, 2: one in eight
    double v = 1.0 + blockIdx.x, acc = 0.0;
    const unsigned long long t0 = global_timer_ns();
    for (int i = 0; i < reps && !gr.aborted(); ++i) {
        double s;
        if (count == 1) {
            const double mine = threadIdx.x == 0 ? v : 0.0;
            s = publish ? gr.sum_threads<true>(mine) : gr.sum_threads<false>(mine);
        } else {
            if (threadIdx.x < count) mailbox.in[threadIdx.x] = v + threadIdx.x;
            main_sync();
            gr.sumN(count);
            s = mailbox.out[count - 1];
            main_sync();
        }
        acc += s;
        v = s * 1e-9 + blockIdx.x;
    }
    const unsigned long long t1 = global_timer_ns();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = acc; out[1] = (double)(t1 - t0); out[7] = gr.aborted() ? 1.0 : 0.0; out[13] = (double)gr.gen; }
    gr.finish();
}

extern "C" int32_t nupgcm_diag_xreduce(nupgcm_comm *comm, int32_t count, int32_t publish, int32_t reps,
                                       float *us_per_reduction) {
    NUPGCM_REQUIRE(nullptr, comm, "comm is NULL");
    nupgcm_ctx *ctx = comm->ctx;
    NUPGCM_REQUIRE(ctx, comm->connected && !comm->broken && comm->nranks > 1, "diag_xreduce: needs a connected multi-rank communicator");
    NUPGCM_REQUIRE(ctx, count >= 1 && count <= kMaxMemory && reps > 0 && us_per_reduction, "diag_xreduce: bad argument");
    const int grid = ctx->coop_grid;
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4 * sizeof(unsigned long long), ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, reduce_scratch_bytes(ctx->coop_grid), ctx->stream));
    KrylovArgs args;
    memset(&args, 0, sizeof(args));
    args.barrier = ctx->d_barrier;
    args.partials = ctx->d_partials;
    args.poll_depth = poll_config(grid);
    args.rank = comm->rank;
    args.nranks = comm->nranks;
    args.xgen_base = comm->xgen;
    args.xmode = 1;
    if (const char *ex = getenv("NUPGCM_XMODE")) args.xmode = atoi(ex) != 0;
    if (const char *ex = getenv("NUPGCM_XFENCE")) args.xfence = atoi(ex) != 0;
    for (int p = 0; p < comm->nranks; ++p) args.arena[p] = comm->peer[p];
    double *out = ctx->d_scalars;
    int c = count, pb = publish, r = reps;
    void *params[] = {&args, &c, &pb, &r, &out};
    NUPGCM_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_diag_xreduce, dim3(grid), dim3(kThreads), params, 0, ctx->stream));
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (es != cudaSuccess) { comm->broken = 1; NUPGCM_CUDA(ctx, es); }
    if (ctx->h_scalars[7] != 0.0) {
        comm->broken = 1;
        return nupgcm_fail(ctx, NUPGCM_ERR_COMM, "%s", "diag kernel aborted: cross-rank wait watchdog expired");
    }
    comm->xgen += (unsigned)ctx->h_scalars[13];
    *us_per_reduction = (float)(ctx->h_scalars[1] * 1e-3 / reps);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_diag_reduce_latency(nupgcm_ctx *ctx, int32_t mode, int32_t reps, int32_t grid,
                                              int32_t threads, float *us_per_reduction) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, mode >= 0 && mode <= 2 && reps > 0 && us_per_reduction, "diag: bad argument");
    NUPGCM_REQUIRE(ctx, grid >= 1 && grid <= ctx->coop_grid, "diag: grid out of range");
    (void)threads;                                      // the kernels always run kThreads per CTA
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4 * sizeof(unsigned long long), ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, reduce_scratch_bytes(ctx->coop_grid), ctx->stream));
    KrylovArgs args;
    memset(&args, 0, sizeof(args));
    args.barrier = ctx->d_barrier;
    args.partials = ctx->d_partials;
    args.poll_depth = poll_config(grid);
    double *out = ctx->d_scalars;
    int m = mode, r = reps;
    void *params[] = {&args, &m, &r, &out};
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev0, ctx->stream));
    NUPGCM_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_diag_reduce, dim3(grid), dim3(kThreads), params, 0, ctx->stream));
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaEventRecord(ctx->sev1, ctx->stream));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_scalars[7] != 0.0)
        return nupgcm_fail(ctx, NUPGCM_ERR_CUDA, "%s", "diag kernel aborted: cross-CTA wait watchdog expired");
    float ms = 0.f;
    NUPGCM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->sev0, ctx->sev1));
    *us_per_reduction = 1e3f * ms / reps;
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
