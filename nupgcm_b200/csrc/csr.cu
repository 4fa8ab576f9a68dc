// CSR matrices on the device and the stand-alone FP64 SpMV.
//
// Replaces `CuSparseMatrixCSR(a)` (ext/nuPGCMCUDAExt.jl:27) and the `cusparseSpMV` behind
// `mul!(y, A, x)`; the same row kernel is instantiated inside the persistent Krylov kernels.
// Rows are processed by sub-warps of T = 2..32 lanes ("vector" CSR); T is chosen from the mean
// row length so that a lane sees ~4-8 entries (inversion matrix: ~50-77 per row -> T = 16;
// evolution matrix: ~23 per row -> T = 4).  Summation order inside a row is fixed (lane-strided
// partial sums, then an xor-shuffle tree), so results are run-to-run reproducible.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
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
// Cost of a row in matrix-entry units (NUPGCM_ROW_WEIGHT overrides): besides its entries a row costs
// the per-row part of the SpMV and, above all, its share of the Krylov vector work (~4 k smem
// fma per iteration for k basis vectors), which does not depend on the row's length.
static double row_weight() {
    if (const char *e = getenv("NUPGCM_ROW_WEIGHT")) {
        const double v = atof(e);
        if (v >= 0.0) return v;
    }
    return 4.0;
}

static void build_partition(const std::vector<int32_t> &rowptr, int64_t n_rows, int parts,
                            std::vector<int32_t> &part) {
    part.assign(parts + 1, 0);
    const double w = row_weight();
    const double total = (double)rowptr[n_rows] + w * (double)n_rows;
    int64_t r = 0;
    for (int p = 1; p < parts; ++p) {
        const double target = total * p / parts;
        while (r < n_rows && (double)rowptr[r] + w * (double)r < target) ++r;
        part[p] = (int32_t)r;
    }
    part[parts] = (int32_t)n_rows;
    for (int p = 1; p <= parts; ++p) part[p] = std::max(part[p], part[p - 1]);
}

// Sharded solves: bounding range [lo, hi) of the rows owned by rank `src` that the rows of rank
// `dst` reference (dst's halo inside src's block); lo = hi = 0 when there are none.  `part` is the
// partition over all ranks' CTAs (rank r owns parts [r*gpr, (r+1)*gpr)).
static void halo_range(const int32_t *rowptr, const int32_t *col, const std::vector<int32_t> &part, int gpr,
                       int dst, int src, int &lo_out, int &hi_out) {
    const int32_t s0 = part[(size_t)src * gpr], s1 = part[(size_t)(src + 1) * gpr];
    const int32_t q0 = part[(size_t)dst * gpr], q1 = part[(size_t)(dst + 1) * gpr];
    int32_t lo = INT32_MAX, hi = -1;
    for (int32_t k = rowptr[q0]; k < rowptr[q1]; ++k) {
        const int32_t c = col[k];
        if (c >= s0 && c < s1) { lo = std::min(lo, c); hi = std::max(hi, c); }
    }
    lo_out = hi >= lo ? lo : 0;
    hi_out = hi >= lo ? hi + 1 : 0;
}

// Reverse Cuthill-McKee ordering of the symmetrised pattern (A + Aᵀ).  The persistent solvers
// work on an internally reordered copy of the matrix: the reference orders the inversion system
// as [u DOFs (RCM) ; p DOFs (RCM)] (src/dofs.jl:38), which puts every pressure row far from the
// velocity columns it couples to — per-CTA column footprints of up to 4.7 k entries and row blocks
// of 82-358 rows at h = 0.08.  One RCM over the whole matrix gives 1.7 k / 169-345 and a bandwidth
// of 3.5 k instead of 30.7 k.  The permutation never leaves the library: solves gather their
// right-hand side / initial guess through it and scatter the solution back.
static void rcm_order(int64_t n, const int32_t *rowptr, const int32_t *col, std::vector<int32_t> &perm) {
    std::vector<int64_t> deg(n, 0);
    for (int64_t r = 0; r < n; ++r)
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
            if (col[k] != r) { deg[r]++; deg[col[k]]++; }
    std::vector<int64_t> ptr(n + 1, 0);
    for (int64_t r = 0; r < n; ++r) ptr[r + 1] = ptr[r] + deg[r];
    std::vector<int32_t> adj(ptr[n]);
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t r = 0; r < n; ++r)
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
            if (col[k] != r) { adj[fill[r]++] = col[k]; adj[fill[col[k]]++] = (int32_t)r; }
    std::vector<int32_t> d(n);                     // degree after removing duplicates
    for (int64_t r = 0; r < n; ++r) {
        std::sort(adj.begin() + ptr[r], adj.begin() + ptr[r + 1]);
        d[r] = (int32_t)(std::unique(adj.begin() + ptr[r], adj.begin() + ptr[r + 1]) - (adj.begin() + ptr[r]));
    }
    perm.clear();
    perm.reserve(n);
    std::vector<char> seen(n, 0);
    std::vector<int32_t> level, nbr;
    auto bfs = [&](int32_t start, std::vector<int32_t> &order) {      // order includes start
        size_t head = order.size();
        order.push_back(start);
        seen[start] = 1;
        while (head < order.size()) {
            const int32_t v = order[head++];
            nbr.clear();
            for (int64_t k = ptr[v]; k < ptr[v] + d[v]; ++k)
                if (!seen[adj[k]]) { seen[adj[k]] = 1; nbr.push_back(adj[k]); }
            std::sort(nbr.begin(), nbr.end(), [&](int32_t x, int32_t y) { return d[x] != d[y] ? d[x] < d[y] : x < y; });
            order.insert(order.end(), nbr.begin(), nbr.end());
        }
    };
    std::vector<int32_t> by_degree(n);
    for (int64_t i = 0; i < n; ++i) by_degree[i] = (int32_t)i;
    std::sort(by_degree.begin(), by_degree.end(), [&](int32_t x, int32_t y) { return d[x] != d[y] ? d[x] < d[y] : x < y; });
    for (int32_t cand : by_degree) {
        if (seen[cand]) continue;
        // pseudo-peripheral start: two BFS sweeps from the minimum-degree node of the component
        int32_t start = cand;
        for (int sweep = 0; sweep < 2; ++sweep) {
            level.clear();
            bfs(start, level);
            for (int32_t v : level) seen[v] = 0;
            start = level.back();
        }
        bfs(start, perm);
    }
    std::reverse(perm.begin(), perm.end());
}

// Structure of P A Pᵀ for the ordering `perm` (internal row i = caller row perm[i]); psrc[k] is the
// position of reordered entry k in the caller-order arrays.
static void permute_structure(int64_t n, const int32_t *rowptr, const int32_t *col, const std::vector<int32_t> &perm,
                              std::vector<int32_t> &prow, std::vector<int32_t> &pcol, std::vector<int32_t> &psrc) {
    const int64_t nnz = rowptr[n];
    std::vector<int32_t> inv(n);
    for (int64_t i = 0; i < n; ++i) inv[perm[i]] = (int32_t)i;
    prow.assign(n + 1, 0);
    pcol.resize(nnz);
    psrc.resize(nnz);
    std::vector<std::pair<int32_t, int32_t>> rowbuf;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t r = perm[i];
        rowbuf.clear();
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) rowbuf.emplace_back(inv[col[k]], k);
        std::sort(rowbuf.begin(), rowbuf.end());
        int32_t o = prow[i];
        for (auto &e : rowbuf) { pcol[o] = e.first; psrc[o] = e.second; ++o; }
        prow[i + 1] = o;
    }
}


// ---- streaming SpMV tables (common.cuh) ------------------------------------------------------
// Built for the CTAs [cta0, cta0 + ncta) of the partition (this rank's CTAs).  Returns false when a
// tile's footprint does not fit the arena / 16-bit offsets (then the solvers fall back to direct loads).
struct StreamTables {
    std::vector<NupgcmTileDesc> tiles;
    std::vector<int32_t> tile_ptr;
    std::vector<NupgcmWarpDesc> wdesc;
    std::vector<NupgcmSlice> slices;
    std::vector<int32_t> srow, slen, foot, ssrc;
    std::vector<uint16_t> scols;
    std::vector<uint8_t> twcnt;                    // [tiles][kMainWarps]: slices of each warp in each tile
    int max_foot = 0;
};

// Which bank pair does lane q of a half-warp read at one slice position?  avail[q] = bank pairs in which
// row q still has entries, size[q][b] = how many.  Maximum bipartite matching lanes <-> banks (Kuhn's
// augmenting paths, fullest bucket first), so that as many lanes as possible gather from distinct banks;
// lanes left over take their fullest bucket.
static void assign_banks(int nl, const uint16_t *avail, const int (*size)[16], int *choice) {
    int owner[16];
    for (int b = 0; b < 16; ++b) owner[b] = -1;
    struct Rec {
        const uint16_t *avail; const int (*size)[16]; int *owner; uint16_t seen;
        bool go(int q) {
            uint16_t cand = avail[q] & (uint16_t)~seen;
            while (cand) {
                int best = -1;
                for (uint16_t m = cand; m; m &= (uint16_t)(m - 1)) {
                    const int b = __builtin_ctz(m);
                    if (best < 0 || size[q][b] > size[q][best]) best = b;
                }
                cand &= (uint16_t)~(1u << best);
                seen |= (uint16_t)(1u << best);
                if (owner[best] < 0 || go(owner[best])) { owner[best] = q; return true; }
            }
            return false;
        }
    } rec{avail, size, owner, 0};
    for (int q = 0; q < nl; ++q) {
        choice[q] = -1;
        if (!avail[q]) continue;
        rec.seen = 0;
        rec.go(q);
    }
    for (int b = 0; b < 16; ++b)
        if (owner[b] >= 0) choice[owner[b]] = b;
    for (int q = 0; q < nl; ++q)
        if (choice[q] < 0 && avail[q]) {
            int best = -1;
            for (uint16_t m = avail[q]; m; m &= (uint16_t)(m - 1)) {
                const int b = __builtin_ctz(m);
                if (best < 0 || size[q][b] > size[q][best]) best = b;
            }
            choice[q] = best;
        }
}

// Tables of one CTA (rows [ra, rb)): tiles, footprints and the kMainWarps entry streams.
struct CtaStream {
    std::vector<NupgcmTileDesc> tiles;             // foot_off relative to `foot`
    std::vector<uint8_t> twcnt;                    // [tiles][kMainWarps]
    std::vector<int32_t> foot;
    std::vector<NupgcmSlice> wsl[kMainWarps];
    std::vector<int32_t> wrow[kMainWarps], wlen[kMainWarps], wsrc[kMainWarps];
    std::vector<uint16_t> wcols[kMainWarps];
    long long wcost[kMainWarps] = {0};
    int max_foot = 0;
    bool ok = true;
};

static void build_cta_stream(const int32_t *rowptr, const int32_t *col, int32_t ra, int32_t rb, int arena,
                             std::vector<int32_t> &mark, std::vector<int32_t> &loc_of, int32_t &stamp, CtaStream &cs) {
    const int W = kMainWarps;
    // rows per tile: kTileRows, doubled until the CTA needs at most kMaxTiles tiles
    int tile_rows = kTileRows;
    while ((int64_t)(rb - ra) > (int64_t)tile_rows * kMaxTiles) tile_rows *= 2;
    std::vector<int32_t> tile_cols, order;
    int32_t arena_pos = 0;
    int tix = 0;
    for (int32_t start = ra; start < rb; start += tile_rows, ++tix) {
        const int32_t r_end = std::min(rb, start + tile_rows);
        ++stamp;
        tile_cols.clear();
        for (int32_t k = rowptr[start]; k < rowptr[r_end]; ++k)
            if (mark[col[k]] != stamp) { mark[col[k]] = stamp; tile_cols.push_back(col[k]); }
        if ((int)tile_cols.size() > arena || tile_cols.size() > 8192) { cs.ok = false; return; }
        std::sort(tile_cols.begin(), tile_cols.end());
        for (size_t i = 0; i < tile_cols.size(); ++i) loc_of[tile_cols[i]] = (int32_t)i;
        cs.max_foot = std::max(cs.max_foot, (int)tile_cols.size());
        // arena slot: first fit going round; `dep` = the youngest earlier tile whose slot it overlaps
        const int32_t flen = ((int32_t)tile_cols.size() + 15) & ~15;
        if (arena_pos + flen > arena) arena_pos = 0;
        int32_t dep = -1;
        for (int u = (int)cs.tiles.size() - 1; u >= 0; --u) {
            const int32_t a0 = cs.tiles[u].xs_off, a1 = a0 + ((cs.tiles[u].foot_len + 15) & ~15);
            if (a0 < arena_pos + flen && arena_pos < a1) { dep = u; break; }
        }
        while (cs.foot.size() % 4) cs.foot.push_back(0);
        cs.tiles.push_back(NupgcmTileDesc{start, r_end - start, (int32_t)cs.foot.size(), (int32_t)tile_cols.size(),
                                          arena_pos, dep, 0, 0});
        arena_pos += flen;
        cs.foot.insert(cs.foot.end(), tile_cols.begin(), tile_cols.end());
        // rows by decreasing length (ties by row id); rows longer than kLongRow become whole-warp items,
        // the rest is cut into slices of 32.  Items go to the warps heaviest-first to the warp with the
        // least work so far (work = slice positions), so that the warps of the CTA stay level
        order.resize(r_end - start);
        for (int32_t i = 0; i < r_end - start; ++i) order[i] = start + i;
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
            return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y];
        });
        size_t nlong = 0;
        while (nlong < order.size() && rowptr[order[nlong] + 1] - rowptr[order[nlong]] > kLongRow) ++nlong;
        struct Item { size_t first; int nrows; int cost; };
        std::vector<Item> items;
        for (size_t i = 0; i < nlong; ++i)
            items.push_back(Item{i, 0, (rowptr[order[i] + 1] - rowptr[order[i]] + 31) / 32 + 4});
        for (size_t i = nlong; i < order.size(); i += 32)
            items.push_back(Item{i, (int)std::min<size_t>(32, order.size() - i), rowptr[order[i] + 1] - rowptr[order[i]] + 3});
        std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) { return x.cost > y.cost; });
        cs.twcnt.insert(cs.twcnt.end(), W, 0);
        for (const Item &it : items) {
            int w = 0;
            for (int v = 1; v < W; ++v)
                if (cs.wcost[v] < cs.wcost[w]) w = v;
            cs.wcost[w] += it.cost;
            cs.twcnt[(size_t)tix * W + w]++;
            if (cs.twcnt[(size_t)tix * W + w] == 255) { cs.ok = false; return; }
            const int lmax = rowptr[order[it.first] + 1] - rowptr[order[it.first]];
            cs.wsl[w].push_back(NupgcmSlice{(int32_t)cs.wcols[w].size(), (int32_t)cs.wrow[w].size(), it.nrows, lmax});
            if (it.nrows == 0) {
                // a long row: entries contiguous, interleaved over the bank pairs so that 16 consecutive
                // entries (one half-warp) mostly sit in 16 different banks
                const int32_t row = order[it.first];
                cs.wrow[w].push_back(row);
                cs.wlen[w].push_back(lmax);
                std::vector<int32_t> bucket[16];
                for (int32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) bucket[loc_of[col[k]] & 15].push_back(k);
                for (size_t r = 0; ; ++r) {
                    bool any = false;
                    for (int bk = 0; bk < 16; ++bk)
                        if (r < bucket[bk].size()) {
                            const int32_t k = bucket[bk][r];
                            cs.wcols[w].push_back((uint16_t)(8 * loc_of[col[k]]));
                            cs.wsrc[w].push_back(k);
                            any = true;
                        }
                    if (!any) break;
                }
                continue;
            }
            const int nr = it.nrows;
            const size_t i = it.first;
            // per row: its entries bucketed by the bank pair (position in the staged footprint mod 16)
            std::vector<int32_t> bucket[32][16];
            int left[32], size[32][16];
            uint16_t avail[32];
            for (int q = 0; q < nr; ++q) {
                const int32_t row = order[i + q];
                cs.wrow[w].push_back(row);
                cs.wlen[w].push_back(rowptr[row + 1] - rowptr[row]);
                left[q] = rowptr[row + 1] - rowptr[row];
                for (int32_t k = rowptr[row + 1] - 1; k >= rowptr[row]; --k) bucket[q][loc_of[col[k]] & 15].push_back(k);
                avail[q] = 0;
                for (int bk = 0; bk < 16; ++bk) {
                    size[q][bk] = (int)bucket[q][bk].size();
                    if (size[q][bk]) avail[q] |= (uint16_t)(1u << bk);
                }
            }
            // jagged-diagonal order, bank-aware: see assign_banks.  A full slice (32 rows) stores its first
            // nblk = lmin / 8 blocks of 8 positions "blocked": lane q's 8 offsets contiguous (one 16-byte load)
            // and its values in pairs (four 16-byte loads) — 13 shared-memory requests per block instead of 24;
            // such a slice starts at a multiple of 8 entries so that those loads are aligned.
            const int lmin = rowptr[order[i + nr - 1] + 1] - rowptr[order[i + nr - 1]];
            const int nblk = nr == 32 ? lmin / 8 : 0;
            if (nblk > 0) {
                while (cs.wcols[w].size() % 8) { cs.wcols[w].push_back(0); cs.wsrc[w].push_back(-1); }
                cs.wsl[w].back().eoff = (int32_t)cs.wcols[w].size();
                cs.wsl[w].back().nrows = nr | (nblk << 8);
            }
            uint16_t bc[8][32];
            int32_t bs[8][32];
            for (int j = 0; j < lmax; ++j) {
                int cnt = 0;
                while (cnt < nr && left[cnt] > 0) ++cnt;            // rows are sorted: lanes 0 .. cnt-1 hold position j
                int choice[32];
                assign_banks(std::min(cnt, 16), avail, size, choice);
                if (cnt > 16) assign_banks(cnt - 16, avail + 16, size + 16, choice + 16);
                for (int q = 0; q < cnt; ++q) {
                    const int bk = choice[q];
                    const int32_t k = bucket[q][bk].back();
                    bucket[q][bk].pop_back();
                    if (--size[q][bk] == 0) avail[q] &= (uint16_t)~(1u << bk);
                    --left[q];
                    if (j < 8 * nblk) {
                        bc[j & 7][q] = (uint16_t)(8 * loc_of[col[k]]);
                        bs[j & 7][q] = k;
                    } else {
                        cs.wcols[w].push_back((uint16_t)(8 * loc_of[col[k]]));
                        cs.wsrc[w].push_back(k);
                    }
                }
                if (j < 8 * nblk && (j & 7) == 7) {                 // a block is complete: emit it in the blocked layout
                    for (int e = 0; e < 256; ++e) {
                        cs.wcols[w].push_back(bc[e & 7][e >> 3]);                          // offsets: [lane][position]
                        cs.wsrc[w].push_back(bs[2 * (e >> 6) + (e & 1)][(e & 63) >> 1]);   // values: [pair][lane][2]
                    }
                }
            }
        }
    }
}

static bool build_stream_tables(const int32_t *rowptr, const int32_t *col, int64_t n_cols,
                                const std::vector<int32_t> &part, int cta0, int ncta, int arena,
                                StreamTables &st) {
    const int W = kMainWarps;
    st.tile_ptr.assign(ncta + 1, 0);
    st.wdesc.assign((size_t)ncta * W, NupgcmWarpDesc{0, 0, 0, 0});
    // the CTAs are independent: build them on all host cores, then concatenate in order
    std::vector<CtaStream> ctas(ncta);
    unsigned nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char *e = getenv("NUPGCM_HOST_THREADS")) nthreads = (unsigned)std::max(1, atoi(e));
    nthreads = std::min<unsigned>(nthreads, (unsigned)std::max(1, ncta));
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<int32_t> mark(n_cols, -1), loc_of(n_cols, 0);
        int32_t stamp = 0;
        for (int b = next.fetch_add(1); b < ncta; b = next.fetch_add(1))
            build_cta_stream(rowptr, col, part[cta0 + b], part[cta0 + b + 1], arena, mark, loc_of, stamp, ctas[b]);
    };
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < nthreads; ++i) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    for (int b = 0; b < ncta; ++b) {
        CtaStream &cs = ctas[b];
        if (!cs.ok) return false;
        st.max_foot = std::max(st.max_foot, cs.max_foot);
        st.tile_ptr[b] = (int32_t)st.tiles.size();
        while (st.foot.size() % 4) st.foot.push_back(0);
        const int32_t fbase = (int32_t)st.foot.size();
        for (NupgcmTileDesc td : cs.tiles) { td.foot_off += fbase; st.tiles.push_back(td); }
        st.twcnt.insert(st.twcnt.end(), cs.twcnt.begin(), cs.twcnt.end());
        st.foot.insert(st.foot.end(), cs.foot.begin(), cs.foot.end());
        for (int w = 0; w < W; ++w) {
            while (st.scols.size() % 8) { st.scols.push_back(0); st.ssrc.push_back(-1); }   // 16-byte aligned streams
            NupgcmWarpDesc &d = st.wdesc[(size_t)b * W + w];
            d.estart = (int32_t)st.scols.size();
            d.elen = (int32_t)cs.wcols[w].size();
            d.stab = (int32_t)st.slices.size();
            d.rtab = (int32_t)st.srow.size();
            st.scols.insert(st.scols.end(), cs.wcols[w].begin(), cs.wcols[w].end());
            st.ssrc.insert(st.ssrc.end(), cs.wsrc[w].begin(), cs.wsrc[w].end());
            st.slices.insert(st.slices.end(), cs.wsl[w].begin(), cs.wsl[w].end());
            st.slices.push_back(NupgcmSlice{0, 0, 0, 0});       // the kernels read two slice headers ahead
            st.slices.push_back(NupgcmSlice{0, 0, 0, 0});
            st.srow.insert(st.srow.end(), cs.wrow[w].begin(), cs.wrow[w].end());
            st.slen.insert(st.slen.end(), cs.wlen[w].begin(), cs.wlen[w].end());
        }
        cs = CtaStream();                                        // free as we go
    }
    st.tile_ptr[ncta] = (int32_t)st.tiles.size();
    // whole pieces are always copied: pad the tail
    for (int i = 0; i < kPieceEntries + 8; ++i) { st.scols.push_back(0); st.ssrc.push_back(-1); }
    for (int i = 0; i < 64; ++i) { st.srow.push_back(0); st.slen.push_back(0); }
    return true;
}

__global__ void k_scatter_stream(double *svals, const double *pvals, const int32_t *ssrc, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = ssrc[i];
        svals[i] = s >= 0 ? pvals[s] : 0.0;
    }
}

static void free_stream_tables(nupgcm_csr *A) {
    cudaFree(A->d_svals); A->d_svals = nullptr;
    cudaFree(A->d_scols); A->d_scols = nullptr;
    cudaFree(A->d_ssrc); A->d_ssrc = nullptr;
    cudaFree(A->d_tiles); A->d_tiles = nullptr;
    cudaFree(A->d_tile_ptr); A->d_tile_ptr = nullptr;
    cudaFree(A->d_wdesc); A->d_wdesc = nullptr;
    cudaFree(A->d_slices); A->d_slices = nullptr;
    cudaFree(A->d_twcnt); A->d_twcnt = nullptr;
    cudaFree(A->d_srow); A->d_srow = nullptr;
    cudaFree(A->d_slen); A->d_slen = nullptr;
    cudaFree(A->d_sfoot); A->d_sfoot = nullptr;
    A->stream_entries = 0;
    A->str_T = A->str_fmax = A->str_max_rows = 0;
    A->svals_version = -1;
}

template <class V>
static cudaError_t upload_vec(void **dst, const std::vector<V> &v, size_t extra = 8) {
    cudaError_t e = cudaMalloc(dst, (v.size() + extra) * sizeof(V));
    if (e != cudaSuccess) return e;
    e = cudaMemset(*dst, 0, (v.size() + extra) * sizeof(V));
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice);
}

// Internal (reordered) copy of the structure: built once per matrix.
static int32_t build_internal_order(nupgcm_csr *A) {
    nupgcm_ctx *ctx = A->ctx;
    if (A->d_perm) return NUPGCM_OK;
    const int64_t n = A->n_rows, nnz = A->nnz;
    std::vector<int32_t> perm;
    const char *env = getenv("NUPGCM_REORDER");
    if (env && atoi(env) == 0) {
        perm.resize(n);
        for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
    } else {
        rcm_order(n, A->h_rowptr, A->h_col, perm);
    }
    std::vector<int32_t> prow, pcol, psrc;
    permute_structure(n, A->h_rowptr, A->h_col, perm, prow, pcol, psrc);
    A->h_prow = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
    A->h_pcol = (int32_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
    if (!A->h_prow || !A->h_pcol) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    memcpy(A->h_prow, prow.data(), (size_t)(n + 1) * sizeof(int32_t));
    if (nnz) memcpy(A->h_pcol, pcol.data(), (size_t)nnz * sizeof(int32_t));
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_perm, (size_t)(n > 0 ? n : 1) * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemcpy(A->d_perm, perm.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_prow, (size_t)(n + 1 + 8) * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemset(A->d_prow, 0, (size_t)(n + 1 + 8) * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemcpy(A->d_prow, prow.data(), (size_t)(n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    // a little padding lets the persistent kernels bulk-copy 16-byte aligned windows
    const size_t nzpad = nz + 64;
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_pcol, nzpad * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemset(A->d_pcol, 0, nzpad * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_psrc, nz * sizeof(int32_t)));
    if (nnz) {
        NUPGCM_CUDA(ctx, cudaMemcpy(A->d_pcol, pcol.data(), (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
        NUPGCM_CUDA(ctx, cudaMemcpy(A->d_psrc, psrc.data(), (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_pvals, nzpad * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaMemset(A->d_pvals, 0, nzpad * sizeof(double)));
    A->pvals_version = -1;
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());   // set-up copies ran on the default stream
    return NUPGCM_OK;
}

// Row partition and SM-resident tables of the persistent solvers for a grid of `grid` CTAs, on
// the internally reordered structure.  Cached in the handle; rebuilt only when a solve asks for
// a different grid.  Also refreshes the reordered values when the caller-order values changed.
int32_t nupgcm_csr_prepare(nupgcm_csr *A, int grid_per_rank) {
    nupgcm_ctx *ctx = A->ctx;
    // sharded solves: the partition covers the CTAs of all ranks (rank r's CTA b is part r*grid+b);
    // every rank builds the same tables from the same host structure
    const int nranks = A->comm ? A->comm->nranks : 1;
    const int grid = grid_per_rank * nranks;
    int32_t rc = build_internal_order(A);
    if (rc) return rc;
    if (A->pvals_version != A->vals_version && A->nnz > 0) {
        int g = (int)std::min<int64_t>((A->nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
        k_scatter_vals<<<g, 256, 0, ctx->stream>>>(A->d_pvals, A->d_vals, A->d_psrc, A->nnz);
        ctx->launches++;
        NUPGCM_CUDA(ctx, cudaGetLastError());
        A->pvals_version = A->vals_version;
    }
    auto refresh_stream_values = [&]() -> int32_t {
        if (A->d_svals && A->svals_version != A->pvals_version) {
            int g = (int)std::min<int64_t>((A->stream_entries + 255) / 256, (int64_t)ctx->sm_count * 8);
            k_scatter_stream<<<g, 256, 0, ctx->stream>>>(A->d_svals, A->d_pvals, A->d_ssrc, A->stream_entries);
            ctx->launches++;
            NUPGCM_CUDA(ctx, cudaGetLastError());
            A->svals_version = A->pvals_version;
        }
        return NUPGCM_OK;
    };
    if (A->prepared_grid == grid && A->prepared_ranks == nranks) return refresh_stream_values();
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_stream_tables(A);
    cudaFree(A->d_part); A->d_part = nullptr;
    cudaFree(A->d_loc); A->d_loc = nullptr;
    cudaFree(A->d_foot_ptr); A->d_foot_ptr = nullptr;
    cudaFree(A->d_foot); A->d_foot = nullptr;
    A->res_max_nnz = A->res_max_foot = A->res_max_rows = 0;
    const int64_t n_rows = A->n_rows, kept = A->nnz;
    std::vector<int32_t> h_rowptr(A->h_prow, A->h_prow + n_rows + 1), part;
    build_partition(h_rowptr, n_rows, grid, part);
    NUPGCM_CUDA(ctx, cudaMalloc(&A->d_part, part.size() * sizeof(int32_t)));
    NUPGCM_CUDA(ctx, cudaMemcpy(A->d_part, part.data(), part.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    // SM-resident tables: only worth building when a CTA's slice can possibly fit on chip
    const int64_t slice_bytes = kept * 10 / (grid > 0 ? grid : 1);
    if (A->n_rows == A->n_cols && kept > 0 && slice_bytes < 230 * 1024) {
        const int32_t *h_col = A->h_pcol;
        std::vector<uint16_t> loc(kept);
        std::vector<int32_t> foot_ptr(grid + 1, 0), foot_len(grid, 0), foot, tmp;
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
            // pad each footprint to a multiple of 4 entries: 16-byte aligned bulk copies
            while (foot.size() % 4) foot.push_back(0);
            foot_ptr[p] = (int32_t)foot.size();
            foot.insert(foot.end(), tmp.begin(), tmp.end());
            foot_ptr[p + 1] = (int32_t)foot.size();
            foot_len[p] = (int32_t)tmp.size();
            max_nnz = std::max(max_nnz, (int)(k1 - k0));
            max_foot = std::max(max_foot, (int)tmp.size());
            max_rows = std::max(max_rows, (int)(part[p + 1] - part[p]));
        }
        if (ok) {
            std::vector<int32_t> fp2(2 * (size_t)grid);
            for (int p = 0; p < grid; ++p) { fp2[2 * p] = foot_ptr[p]; fp2[2 * p + 1] = foot_len[p]; }
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_loc, ((size_t)kept + 16) * sizeof(uint16_t)));
            NUPGCM_CUDA(ctx, cudaMemset(A->d_loc, 0, ((size_t)kept + 16) * sizeof(uint16_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_loc, loc.data(), (size_t)kept * sizeof(uint16_t), cudaMemcpyHostToDevice));
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_foot_ptr, fp2.size() * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_foot_ptr, fp2.data(), fp2.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            NUPGCM_CUDA(ctx, cudaMalloc(&A->d_foot, (foot.size() + 8) * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemset(A->d_foot, 0, (foot.size() + 8) * sizeof(int32_t)));
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_foot, foot.data(), foot.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            A->res_max_nnz = max_nnz;
            A->res_max_foot = max_foot;
            A->res_max_rows = max_rows;
        }
    }
    // streaming tables: whenever the SM-resident form cannot hold the slice (or is switched off)
    {
        const char *env = getenv("NUPGCM_RESIDENT");
        const bool resident_off = env && atoi(env) == 0;
        const long long res_bytes = 10LL * A->res_max_nnz + 12LL * A->res_max_foot + 4LL * A->res_max_rows + 256;
        const bool resident_fits = A->res_max_nnz > 0 && res_bytes <= 215 * 1024;
        if (A->n_rows == A->n_cols && kept > 0 && (resident_off || !resident_fits)) {
            const int me = nranks > 1 ? A->comm->rank : 0;
            StreamTables st;
            if (build_stream_tables(A->h_prow, A->h_pcol, A->n_cols, part, me * grid_per_rank, grid_per_rank, kArenaEntries, st)) {
                const int T = 8, fmax = st.max_foot;
                A->stream_entries = (int64_t)st.scols.size();
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_scols, st.scols));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_ssrc, st.ssrc));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_tiles, st.tiles));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_tile_ptr, st.tile_ptr));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_wdesc, st.wdesc));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_slices, st.slices));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_twcnt, st.twcnt));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_srow, st.srow));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_slen, st.slen));
                NUPGCM_CUDA(ctx, upload_vec((void **)&A->d_sfoot, st.foot));
                NUPGCM_CUDA(ctx, cudaMalloc(&A->d_svals, (size_t)(A->stream_entries + 8) * sizeof(double)));
                A->str_T = T;
                A->str_fmax = fmax;
                int max_rows = 0;
                for (int b = 0; b < grid_per_rank; ++b)
                    max_rows = std::max(max_rows, (int)(part[me * grid_per_rank + b + 1] - part[me * grid_per_rank + b]));
                A->str_max_rows = max_rows;
                A->svals_version = -1;
            }
        }
    }
    // halo push ranges: for every peer, the bounding range of this rank's rows it gathers
    for (int p = 0; p < kMaxRanks; ++p) A->push_lo[p] = A->push_hi[p] = 0;
    if (nranks > 1) {
        const int me = A->comm->rank;
        for (int p = 0; p < nranks; ++p)
            if (p != me) halo_range(h_rowptr.data(), A->h_pcol, part, grid_per_rank, p, me, A->push_lo[p], A->push_hi[p]);
    }
    // per CTA of this rank: the columns of its rows that other ranks own (unpacked from the
    // flagged halo words before each SpMV)
    cudaFree(A->d_halo_ptr); A->d_halo_ptr = nullptr;
    cudaFree(A->d_halo_idx); A->d_halo_idx = nullptr;
    A->halo_total = 0;
    if (nranks > 1) {
        const int me = A->comm->rank;
        const int32_t my0 = part[(size_t)me * grid_per_rank], my1 = part[(size_t)(me + 1) * grid_per_rank];
        std::vector<int32_t> hptr(grid_per_rank + 1, 0), hidx, tmp, all;
        for (int b = 0; b < grid_per_rank; ++b) {
            const int lgb = me * grid_per_rank + b;
            tmp.clear();
            for (int32_t k = h_rowptr[part[lgb]]; k < h_rowptr[part[lgb + 1]]; ++k) {
                const int32_t c = A->h_pcol[k];
                if (c < my0 || c >= my1) tmp.push_back(c);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            hidx.insert(hidx.end(), tmp.begin(), tmp.end());
            all.insert(all.end(), tmp.begin(), tmp.end());
            hptr[b + 1] = (int32_t)hidx.size();
        }
        std::sort(all.begin(), all.end());
        A->halo_total = (int64_t)(std::unique(all.begin(), all.end()) - all.begin());
        NUPGCM_CUDA(ctx, cudaMalloc(&A->d_halo_ptr, hptr.size() * sizeof(int32_t)));
        NUPGCM_CUDA(ctx, cudaMemcpy(A->d_halo_ptr, hptr.data(), hptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        NUPGCM_CUDA(ctx, cudaMalloc(&A->d_halo_idx, (hidx.size() + 1) * sizeof(int32_t)));
        if (!hidx.empty())
            NUPGCM_CUDA(ctx, cudaMemcpy(A->d_halo_idx, hidx.data(), hidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    A->prepared_grid = grid;
    A->prepared_ranks = nranks;
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());   // set-up copies ran on the default stream
    return refresh_stream_values();
}

extern "C" int32_t nupgcm_csr_shard(nupgcm_csr *A, nupgcm_comm *comm) {
    NUPGCM_REQUIRE(nullptr, A, "csr is NULL");
    nupgcm_ctx *ctx = A->ctx;
    if (comm) {
        NUPGCM_REQUIRE(ctx, comm->ctx == ctx, "csr_shard: matrix and communicator belong to different contexts");
        NUPGCM_REQUIRE(ctx, comm->connected, "csr_shard: communicator is not connected");
        NUPGCM_REQUIRE(ctx, A->n_rows == A->n_cols && A->n_rows <= comm->max_n,
                       "csr_shard: matrix must be square with n <= the communicator's max_n");
    }
    A->comm = (comm && comm->nranks > 1) ? comm : nullptr;
    if (A->comm) {
        int32_t rc = nupgcm_reserve_solver_workspace(ctx, comm->max_n);
        if (rc) return rc;
    }
    return nupgcm_csr_prepare(A, ctx->coop_grid);
}

extern "C" int32_t nupgcm_csr_shard_info(nupgcm_csr *A, int32_t rank, int64_t *row_begin, int64_t *row_end,
                                         int64_t *nnz_owned, int64_t *halo_rows) {
    NUPGCM_REQUIRE(nullptr, A, "csr is NULL");
    nupgcm_ctx *ctx = A->ctx;
    const int nranks = A->comm ? A->comm->nranks : 1;
    NUPGCM_REQUIRE(ctx, rank >= 0 && rank < nranks, "csr_shard_info: rank out of range");
    NUPGCM_REQUIRE(ctx, A->prepared_grid > 0 && A->prepared_ranks == nranks, "csr_shard_info: matrix not prepared");
    const int gpr = A->prepared_grid / nranks;
    std::vector<int32_t> h_rowptr(A->h_prow, A->h_prow + A->n_rows + 1), part;
    build_partition(h_rowptr, A->n_rows, A->prepared_grid, part);
    const int32_t q0 = part[(size_t)rank * gpr], q1 = part[(size_t)(rank + 1) * gpr];
    if (row_begin) *row_begin = q0;
    if (row_end) *row_end = q1;
    if (nnz_owned) *nnz_owned = h_rowptr[q1] - h_rowptr[q0];
    if (halo_rows) {
        // distinct columns outside the rank's own block
        std::vector<int32_t> cols;
        for (int32_t k = h_rowptr[q0]; k < h_rowptr[q1]; ++k) {
            const int32_t c = A->h_pcol[k];
            if (c < q0 || c >= q1) cols.push_back(c);
        }
        std::sort(cols.begin(), cols.end());
        *halo_rows = (int64_t)(std::unique(cols.begin(), cols.end()) - cols.begin());
    }
    return NUPGCM_OK;
}

// ---- C ABI --------------------------------------------------------------------------------
// Host-only: y = A x computed by walking the streaming tables exactly as the persistent kernels do
// (tiles, footprints staged in the arena, per-warp streams of jagged-diagonal slices), for `grid` CTAs
// on the structure as given (no reordering).  Lets the CPU tests validate the table builder without a
// device.  Also returns the number of tiles, of stream entries (padding included) and the number of
// shared-memory wavefronts the vector gathers of all positions need (2 per position = conflict-free).
extern "C" int32_t nupgcm_diag_stream_spmv_host(int64_t n, const int64_t *rowptr, const int64_t *colidx,
                                                const double *vals, const double *x, int32_t grid,
                                                int32_t arena, double *y, int64_t *n_tiles,
                                                int64_t *n_entries, int64_t *gather_wavefronts, int64_t *positions) {
    if (n < 1 || !rowptr || !colidx || !vals || !x || !y || grid < 1 || arena < 16 || arena % 16)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "diag_stream_spmv_host");
    const int64_t nnz = rowptr[n];
    if (nnz < 0 || nnz >= INT32_MAX || n >= INT32_MAX)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "diag_stream_spmv_host: size");
    std::vector<int32_t> rp(n + 1), col(nnz > 0 ? nnz : 1), part;
    for (int64_t i = 0; i <= n; ++i) rp[i] = (int32_t)rowptr[i];
    for (int64_t k = 0; k < nnz; ++k) col[k] = (int32_t)colidx[k];
    build_partition(rp, n, grid, part);
    StreamTables st;
    if (!build_stream_tables(rp.data(), col.data(), n, part, 0, grid, arena, st))
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "a tile's footprint exceeds the arena");
    std::vector<double> sv(st.scols.size());
    for (size_t i = 0; i < sv.size(); ++i) sv[i] = st.ssrc[i] >= 0 ? vals[st.ssrc[i]] : 0.0;
    for (int64_t i = 0; i < n; ++i) y[i] = std::nan("");         // every row must be written exactly once
    std::vector<double> xs(arena);
    std::vector<int> written(n, 0), owner(arena);
    const int W = kMainWarps;
    int64_t waves = 0, npos = 0;
    for (int b = 0; b < grid; ++b) {
        std::vector<int64_t> walked(W, 0);                       // a warp's slices must tile its stream in order
        std::vector<int> next_slice(W, 0);
        std::fill(owner.begin(), owner.end(), -1);
        const int nt = st.tile_ptr[b + 1] - st.tile_ptr[b];
        if (nt > kMaxTiles) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "too many tiles in a CTA");
        for (int tix = 0; tix < nt; ++tix) {
            const NupgcmTileDesc td = st.tiles[st.tile_ptr[b] + tix];
            if (td.xs_off % 16 || td.xs_off + td.foot_len > arena || td.dep >= tix)
                return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "bad arena slot");
            // staging tile tix overwrites its slot: every tile that still owns a part of it must be <= dep
            for (int i = 0; i < td.foot_len; ++i) {
                if (owner[td.xs_off + i] > td.dep) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "arena slot reused before its dependency");
                owner[td.xs_off + i] = tix;
                xs[td.xs_off + i] = x[st.foot[td.foot_off + i]];
            }
            for (int w = 0; w < W; ++w) {
                const NupgcmWarpDesc wd = st.wdesc[(size_t)b * W + w];
                if (wd.estart % 8) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "misaligned stream");
                for (int k = 0; k < st.twcnt[(size_t)(st.tile_ptr[b] + tix) * W + w]; ++k) {
                    NupgcmSlice sl = st.slices[(size_t)wd.stab + next_slice[w]++];
                    if (sl.eoff < walked[w] || sl.eoff > walked[w] + 7) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "slices do not tile the stream");
                    const int32_t *rows = st.srow.data() + wd.rtab + sl.roff, *lens = st.slen.data() + wd.rtab + sl.roff;
                    double acc[32] = {0};
                    int64_t off = sl.eoff;
                    const int nblk = sl.nrows >> 8;
                    sl.nrows &= 0xff;
                    if (sl.nrows < 0 || sl.nrows > 32 || lens[0] != sl.lmax || (nblk && (sl.nrows != 32 || sl.eoff % 8)))
                        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "bad slice header");
                    if (sl.nrows == 0) {                         // a long row: 32 consecutive entries per step
                        if (sl.lmax <= kLongRow) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "short row stored as a long one");
                        double sum = 0.0;
                        for (int k0 = 0; k0 < sl.lmax; k0 += 32) {
                            int load[2][16] = {{0}};
                            for (int q = 0; q < 32 && k0 + q < sl.lmax; ++q) {
                                const size_t e = (size_t)wd.estart + off + k0 + q;
                                if (off + k0 + q >= wd.elen || st.scols[e] % 8 || st.scols[e] / 8 >= td.foot_len)
                                    return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "entry outside its stream or footprint");
                                sum += sv[e] * xs[td.xs_off + st.scols[e] / 8];
                                load[q >> 4][(st.scols[e] / 8) & 15]++;
                            }
                            for (int hw = 0; hw < 2; ++hw) {
                                int mx = 0;
                                for (int bk = 0; bk < 16; ++bk) mx = std::max(mx, load[hw][bk]);
                                waves += mx;
                            }
                            ++npos;
                        }
                        walked[w] = off + sl.lmax;
                        if (rows[0] < td.row0 || rows[0] >= td.row0 + td.nrows) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "row outside its tile");
                        y[rows[0]] = sum;
                        written[rows[0]]++;
                        continue;
                    }
                    for (int bq = 0; bq < nblk; ++bq) {           // blocked part: [lane][8] offsets, [pair][lane][2] values
                        for (int pz = 0; pz < 8; ++pz) {
                            int load[2][16] = {{0}};
                            for (int q = 0; q < 32; ++q) {
                                const size_t base = (size_t)wd.estart + off;
                                const uint16_t c = st.scols[base + q * 8 + pz];
                                const double v = sv[base + (pz >> 1) * 64 + q * 2 + (pz & 1)];
                                if (lens[q] < 8 * nblk || off + 256 > wd.elen || c % 8 || c / 8 >= td.foot_len)
                                    return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "bad blocked entry");
                                acc[q] += v * xs[td.xs_off + c / 8];
                                load[q >> 4][(c / 8) & 15]++;
                            }
                            for (int hw = 0; hw < 2; ++hw) {
                                int mx = 0;
                                for (int bk = 0; bk < 16; ++bk) mx = std::max(mx, load[hw][bk]);
                                waves += mx;
                            }
                            ++npos;
                        }
                        off += 256;
                    }
                    for (int j = 8 * nblk; j < sl.lmax; ++j) {
                        int cnt = 0, load[2][16] = {{0}};
                        for (int q = 0; q < sl.nrows; ++q) {
                            if (q + 1 < sl.nrows && lens[q] < lens[q + 1]) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "slice rows not sorted");
                            if (lens[q] > j) {
                                const size_t e = (size_t)wd.estart + off + cnt;
                                if (off + cnt >= wd.elen || st.scols[e] % 8 || st.scols[e] / 8 >= td.foot_len)
                                    return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "entry outside its stream or footprint");
                                acc[q] += sv[e] * xs[td.xs_off + st.scols[e] / 8];
                                load[q >> 4][(st.scols[e] / 8) & 15]++;
                                ++cnt;
                            }
                        }
                        for (int hw = 0; hw < 2; ++hw) {
                            int mx = 0;
                            for (int bk = 0; bk < 16; ++bk) mx = std::max(mx, load[hw][bk]);
                            waves += mx;
                        }
                        ++npos;
                        off += cnt;
                    }
                    walked[w] = off;
                    for (int q = 0; q < sl.nrows; ++q) {
                        if (rows[q] < td.row0 || rows[q] >= td.row0 + td.nrows) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "row outside its tile");
                        y[rows[q]] = acc[q];
                        written[rows[q]]++;
                    }
                }
            }
        }
        for (int w = 0; w < W; ++w)
            if (walked[w] != st.wdesc[(size_t)b * W + w].elen) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "stream not consumed completely");
    }
    for (int64_t i = 0; i < n; ++i)
        if (written[i] != 1) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "%s", "a row was not written exactly once");
    if (n_tiles) *n_tiles = (int64_t)st.tiles.size();
    if (n_entries) *n_entries = (int64_t)st.scols.size();
    if (gather_wavefronts) *gather_wavefronts = waves;
    if (positions) *positions = npos;
    return NUPGCM_OK;
}

// Host-only utility: the reordering the solvers apply internally, for callers that want it too
// (the reference computes its per-field orderings with CuthillMcKee.symrcm, src/dofs.jl:98-100).
extern "C" int32_t nupgcm_rcm_order(int64_t n, const int64_t *rowptr, const int64_t *colidx,
                                    int32_t index_base, int64_t *perm_out) {
    if (n < 0 || !rowptr || !perm_out || (index_base != 0 && index_base != 1) || n >= INT32_MAX)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "rcm_order");
    const int64_t nnz = rowptr[n] - index_base;
    if (nnz < 0 || nnz >= INT32_MAX || (nnz > 0 && !colidx))
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "rcm_order: bad nnz");
    std::vector<int32_t> rp(n + 1), col(nnz > 0 ? nnz : 1), perm;
    for (int64_t i = 0; i <= n; ++i) rp[i] = (int32_t)(rowptr[i] - index_base);
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t c = colidx[k] - index_base;
        if (c < 0 || c >= n) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "rcm_order: column out of range");
        col[k] = (int32_t)c;
    }
    rcm_order(n, rp.data(), col.data(), perm);
    for (int64_t i = 0; i < n; ++i) perm_out[i] = perm[i] + index_base;
    return NUPGCM_OK;
}

// Host-only: how a sharded solve lays a matrix out over `nranks` ranks of `grid_per_rank` CTAs —
// the internal ordering, the row block of every rank and every (dst, src) halo range.  The solvers
// derive exactly this plan internally; it is exported so that hosts (and the CPU tests) can reason
// about ownership and halo sizes without a device.
extern "C" int32_t nupgcm_shard_plan(int64_t n, const int64_t *rowptr, const int64_t *colidx, int32_t index_base,
                                     int32_t nranks, int32_t grid_per_rank, int64_t *perm_out,
                                     int64_t *row_begin, int64_t *halo_lo, int64_t *halo_hi) {
    if (n < 1 || !rowptr || !colidx || (index_base != 0 && index_base != 1) || n >= INT32_MAX ||
        nranks < 1 || nranks > kMaxRanks || grid_per_rank < 1)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "shard_plan");
    const int64_t nnz = rowptr[n] - index_base;
    if (nnz < 0 || nnz >= INT32_MAX)
        return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "shard_plan: bad nnz");
    std::vector<int32_t> rp(n + 1), col(nnz > 0 ? nnz : 1), perm, prow, pcol, psrc, part;
    for (int64_t i = 0; i <= n; ++i) rp[i] = (int32_t)(rowptr[i] - index_base);
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t c = colidx[k] - index_base;
        if (c < 0 || c >= n) return nupgcm_fail(nullptr, NUPGCM_ERR_INVALID, "invalid argument: %s", "shard_plan: column out of range");
        col[k] = (int32_t)c;
    }
    rcm_order(n, rp.data(), col.data(), perm);
    permute_structure(n, rp.data(), col.data(), perm, prow, pcol, psrc);
    build_partition(prow, n, nranks * grid_per_rank, part);
    if (perm_out) for (int64_t i = 0; i < n; ++i) perm_out[i] = perm[i] + index_base;
    if (row_begin) for (int r = 0; r <= nranks; ++r) row_begin[r] = part[(size_t)r * grid_per_rank];
    for (int d = 0; d < nranks; ++d)
        for (int s_ = 0; s_ < nranks; ++s_) {
            int lo = 0, hi = 0;
            if (d != s_) halo_range(prow.data(), pcol.data(), part, grid_per_rank, d, s_, lo, hi);
            if (halo_lo) halo_lo[d * nranks + s_] = lo;
            if (halo_hi) halo_hi[d * nranks + s_] = hi;
        }
    return NUPGCM_OK;
}

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
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());   // set-up copies ran on the default stream
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
    cudaFree(A->d_perm);
    cudaFree(A->d_prow);
    cudaFree(A->d_pcol);
    cudaFree(A->d_psrc);
    cudaFree(A->d_pvals);
    free_stream_tables(A);
    cudaFree(A->d_halo_ptr);
    cudaFree(A->d_halo_idx);
    free(A->h_rowptr);
    free(A->h_col);
    free(A->h_prow);
    free(A->h_pcol);
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
    A->vals_version++;
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
    out->vals_version++;
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
