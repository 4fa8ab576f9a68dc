// Per-step element right-hand side of the buoyancy equation and the RHS combine.
//
// Replaces the CPU Gridap `assemble_vector(d -> advection_lform(...), B_test)` of reference
// src/model.jl:269-273 (forms :292-300), the `rhs_adv[perm]` + host->device copy of :274-275 and
// the broadcast of :278.  Two kernels, no atomics:
//   1. k_elem: one thread per cell evaluates the fields (velocity P2; buoyancy P2 or P1, `b_order` of
//      src/spaces.jl:31-39) at the quadrature points and writes the cell's elemental integrals — one per local
//      buoyancy DOF — to d_elem[i][cell] (tables are stored transposed so
//      that consecutive threads read consecutive addresses; u and b are gathered through the
//      cell's DOF indices, Dirichlet values living behind the free ones);
//   2. k_gather: one thread per free buoyancy DOF sums its elemental slots in a FIXED order
//      (sorted by cell id at set-up), so the result is independent of scheduling and GPU count.
#include <algorithm>
#include <vector>

#include "common.cuh"

template <int NV>   // vertices per cell: 4 (tetrahedra) or 3 (triangles embedded in 3-D)
struct P2 {
    static constexpr int NE = NV * (NV - 1) / 2;
    static constexpr int NLOC = NV + NE;
};

__device__ __constant__ int c_edge_a[6] = {0, 0, 1, 0, 1, 2};
__device__ __constant__ int c_edge_b[6] = {1, 2, 2, 3, 3, 3};

// NLB: local buoyancy DOFs per cell — P2<NV>::NLOC (reference default b_order = 2) or NV (b_order = 1,
// the production set-up of scratch/run.jl:152: φ_i = λ_i, ∇φ_i = ∇λ_i).  The velocity is always P2.
template <int NV, int NLB>
__global__ void __launch_bounds__(128)
k_elem(const int32_t *__restrict__ cell_b, const int32_t *__restrict__ cell_u,
       const double *__restrict__ grad, const double *__restrict__ vol,
       const double *__restrict__ bary, const double *__restrict__ w, int nq, int64_t n_cells,
       const double *__restrict__ b, const double *__restrict__ bp, const double *__restrict__ bdir,
       int64_t nb, const double *__restrict__ u, const double *__restrict__ up,
       const double *__restrict__ udir, int64_t nu, int scheme, double dt, double N2,
       double *__restrict__ elem) {
    constexpr int NLOC = P2<NV>::NLOC;
    constexpr int NE = P2<NV>::NE;
    constexpr bool BP1 = NLB == NV;
    static_assert(BP1 || NLB == NLOC, "buoyancy is P1 or P2");
    extern __shared__ double s_q[];               // bary[nq][NV], w[nq]
    for (int i = threadIdx.x; i < nq * NV; i += blockDim.x) s_q[i] = bary[i];
    for (int i = threadIdx.x; i < nq; i += blockDim.x) s_q[nq * NV + i] = w[i];
    __syncthreads();
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_cells) return;

    // gather the cell's fields: b* (advected), lin (time-derivative part), u* (advecting)
    double bs[NLB], lin[NLB], us[NLOC][3];
#pragma unroll
    for (int i = 0; i < NLB; ++i) {
        const int32_t ib = cell_b[i * n_cells + c];
        const double b0 = ib < nb ? b[ib] : bdir[ib - nb];
        const double b1 = ib < nb ? bp[ib] : bdir[ib - nb];
        if (scheme == 2) {
            bs[i] = 2.0 * b0 - b1;
            lin[i] = (4.0 / 3.0) * b0 - (1.0 / 3.0) * b1;
        } else {
            bs[i] = b0;
            lin[i] = b0;
        }
    }
#pragma unroll
    for (int i = 0; i < NLOC; ++i) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int32_t iu = cell_u[(i * 3 + d) * n_cells + c];
            const double u0 = iu < nu ? u[iu] : udir[iu - nu];
            const double u1 = iu < nu ? up[iu] : udir[iu - nu];
            us[i][d] = scheme == 2 ? 2.0 * u0 - u1 : u0;
        }
    }
    double gl[NV][3];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) gl[k][d] = grad[(k * 3 + d) * n_cells + c];
    const double fac = scheme == 2 ? (2.0 / 3.0) * dt : dt;
    const double vc = vol[c];

    double out[NLB];
#pragma unroll
    for (int i = 0; i < NLB; ++i) out[i] = 0.0;

    for (int q = 0; q < nq; ++q) {
        double lam[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) lam[k] = s_q[q * NV + k];
        // P2 basis values and gradients at this point
        double phi[NLOC];
        double gb[3] = {0.0, 0.0, 0.0}, uq[3] = {0.0, 0.0, 0.0}, lq = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            phi[i] = lam[i] * (2.0 * lam[i] - 1.0);
            const double dl = BP1 ? 1.0 : 4.0 * lam[i] - 1.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) gb[d] = fma(bs[i] * dl, gl[i][d], gb[d]);
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int ia = c_edge_a[e], ib = c_edge_b[e];
            phi[NV + e] = 4.0 * lam[ia] * lam[ib];
            if constexpr (!BP1) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    gb[d] = fma(4.0 * bs[NV + e], fma(lam[ia], gl[ib][d], lam[ib] * gl[ia][d]), gb[d]);
            }
        }
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) uq[d] = fma(phi[i], us[i][d], uq[d]);
        }
#pragma unroll
        for (int i = 0; i < NLB; ++i) lq = fma(BP1 ? lam[i] : phi[i], lin[i], lq);
        const double adv = uq[0] * gb[0] + uq[1] * gb[1] + uq[2] * gb[2] + uq[2] * N2;
        const double val = (lq - fac * adv) * (s_q[nq * NV + q] * vc);
#pragma unroll
        for (int i = 0; i < NLB; ++i) out[i] = fma(val, BP1 ? lam[i] : phi[i], out[i]);
    }
#pragma unroll
    for (int i = 0; i < NLB; ++i) elem[i * n_cells + c] = out[i];
}

// Adaptive timestep (reference src/timesteppers.jl:108-119): Δt = c · min_K h_K / max(|u|_{L∞(K)}, u_min),
// |u|_{L∞(K)} being the largest Euclidean speed over the cell's quadrature points.  One thread per
// cell; block minimum by shuffles, then atomicMin on the bit pattern (positive doubles order like
// unsigned integers; min is exact, so the result does not depend on the order).
template <int NV>
__global__ void __launch_bounds__(128)
k_cfl(const int32_t *__restrict__ cell_u, const double *__restrict__ bary, int nq, int64_t n_cells,
      const double *__restrict__ u, const double *__restrict__ udir, int64_t nu,
      const double *__restrict__ h_cells, double u_min, unsigned long long *__restrict__ min_bits) {
    constexpr int NLOC = P2<NV>::NLOC;
    constexpr int NE = P2<NV>::NE;
    extern __shared__ double s_q[];               // bary[nq][NV]
    __shared__ double s_min[4];
    for (int i = threadIdx.x; i < nq * NV; i += blockDim.x) s_q[i] = bary[i];
    __syncthreads();
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double ratio = 1.7976931348623157e308;
    if (c < n_cells) {
        double us[NLOC][3];
#pragma unroll
        for (int i = 0; i < NLOC; ++i)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int32_t iu = cell_u[(i * 3 + d) * n_cells + c];
                us[i][d] = iu < nu ? u[iu] : udir[iu - nu];
            }
        double smax = 0.0;
        for (int q = 0; q < nq; ++q) {
            double lam[NV], uq[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < NV; ++k) lam[k] = s_q[q * NV + k];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double ph = lam[i] * (2.0 * lam[i] - 1.0);
#pragma unroll
                for (int d = 0; d < 3; ++d) uq[d] = fma(ph, us[i][d], uq[d]);
            }
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const double ph = 4.0 * lam[c_edge_a[e]] * lam[c_edge_b[e]];
#pragma unroll
                for (int d = 0; d < 3; ++d) uq[d] = fma(ph, us[NV + e][d], uq[d]);
            }
            smax = fmax(smax, sqrt(uq[0] * uq[0] + uq[1] * uq[1] + uq[2] * uq[2]));
        }
        ratio = h_cells[c] / fmax(smax, u_min);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ratio = fmin(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = ratio;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = s_min[0];
        for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) m = fmin(m, s_min[wv]);
        atomicMin(min_bits, (unsigned long long)__double_as_longlong(m));
    }
}

// Convection parameterisation (reference src/model.jl:229-246, src/inputs.jl:87-91): every step
//   κᵥ(x_q) = κᵥ⁰(x_q) + κᶜ (1 + tanh(−α(N² + ∂z b)(x_q) / N²min)) / 2
// and Kᵥ = ∫ κᵥ ∂z b ∂z d, rhsᵥ = ∫ κᵥ ∂z b_diri ∂z d, rhs_diff = ∫ −N² κᵥ ∂z d are re-assembled
// (src/evolution.jl:243-246,256-260,269-278).  The reference does this on the CPU with Gridap and
// uploads the results; here one thread per cell writes the cell's 10x10 element matrix and two
// element vectors to [slot][cell] arrays, and gather kernels add each matrix entry's / DOF's
// slots in a fixed order (sorted by cell): no atomics, bitwise reproducible.
template <int NV, int NLOC>   // NLOC: local buoyancy DOFs (P2: NV + edges, P1: NV)
__global__ void __launch_bounds__(128)
k_kv_elem(const int32_t *__restrict__ cell_b, const double *__restrict__ grad, const double *__restrict__ vol,
          const double *__restrict__ bary, const double *__restrict__ w, int nq, int64_t n_cells,
          const double *__restrict__ b, const double *__restrict__ bdir, int64_t nb,
          const double *__restrict__ kv_q, double alpha, double N2, double kappa_c, double N2min,
          double *__restrict__ emat, double *__restrict__ evec_v, double *__restrict__ evec_d) {
    constexpr bool BP1 = NLOC == NV;
    static_assert(BP1 || NLOC == P2<NV>::NLOC, "buoyancy is P1 or P2");
    constexpr int MAXQ = 16;
    extern __shared__ double s_q[];               // bary[nq][NV], w[nq]
    for (int i = threadIdx.x; i < nq * NV; i += blockDim.x) s_q[i] = bary[i];
    for (int i = threadIdx.x; i < nq; i += blockDim.x) s_q[nq * NV + i] = w[i];
    __syncthreads();
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double bv[NLOC], bd[NLOC], gz[NV];
#pragma unroll
    for (int i = 0; i < NLOC; ++i) {
        const int32_t ib = cell_b[i * n_cells + c];
        bv[i] = ib < nb ? b[ib] : bdir[ib - nb];
        bd[i] = ib < nb ? 0.0 : bdir[ib - nb];        // b_diri: Dirichlet values, zero elsewhere
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) gz[k] = grad[(k * 3 + 2) * n_cells + c];
    const double vc = vol[c];
    // ∂z φ_i at quadrature point q
    auto dz = [&](int q, int i) -> double {
        if (BP1) return gz[i];
        const double *lam = s_q + q * NV;
        if (i < NV) return (4.0 * lam[i] - 1.0) * gz[i];
        const int ia = c_edge_a[i - NV], ib = c_edge_b[i - NV];
        return 4.0 * fma(lam[ia], gz[ib], lam[ib] * gz[ia]);
    };
    double kq[MAXQ], dbd[MAXQ];                       // w_q |K| κᵥ(x_q) and ∂z b_diri(x_q)
    for (int q = 0; q < nq; ++q) {
        double dzb = 0.0, dzd = 0.0;
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
            const double d = dz(q, i);
            dzb = fma(bv[i], d, dzb);
            dzd = fma(bd[i], d, dzd);
        }
        const double abz = alpha * (N2 + dzb);
        const double kap = kv_q[c * (int64_t)nq + q] + kappa_c * (1.0 + tanh(-abz / N2min)) * 0.5;
        kq[q] = s_q[nq * NV + q] * vc * kap;
        dbd[q] = dzd;
    }
    for (int i = 0; i < NLOC; ++i) {
        double row[NLOC], rv = 0.0, rd = 0.0;
#pragma unroll
        for (int j = 0; j < NLOC; ++j) row[j] = 0.0;
        for (int q = 0; q < nq; ++q) {
            const double di = kq[q] * dz(q, i);
            rv = fma(di, dbd[q], rv);
            rd = fma(di, -N2, rd);
#pragma unroll
            for (int j = 0; j < NLOC; ++j) row[j] = fma(di, dz(q, j), row[j]);
        }
#pragma unroll
        for (int j = 0; j < NLOC; ++j) emat[(size_t)(i * NLOC + j) * n_cells + c] = row[j];
        evec_v[(size_t)i * n_cells + c] = rv;
        evec_d[(size_t)i * n_cells + c] = rd;
    }
}

// Eddy parameterisation (reference src/model.jl:160-170, src/inputs.jl:130-137): every 10 steps
//   ν(x_q) = LogSumExp_s(ν_min, f² / sqrt(N²min² + (α(N² + ∂z b))²))
// and the friction block of the inversion matrix, ∫ 2α²ε² ν σ(u)⊙σ(v) (src/inversion.jl:172-182), is
// re-assembled; the reference rebuilds the whole matrix on the CPU with Gridap, permutes and uploads
// it.  Here the constant part (pressure gradient, divergence, Coriolis) stays on the device and one
// thread per cell writes the 3x3 blocks of the cell's 10x10 (6x6) node pairs:
//   block(i,j)[a][b] = α²ε² Σ_q w_q |K| ν_q (δ_ab ∇φ_i·∇φ_j + ∂_a φ_j ∂_b φ_i).
template <int NV, int NLB>   // NLB: local buoyancy DOFs (only ∂z b enters); the velocity blocks are P2
__global__ void __launch_bounds__(64)
k_nu_elem(const int32_t *__restrict__ cell_b, const double *__restrict__ grad, const double *__restrict__ vol,
          const double *__restrict__ bary, const double *__restrict__ w, int nq, int64_t n_cells,
          const double *__restrict__ b, const double *__restrict__ bdir, int64_t nb,
          const double *__restrict__ f_q, double a2e2, double alpha, double N2, double N2min, double smoothing,
          double nu_min, double *__restrict__ emat) {
    constexpr int NLOC = P2<NV>::NLOC;
    constexpr bool BP1 = NLB == NV;
    static_assert(BP1 || NLB == NLOC, "buoyancy is P1 or P2");
    constexpr int MAXQ = 16;
    extern __shared__ double s_q[];               // bary[nq][NV], w[nq]
    for (int i = threadIdx.x; i < nq * NV; i += blockDim.x) s_q[i] = bary[i];
    for (int i = threadIdx.x; i < nq; i += blockDim.x) s_q[nq * NV + i] = w[i];
    __syncthreads();
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double gl[NV][3];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) gl[k][d] = grad[(k * 3 + d) * n_cells + c];
    auto gphi = [&](int q, int i, double g[3]) {
        const double *lam = s_q + q * NV;
        if (i < NV) {
            const double t = 4.0 * lam[i] - 1.0;
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] = t * gl[i][d];
        } else {
            const int ia = c_edge_a[i - NV], ib = c_edge_b[i - NV];
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] = 4.0 * fma(lam[ia], gl[ib][d], lam[ib] * gl[ia][d]);
        }
    };
    double cq[MAXQ];                                  // α²ε² w_q |K| ν(x_q)
    {
        double bv[NLB];
#pragma unroll
        for (int i = 0; i < NLB; ++i) {
            const int32_t ib = cell_b[i * n_cells + c];
            bv[i] = ib < nb ? b[ib] : bdir[ib - nb];
        }
        const double vc = vol[c];
        for (int q = 0; q < nq; ++q) {
            double dzb = 0.0;
#pragma unroll
            for (int i = 0; i < NLB; ++i) {
                if constexpr (BP1) {
                    dzb = fma(bv[i], gl[i][2], dzb);
                } else {
                    double g[3];
                    gphi(q, i, g);
                    dzb = fma(bv[i], g[2], dzb);
                }
            }
            const double abz = alpha * (N2 + dzb);
            const double f = f_q[c * (int64_t)nq + q];
            const double nu_e = f * (f / sqrt(N2min * N2min + abz * abz));
            // LogSumExp, written so that the larger exponent is factored out (no overflow)
            const double hi = fmax(nu_min, nu_e), lo = fmin(nu_min, nu_e);
            const double nu = hi + log1p(exp(smoothing * (lo - hi))) / smoothing;
            cq[q] = a2e2 * s_q[nq * NV + q] * vc * nu;
        }
    }
    for (int i = 0; i < NLOC; ++i)
        for (int j = 0; j < NLOC; ++j) {
            double S = 0.0, T[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
            for (int q = 0; q < nq; ++q) {
                double gi[3], gj[3];
                gphi(q, i, gi);
                gphi(q, j, gj);
                const double cw = cq[q];
                S = fma(cw, gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2], S);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) T[a][bb] = fma(cw * gj[a], gi[bb], T[a][bb]);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb)
                    emat[(size_t)(((i * NLOC + j) * 3 + a) * 3 + bb) * n_cells + c] = T[a][bb] + (a == bb ? S : 0.0);
        }
}

// A.vals[e] = A0[e] + (sum of the entry's friction slots, fixed order)
__global__ void k_gather_mat_add(const int32_t *__restrict__ kptr, const int32_t *__restrict__ kidx,
                                 const double *__restrict__ emat, const double *__restrict__ base,
                                 double *__restrict__ vals, int64_t nnz) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz;
         e += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int32_t k = kptr[e]; k < kptr[e + 1]; ++k) acc += emat[kidx[k]];
        vals[e] = base[e] + acc;
    }
}

// one thread per stored matrix entry: add its element-matrix slots in the fixed (cell) order
__global__ void k_gather_mat(const int32_t *__restrict__ kptr, const int32_t *__restrict__ kidx,
                             const double *__restrict__ emat, double *__restrict__ vals, int64_t nnz) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz;
         e += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int32_t k = kptr[e]; k < kptr[e + 1]; ++k) acc += emat[kidx[k]];
        vals[e] = acc;
    }
}

__global__ void k_gather_elem(const int32_t *__restrict__ gptr, const int32_t *__restrict__ gidx,
                              const double *__restrict__ elem, double *__restrict__ out, int64_t nb) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb;
         i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int32_t k = gptr[i]; k < gptr[i + 1]; ++k) acc += elem[gidx[k]];
        out[i] = acc;
    }
}

__global__ void k_rhs_combine(double *out, const double *adv, double theta, double dt,
                              const double *diff, const double *flux, const double *rm,
                              const double *rh, const double *rv, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        // rhs_adv + θ*rhs_diff + Δt*rhs_flux − (rhsₘ + θ*(rhsₕ + rhsᵥ)), model.jl:278
        out[i] = adv[i] + theta * diff[i] + dt * flux[i] - (rm[i] + theta * (rh[i] + rv[i]));
}

// ---- C ABI --------------------------------------------------------------------------------

template <class Tp>
static cudaError_t upload(Tp **dst, const std::vector<Tp> &src) {
    cudaError_t e = cudaMalloc(dst, (src.size() ? src.size() : 1) * sizeof(Tp));
    if (e != cudaSuccess) return e;
    if (src.size()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice);
    return e;
}

// Launch `KERNEL<NV, NLB>` for the mesh's cell type and buoyancy order.
#define NUPGCM_MESH_DISPATCH(m, KERNEL, GRID, BLOCK, SMEM, STREAM, ...)                                    \
    do {                                                                                                   \
        if ((m)->n_vert == 4 && (m)->n_loc_b == 10) KERNEL<4, 10><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); \
        else if ((m)->n_vert == 4) KERNEL<4, 4><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                \
        else if ((m)->n_loc_b == 6) KERNEL<3, 6><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);               \
        else KERNEL<3, 3><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                                      \
    } while (0)

extern "C" int32_t nupgcm_mesh_create(nupgcm_ctx *ctx, int64_t n_cells, int32_t n_loc,
                                      const int32_t *cell_b, const int32_t *cell_u,
                                      const double *grad, const double *vol, int32_t nq,
                                      const double *bary, const double *w, int64_t nb,
                                      const double *b_dirichlet, int64_t nbd, int64_t nu,
                                      const double *u_dirichlet, int64_t nud, nupgcm_mesh **out) {
    return nupgcm_mesh_create_orders(ctx, n_cells, n_loc, n_loc, cell_b, cell_u, grad, vol, nq, bary, w, nb,
                                     b_dirichlet, nbd, nu, u_dirichlet, nud, out);
}

extern "C" int32_t nupgcm_mesh_create_orders(nupgcm_ctx *ctx, int64_t n_cells, int32_t n_loc, int32_t n_loc_b,
                                             const int32_t *cell_b, const int32_t *cell_u,
                                             const double *grad, const double *vol, int32_t nq,
                                             const double *bary, const double *w, int64_t nb,
                                             const double *b_dirichlet, int64_t nbd, int64_t nu,
                                             const double *u_dirichlet, int64_t nud, nupgcm_mesh **out) {
    NUPGCM_REQUIRE(nullptr, ctx, "ctx is NULL");
    NUPGCM_REQUIRE(ctx, out && cell_b && cell_u && grad && vol && bary && w, "mesh_create: NULL argument");
    NUPGCM_REQUIRE(ctx, n_loc == 10 || n_loc == 6, "mesh_create: n_loc must be 10 (tets) or 6 (triangles)");
    NUPGCM_REQUIRE(ctx, n_loc_b == n_loc || n_loc_b == (n_loc == 10 ? 4 : 3),
                   "mesh_create: n_loc_b must be n_loc (P2 buoyancy) or the vertex count (P1 buoyancy)");
    NUPGCM_REQUIRE(ctx, n_cells > 0 && nq > 0 && nq <= 64, "mesh_create: bad n_cells or nq");
    NUPGCM_REQUIRE(ctx, nb >= 0 && nbd >= 0 && nu >= 0 && nud >= 0, "mesh_create: negative size");
    NUPGCM_REQUIRE(ctx, (nbd == 0 || b_dirichlet) && (nud == 0 || u_dirichlet), "mesh_create: NULL Dirichlet values");
    NUPGCM_REQUIRE(ctx, n_cells * n_loc * 3 < INT32_MAX, "mesh_create: mesh too large for int32 slots");
    const int nv = n_loc == 10 ? 4 : 3;
    // transpose tables to [local][cell]; validate indices; build the per-DOF gather lists
    std::vector<int32_t> tb((size_t)n_cells * n_loc_b), tu((size_t)n_cells * n_loc * 3);
    std::vector<double> tg((size_t)n_cells * nv * 3);
    std::vector<int32_t> gptr(nb + 1, 0);
    for (int64_t c = 0; c < n_cells; ++c) {
        for (int i = 0; i < n_loc_b; ++i) {
            const int32_t ib = cell_b[c * n_loc_b + i];
            if (ib < 0 || ib >= nb + nbd)
                return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s", "mesh_create: cell_b index out of range");
            tb[(size_t)i * n_cells + c] = ib;
            if (ib < nb) gptr[ib + 1]++;
        }
        for (int i = 0; i < n_loc; ++i) {
            for (int d = 0; d < 3; ++d) {
                const int32_t iu = cell_u[(c * n_loc + i) * 3 + d];
                if (iu < 0 || iu >= nu + nud)
                    return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s", "mesh_create: cell_u index out of range");
                tu[((size_t)i * 3 + d) * n_cells + c] = iu;
            }
        }
        for (int k = 0; k < nv; ++k)
            for (int d = 0; d < 3; ++d) tg[((size_t)k * 3 + d) * n_cells + c] = grad[(c * nv + k) * 3 + d];
    }
    for (int64_t i = 0; i < nb; ++i) gptr[i + 1] += gptr[i];
    std::vector<int32_t> gidx(gptr[nb]), fill(gptr.begin(), gptr.end() - 1);
    for (int64_t c = 0; c < n_cells; ++c)          // cells in order -> lists sorted by cell id
        for (int i = 0; i < n_loc_b; ++i) {
            const int32_t ib = cell_b[c * n_loc_b + i];
            if (ib < nb) gidx[fill[ib]++] = (int32_t)((int64_t)i * n_cells + c);
        }
    nupgcm_mesh *m = (nupgcm_mesh *)calloc(1, sizeof(nupgcm_mesh));
    if (!m) return nupgcm_fail(ctx, NUPGCM_ERR_ALLOC, "%s", "host allocation failed");
    m->ctx = ctx;
    m->n_cells = n_cells;
    m->n_loc = n_loc;
    m->n_loc_b = n_loc_b;
    m->n_vert = nv;
    m->nq = nq;
    m->nb = nb;
    m->nbd = nbd;
    m->nu = nu;
    m->nud = nud;
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, upload(&m->d_cell_b, tb));
    NUPGCM_CUDA(ctx, upload(&m->d_cell_u, tu));
    NUPGCM_CUDA(ctx, upload(&m->d_grad, tg));
    NUPGCM_CUDA(ctx, upload(&m->d_vol, std::vector<double>(vol, vol + n_cells)));
    NUPGCM_CUDA(ctx, upload(&m->d_bdir, std::vector<double>(b_dirichlet, b_dirichlet + nbd)));
    NUPGCM_CUDA(ctx, upload(&m->d_udir, std::vector<double>(u_dirichlet, u_dirichlet + nud)));
    NUPGCM_CUDA(ctx, upload(&m->d_gptr, gptr));
    NUPGCM_CUDA(ctx, upload(&m->d_gidx, gidx));
    NUPGCM_CUDA(ctx, upload(&m->d_phi, std::vector<double>(bary, bary + (size_t)nq * nv)));   // bary
    NUPGCM_CUDA(ctx, upload(&m->d_w, std::vector<double>(w, w + nq)));
    NUPGCM_CUDA(ctx, cudaMalloc(&m->d_elem, (size_t)n_cells * n_loc_b * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());   // set-up copies ran on the default stream
    *out = m;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_mesh_destroy(nupgcm_mesh *m) {
    if (!m) return NUPGCM_OK;
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_cell_b);
    cudaFree(m->d_cell_u);
    cudaFree(m->d_grad);
    cudaFree(m->d_vol);
    cudaFree(m->d_bdir);
    cudaFree(m->d_udir);
    cudaFree(m->d_gptr);
    cudaFree(m->d_gidx);
    cudaFree(m->d_phi);
    cudaFree(m->d_w);
    cudaFree(m->d_elem);
    cudaFree(m->d_hcells);
    cudaFree(m->d_minbits);
    cudaFree(m->d_kptr);
    cudaFree(m->d_kidx);
    cudaFree(m->d_kvq);
    cudaFree(m->d_emat);
    cudaFree(m->d_evec);
    cudaFree(m->d_nptr);
    cudaFree(m->d_nidx);
    cudaFree(m->d_fq);
    cudaFree(m->d_nmat);
    cudaFree(m->d_A0);
    free(m);
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_rhs_adv(nupgcm_mesh *m, int32_t scheme, double dt, double N2,
                                  const nupgcm_vec *b, const nupgcm_vec *b_prev, const nupgcm_vec *u,
                                  const nupgcm_vec *u_prev, nupgcm_vec *out) {
    NUPGCM_REQUIRE(nullptr, m && b && b_prev && u && u_prev && out, "rhs_adv: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, scheme == 1 || scheme == 2, "rhs_adv: scheme must be 1 (BDF1) or 2 (BDF2)");
    NUPGCM_REQUIRE(ctx, b->n == m->nb && b_prev->n == m->nb && out->n == m->nb, "rhs_adv: buoyancy length mismatch");
    NUPGCM_REQUIRE(ctx, u->n >= m->nu && u_prev->n >= m->nu, "rhs_adv: velocity vector shorter than nu");
    NUPGCM_REQUIRE(ctx, out->d != b->d && out->d != b_prev->d, "rhs_adv: out must not alias b");
    const int block = 128;
    const int grid = (int)((m->n_cells + block - 1) / block);
    const size_t smem = (size_t)m->nq * (m->n_vert + 1) * sizeof(double);
    NUPGCM_MESH_DISPATCH(m, k_elem, grid, block, smem, ctx->stream, m->d_cell_b, m->d_cell_u, m->d_grad, m->d_vol, m->d_phi,
                         m->d_w, m->nq, m->n_cells, b->d, b_prev->d, m->d_bdir, m->nb, u->d, u_prev->d, m->d_udir, m->nu,
                         scheme, dt, N2, m->d_elem);
    NUPGCM_CUDA(ctx, cudaGetLastError());
    if (m->nb > 0) {
        int g2 = (int)((m->nb + 255) / 256);
        if (g2 > ctx->sm_count * 8) g2 = ctx->sm_count * 8;
        k_gather_elem<<<g2, 256, 0, ctx->stream>>>(m->d_gptr, m->d_gidx, m->d_elem, out->d, m->nb);
        NUPGCM_CUDA(ctx, cudaGetLastError());
    }
    ctx->launches += 2;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_mesh_set_cell_sizes(nupgcm_mesh *m, const double *h_cells, int64_t n_cells) {
    NUPGCM_REQUIRE(nullptr, m, "mesh is NULL");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, h_cells && n_cells == m->n_cells, "mesh_set_cell_sizes: NULL sizes or cell count mismatch");
    for (int64_t c = 0; c < n_cells; ++c)
        if (!(h_cells[c] > 0.0)) return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s", "mesh_set_cell_sizes: sizes must be positive");
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!m->d_hcells) NUPGCM_CUDA(ctx, cudaMalloc(&m->d_hcells, (size_t)n_cells * sizeof(double)));
    if (!m->d_minbits) NUPGCM_CUDA(ctx, cudaMalloc(&m->d_minbits, sizeof(unsigned long long)));
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(m->d_hcells, h_cells, (size_t)n_cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_cfl_dt(nupgcm_mesh *m, const nupgcm_vec *u, double cfl_factor, double u_min,
                                 double *dt_out) {
    NUPGCM_REQUIRE(nullptr, m && u && dt_out, "cfl_dt: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, m->d_hcells, "cfl_dt: call nupgcm_mesh_set_cell_sizes first");
    NUPGCM_REQUIRE(ctx, u->n >= m->nu, "cfl_dt: velocity vector shorter than nu");
    NUPGCM_REQUIRE(ctx, cfl_factor > 0.0 && u_min > 0.0, "cfl_dt: cfl_factor and u_min must be positive");
    const unsigned long long inf_bits = 0x7fefffffffffffffULL;     // largest finite double
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(m->d_minbits, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, ctx->stream));
    const int block = 128;
    const int grid = (int)((m->n_cells + block - 1) / block);
    const size_t smem = (size_t)m->nq * m->n_vert * sizeof(double);
    if (m->n_vert == 4)
        k_cfl<4><<<grid, block, smem, ctx->stream>>>(m->d_cell_u, m->d_phi, m->nq, m->n_cells, u->d, m->d_udir, m->nu, m->d_hcells, u_min, m->d_minbits);
    else
        k_cfl<3><<<grid, block, smem, ctx->stream>>>(m->d_cell_u, m->d_phi, m->nq, m->n_cells, u->d, m->d_udir, m->nu, m->d_hcells, u_min, m->d_minbits);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    unsigned long long bits = 0;
    NUPGCM_CUDA(ctx, cudaMemcpyAsync(&bits, m->d_minbits, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    NUPGCM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double r;
    memcpy(&r, &bits, sizeof(r));
    *dt_out = cfl_factor * r;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_mesh_enable_kv_rebuild(nupgcm_mesh *m, const nupgcm_csr *pattern, const double *kv_q) {
    NUPGCM_REQUIRE(nullptr, m && pattern, "mesh_enable_kv_rebuild: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, kv_q, "mesh_enable_kv_rebuild: NULL kv_q");
    NUPGCM_REQUIRE(ctx, !pattern->dropped && pattern->n_rows == m->nb && pattern->n_cols == m->nb,
                   "mesh_enable_kv_rebuild: pattern must be the nb x nb evolution pattern created with drop_zeros=0");
    NUPGCM_REQUIRE(ctx, m->nq <= 16, "mesh_enable_kv_rebuild: at most 16 quadrature points");
    const int64_t nc = m->n_cells, nnz = pattern->nnz;
    const int nl = m->n_loc_b;
    NUPGCM_REQUIRE(ctx, (int64_t)nl * nl * nc < INT32_MAX, "mesh_enable_kv_rebuild: mesh too large for int32 slots");
    // host copy of the transposed cell_b table
    std::vector<int32_t> tb((size_t)nc * nl);
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, cudaMemcpy(tb.data(), m->d_cell_b, tb.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int32_t *rp = pattern->h_rowptr, *col = pattern->h_col;
    std::vector<int32_t> kptr(nnz + 1, 0), pos((size_t)nc * nl * nl, -1);
    for (int64_t c = 0; c < nc; ++c)
        for (int i = 0; i < nl; ++i) {
            const int32_t r = tb[(size_t)i * nc + c];
            if (r >= m->nb) continue;
            for (int j = 0; j < nl; ++j) {
                const int32_t cc = tb[(size_t)j * nc + c];
                if (cc >= m->nb) continue;
                const int32_t *lo = std::lower_bound(col + rp[r], col + rp[r + 1], cc);
                if (lo == col + rp[r + 1] || *lo != cc)
                    return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s",
                                       "mesh_enable_kv_rebuild: a cell couples two DOFs that the pattern does not store");
                const int32_t e = (int32_t)(lo - col);
                pos[((size_t)c * nl + i) * nl + j] = e;
                kptr[e + 1]++;
            }
        }
    for (int64_t e = 0; e < nnz; ++e) kptr[e + 1] += kptr[e];
    std::vector<int32_t> kidx(kptr[nnz]), fill(kptr.begin(), kptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)               // cells in order -> each entry's slots sorted by cell
        for (int i = 0; i < nl; ++i)
            for (int j = 0; j < nl; ++j) {
                const int32_t e = pos[((size_t)c * nl + i) * nl + j];
                if (e >= 0) kidx[fill[e]++] = (int32_t)((int64_t)(i * nl + j) * nc + c);
            }
    cudaFree(m->d_kptr); cudaFree(m->d_kidx); cudaFree(m->d_kvq); cudaFree(m->d_emat); cudaFree(m->d_evec);
    m->d_kptr = m->d_kidx = nullptr; m->d_kvq = m->d_emat = m->d_evec = nullptr;
    NUPGCM_CUDA(ctx, upload(&m->d_kptr, kptr));
    NUPGCM_CUDA(ctx, upload(&m->d_kidx, kidx));
    NUPGCM_CUDA(ctx, upload(&m->d_kvq, std::vector<double>(kv_q, kv_q + (size_t)nc * m->nq)));
    NUPGCM_CUDA(ctx, cudaMalloc(&m->d_emat, (size_t)nc * nl * nl * sizeof(double)));
    NUPGCM_CUDA(ctx, cudaMalloc(&m->d_evec, 2 * (size_t)nc * nl * sizeof(double)));
    m->kv_nnz = nnz;
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_rebuild_kv(nupgcm_mesh *m, double alpha, double N2, double kappa_c, double N2min,
                                     const nupgcm_vec *b, nupgcm_csr *Kv, nupgcm_vec *rhs_v, nupgcm_vec *rhs_diff) {
    NUPGCM_REQUIRE(nullptr, m && b && Kv && rhs_v && rhs_diff, "rebuild_kv: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, m->d_kptr, "rebuild_kv: call nupgcm_mesh_enable_kv_rebuild first");
    NUPGCM_REQUIRE(ctx, !Kv->dropped && Kv->nnz == m->kv_nnz && Kv->n_rows == m->nb, "rebuild_kv: Kv does not have the bound pattern");
    NUPGCM_REQUIRE(ctx, b->n == m->nb && rhs_v->n == m->nb && rhs_diff->n == m->nb, "rebuild_kv: vector length mismatch");
    NUPGCM_REQUIRE(ctx, N2min != 0.0, "rebuild_kv: N2min must be non-zero");
    const int block = 128;
    const int grid = (int)((m->n_cells + block - 1) / block);
    const size_t smem = (size_t)m->nq * (m->n_vert + 1) * sizeof(double);
    double *ev = m->d_evec, *ed = m->d_evec + (size_t)m->n_cells * m->n_loc_b;
    NUPGCM_MESH_DISPATCH(m, k_kv_elem, grid, block, smem, ctx->stream, m->d_cell_b, m->d_grad, m->d_vol, m->d_phi, m->d_w,
                         m->nq, m->n_cells, b->d, m->d_bdir, m->nb, m->d_kvq, alpha, N2, kappa_c, N2min, m->d_emat, ev, ed);
    NUPGCM_CUDA(ctx, cudaGetLastError());
    int g = (int)std::min<int64_t>((m->kv_nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
    if (g < 1) g = 1;
    k_gather_mat<<<g, 256, 0, ctx->stream>>>(m->d_kptr, m->d_kidx, m->d_emat, Kv->d_vals, m->kv_nnz);
    NUPGCM_CUDA(ctx, cudaGetLastError());
    Kv->vals_version++;
    if (m->nb > 0) {
        int g2 = (int)std::min<int64_t>((m->nb + 255) / 256, (int64_t)ctx->sm_count * 8);
        k_gather_elem<<<g2, 256, 0, ctx->stream>>>(m->d_gptr, m->d_gidx, ev, rhs_v->d, m->nb);
        k_gather_elem<<<g2, 256, 0, ctx->stream>>>(m->d_gptr, m->d_gidx, ed, rhs_diff->d, m->nb);
        NUPGCM_CUDA(ctx, cudaGetLastError());
    }
    ctx->launches += 4;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_mesh_enable_nu_rebuild(nupgcm_mesh *m, const nupgcm_csr *A, const double *A0_vals,
                                                const double *f_q) {
    NUPGCM_REQUIRE(nullptr, m && A, "mesh_enable_nu_rebuild: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, A0_vals && f_q, "mesh_enable_nu_rebuild: NULL A0_vals or f_q");
    NUPGCM_REQUIRE(ctx, !A->dropped && A->n_rows == A->n_cols && A->n_rows >= m->nu,
                   "mesh_enable_nu_rebuild: A must be the square inversion matrix created with drop_zeros=0");
    NUPGCM_REQUIRE(ctx, m->nq <= 16, "mesh_enable_nu_rebuild: at most 16 quadrature points");
    const int64_t nc = m->n_cells, nnz = A->nnz;
    const int nl = m->n_loc, nd = nl * 3;
    NUPGCM_REQUIRE(ctx, (int64_t)nd * nd * nc < INT32_MAX, "mesh_enable_nu_rebuild: mesh too large for int32 slots");
    std::vector<int32_t> tu((size_t)nc * nd);
    NUPGCM_CUDA(ctx, cudaSetDevice(ctx->device));
    NUPGCM_CUDA(ctx, cudaMemcpy(tu.data(), m->d_cell_u, tu.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int32_t *rp = A->h_rowptr, *col = A->h_col;
    std::vector<int32_t> kptr(nnz + 1, 0);
    std::vector<int32_t> pos((size_t)nc * nd * nd, -1);
    // local velocity DOF (i, a) of cell c is tu[(i*3 + a)*nc + c]; free ones (< nu) are rows/columns of A
    for (int64_t c = 0; c < nc; ++c)
        for (int ia = 0; ia < nd; ++ia) {
            const int32_t r = tu[(size_t)ia * nc + c];
            if (r >= m->nu) continue;
            for (int jb = 0; jb < nd; ++jb) {
                const int32_t cc = tu[(size_t)jb * nc + c];
                if (cc >= m->nu) continue;
                const int32_t *lo = std::lower_bound(col + rp[r], col + rp[r + 1], cc);
                if (lo == col + rp[r + 1] || *lo != cc)
                    return nupgcm_fail(ctx, NUPGCM_ERR_INVALID, "invalid argument: %s",
                                       "mesh_enable_nu_rebuild: a cell couples two velocity DOFs that A does not store");
                const int32_t e = (int32_t)(lo - col);
                pos[((size_t)c * nd + ia) * nd + jb] = e;
                kptr[e + 1]++;
            }
        }
    for (int64_t e = 0; e < nnz; ++e) kptr[e + 1] += kptr[e];
    std::vector<int32_t> kidx(kptr[nnz]), fill(kptr.begin(), kptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
        for (int ia = 0; ia < nd; ++ia)
            for (int jb = 0; jb < nd; ++jb) {
                const int32_t e = pos[((size_t)c * nd + ia) * nd + jb];
                if (e < 0) continue;
                const int i = ia / 3, a = ia % 3, j = jb / 3, b = jb % 3;
                kidx[fill[e]++] = (int32_t)((int64_t)(((i * nl + j) * 3 + a) * 3 + b) * nc + c);
            }
    cudaFree(m->d_nptr); cudaFree(m->d_nidx); cudaFree(m->d_fq); cudaFree(m->d_nmat); cudaFree(m->d_A0);
    m->d_nptr = m->d_nidx = nullptr; m->d_fq = m->d_nmat = m->d_A0 = nullptr;
    NUPGCM_CUDA(ctx, upload(&m->d_nptr, kptr));
    NUPGCM_CUDA(ctx, upload(&m->d_nidx, kidx));
    NUPGCM_CUDA(ctx, upload(&m->d_fq, std::vector<double>(f_q, f_q + (size_t)nc * m->nq)));
    NUPGCM_CUDA(ctx, upload(&m->d_A0, std::vector<double>(A0_vals, A0_vals + nnz)));
    NUPGCM_CUDA(ctx, cudaMalloc(&m->d_nmat, (size_t)nc * nd * nd * sizeof(double)));
    m->nu_nnz = nnz;
    NUPGCM_CUDA(ctx, cudaDeviceSynchronize());
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_rebuild_friction(nupgcm_mesh *m, double a2e2, double alpha, double N2, double N2min,
                                             double smoothing, double nu_min, const nupgcm_vec *b, nupgcm_csr *A) {
    NUPGCM_REQUIRE(nullptr, m && b && A, "rebuild_friction: NULL argument");
    nupgcm_ctx *ctx = m->ctx;
    NUPGCM_REQUIRE(ctx, m->d_nptr, "rebuild_friction: call nupgcm_mesh_enable_nu_rebuild first");
    NUPGCM_REQUIRE(ctx, !A->dropped && A->nnz == m->nu_nnz, "rebuild_friction: A does not have the bound pattern");
    NUPGCM_REQUIRE(ctx, b->n == m->nb, "rebuild_friction: buoyancy length mismatch");
    NUPGCM_REQUIRE(ctx, smoothing > 0.0, "rebuild_friction: smoothing must be positive");
    const int block = 64;
    const int grid = (int)((m->n_cells + block - 1) / block);
    const size_t smem = (size_t)m->nq * (m->n_vert + 1) * sizeof(double);
    NUPGCM_MESH_DISPATCH(m, k_nu_elem, grid, block, smem, ctx->stream, m->d_cell_b, m->d_grad, m->d_vol, m->d_phi, m->d_w,
                         m->nq, m->n_cells, b->d, m->d_bdir, m->nb, m->d_fq, a2e2, alpha, N2, N2min, smoothing, nu_min, m->d_nmat);
    NUPGCM_CUDA(ctx, cudaGetLastError());
    int g = (int)std::min<int64_t>((m->nu_nnz + 255) / 256, (int64_t)ctx->sm_count * 8);
    if (g < 1) g = 1;
    k_gather_mat_add<<<g, 256, 0, ctx->stream>>>(m->d_nptr, m->d_nidx, m->d_nmat, m->d_A0, A->d_vals, m->nu_nnz);
    NUPGCM_CUDA(ctx, cudaGetLastError());
    A->vals_version++;
    ctx->launches += 2;
    return NUPGCM_OK;
}

extern "C" int32_t nupgcm_rhs_combine(nupgcm_vec *out, const nupgcm_vec *rhs_adv, double theta,
                                      double dt, const nupgcm_vec *rhs_diff, const nupgcm_vec *rhs_flux,
                                      const nupgcm_vec *rhs_m, const nupgcm_vec *rhs_h,
                                      const nupgcm_vec *rhs_v) {
    NUPGCM_REQUIRE(nullptr, out && rhs_adv && rhs_diff && rhs_flux && rhs_m && rhs_h && rhs_v, "rhs_combine: NULL argument");
    nupgcm_ctx *ctx = out->ctx;
    const int64_t n = out->n;
    NUPGCM_REQUIRE(ctx, rhs_adv->n == n && rhs_diff->n == n && rhs_flux->n == n && rhs_m->n == n && rhs_h->n == n && rhs_v->n == n, "rhs_combine: length mismatch");
    if (n == 0) return NUPGCM_OK;
    int g = (int)((n + 255) / 256);
    if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
    k_rhs_combine<<<g, 256, 0, ctx->stream>>>(out->d, rhs_adv->d, theta, dt, rhs_diff->d, rhs_flux->d, rhs_m->d, rhs_h->d, rhs_v->d, n);
    ctx->launches++;
    NUPGCM_CUDA(ctx, cudaGetLastError());
    return NUPGCM_OK;
}
