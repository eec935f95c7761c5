// mg.cuh -- aggregation multigrid V-cycle used as the preconditioner of CG (SURVEY 8f rank 1:
// the reference's own algorithm class -- it preconditions CG with AlgebraicMultigrid.ruge_stuben,
// src/FiniteVolume.jl:160 -- where Jacobi needs O(n) iterations on an n^3 grid).
//
// Available when the assembled matrix is a 7-point stencil on a box (the diagonal format of
// dia.cuh with offsets 1, nz, ny*nz and no wrap-around couplings: regulargrid numbering with the
// Dirichlet nodes forming whole x-planes).  Everything stays index-free:
//   * coarsening: 2x2x2 aggregates of nodes, piecewise-constant prolongation P;
//   * coarse operators: Galerkin P^T A P, which for box aggregates is again a 7-point stencil --
//     off-diagonal = sum of the fine couplings crossing the common face, diagonal = sum of the fine
//     diagonals plus twice the couplings inside the aggregate (k_mg_coarsen);
//   * smoother: nu damped-Jacobi sweeps (omega) before and after, the first pre-sweep from a zero
//     guess being a pure scaling; same count on both sides keeps the cycle symmetric, so M^-1 is
//     SPD and legal inside CG;
//   * coarse-grid correction scaled by `oc` (over-correction; plain aggregation under-corrects
//     smooth error -- measured 38 -> 20 PCG iterations at 128^3 with oc = 1.5);
//   * the coarsest level (<= kMgCoarsest unknowns) is "solved" by a fixed number of sweeps in one CTA.
// Levels store only upper couplings up_k[r] = A[r, r+o_k] (the lower one is up_k[r-o_k]).  On level 0
// they alias the handle's diagonal copy.
// Multi-GPU (x-slabs): aggregates never straddle ranks, so every level is again slab-partitioned with
// the same (y,z) plane on all ranks.  The iterate buffers carry one halo plane per side
// ([owned n | plane from the rank below | plane from the rank above]); the coupling of a first-plane
// row to the rank below is lo_c[r] (level 0: the diagonal copy's prefix; coarser: sum over the
// children, k_mg_coarsen), the coupling of a last-plane row to the rank above is its up[2][r].  The
// host exchanges the boundary planes before every operator application (4 per level and V-cycle at
// nu = 2), so the distributed V-cycle is the same operator as the single-GPU one up to the
// rank-aligned aggregate boundaries and the inexact coarsest solve.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace fvb {

constexpr int kMgMaxLevels = 12;
constexpr int kMgCoarsest = 256;     // stop coarsening at or below this many unknowns
constexpr int kMgCoarseSweeps = 24;  // damped-Jacobi sweeps on the coarsest level (fixed => linear operator; even)
constexpr int kMgCtasPerSm = 4;

struct MgLevel {
  int nx, ny, nz;        // box of unknowns, z fastest
  int64_t n;
  const double *diag;    // [n]
  const double *up[3];   // up[k][r] = A[r, r + o_k], o = (1, nz, ny*nz); may be read at r - o_k >= 0
  double *x, *r, *t;     // correction, right-hand side, scratch (ping-pong for Jacobi); x,t: n + 2*ny*nz
  const double *lo_c;    // [ny*nz] couplings of the first plane to the rank below, or null
  int has_hi;            // the last plane's up[2] couples to the rank above (halo at v[n + ny*nz + j])
};

// (A v)[r] for the level's stencil, neighbours outside [0,n) skipped
__device__ __forceinline__ double mg_apply(const MgLevel &L, const double *__restrict__ v, int64_t r, double vr) {
  const int64_t o1 = L.nz, o2 = (int64_t)L.ny * L.nz;
  double acc = L.diag[r] * vr;
  if (r >= 1) acc += L.up[0][r - 1] * v[r - 1];
  if (r + 1 < L.n) acc += L.up[0][r] * v[r + 1];
  if (r >= o1) acc += L.up[1][r - o1] * v[r - o1];
  if (r + o1 < L.n) acc += L.up[1][r] * v[r + o1];
  if (r >= o2) acc += L.up[2][r - o2] * v[r - o2];
  else if (L.lo_c) acc += L.lo_c[r] * v[L.n + r];                       // plane of the rank below
  if (r + o2 < L.n) acc += L.up[2][r] * v[r + o2];
  else if (L.has_hi) acc += L.up[2][r] * v[L.n + o2 + (r - (L.n - o2))];  // plane of the rank above
  return acc;
}

// x = omega * rhs / diag   (first sweep from a zero guess)
__global__ void __launch_bounds__(kBlock, kMgCtasPerSm)
k_mg_smooth0(MgLevel L, const double *__restrict__ rhs, double *__restrict__ x, double omega,
             const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  for (int64_t r = (int64_t)blockIdx.x * kBlock + threadIdx.x; r < L.n; r += (int64_t)gridDim.x * kBlock) {
    const double d = L.diag[r];
    x[r] = d != 0.0 ? omega * rhs[r] / d : 0.0;
  }
}

// xout = xin + omega * (rhs - A xin) / diag
__global__ void __launch_bounds__(kBlock, kMgCtasPerSm)
k_mg_smooth(MgLevel L, const double *__restrict__ rhs, const double *__restrict__ xin, double *__restrict__ xout,
            double omega, const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  for (int64_t r = (int64_t)blockIdx.x * kBlock + threadIdx.x; r < L.n; r += (int64_t)gridDim.x * kBlock) {
    const double xr = xin[r];
    const double d = L.diag[r];
    const double res = rhs[r] - mg_apply(L, xin, r, xr);
    xout[r] = d != 0.0 ? xr + omega * res / d : xr;
  }
}

// rc[I] = sum over the children i of aggregate I of (rhs - A x)[i]     (residual + P^T in one pass)
__global__ void __launch_bounds__(kBlock, kMgCtasPerSm)
k_mg_restrict(MgLevel L, const double *__restrict__ rhs, const double *__restrict__ x, int cx, int cy, int cz,
              double *__restrict__ rc, const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const int64_t nc = (int64_t)cx * cy * cz;
  for (int64_t I = (int64_t)blockIdx.x * kBlock + threadIdx.x; I < nc; I += (int64_t)gridDim.x * kBlock) {
    const int Iz = (int)(I % cz), Iy = (int)((I / cz) % cy), Ix = (int)(I / ((int64_t)cz * cy));
    double s = 0.0;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) {
          const int ix = 2 * Ix + dx, iy = 2 * Iy + dy, iz = 2 * Iz + dz;
          if (ix < L.nx && iy < L.ny && iz < L.nz) {
            const int64_t r = ((int64_t)ix * L.ny + iy) * L.nz + iz;
            s += rhs[r] - mg_apply(L, x, r, x[r]);
          }
        }
    rc[I] = s;
  }
}

// x[i] += oc * ec[aggregate(i)]
__global__ void __launch_bounds__(kBlock, kMgCtasPerSm)
k_mg_prolong(MgLevel L, int cy, int cz, const double *__restrict__ ec, double *__restrict__ x, double oc,
             const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  for (int64_t r = (int64_t)blockIdx.x * kBlock + threadIdx.x; r < L.n; r += (int64_t)gridDim.x * kBlock) {
    const int iz = (int)(r % L.nz), iy = (int)((r / L.nz) % L.ny), ix = (int)(r / ((int64_t)L.nz * L.ny));
    const int64_t I = ((int64_t)(ix >> 1) * cy + (iy >> 1)) * cz + (iz >> 1);
    x[r] += oc * ec[I];
  }
}

// Galerkin coarse operator of one level (fine level Lf -> arrays of the next level)
__global__ void __launch_bounds__(kBlock)
k_mg_coarsen(MgLevel Lf, int cx, int cy, int cz, double *__restrict__ dc, double *__restrict__ u0,
             double *__restrict__ u1, double *__restrict__ u2, double *__restrict__ lo_c_coarse) {
  const int64_t nc = (int64_t)cx * cy * cz;
  for (int64_t I = (int64_t)blockIdx.x * kBlock + threadIdx.x; I < nc; I += (int64_t)gridDim.x * kBlock) {
    const int Iz = (int)(I % cz), Iy = (int)((I / cz) % cy), Ix = (int)(I / ((int64_t)cz * cy));
    double d = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int dx = 0; dx < 2; ++dx)
      for (int dy = 0; dy < 2; ++dy)
        for (int dz = 0; dz < 2; ++dz) {
          const int ix = 2 * Ix + dx, iy = 2 * Iy + dy, iz = 2 * Iz + dz;
          if (ix >= Lf.nx || iy >= Lf.ny || iz >= Lf.nz) continue;
          const int64_t r = ((int64_t)ix * Lf.ny + iy) * Lf.nz + iz;
          d += Lf.diag[r];
          // +z, +y, +x couplings of the child; couplings leaving the local box are not part of it
          const double cz_ = (iz + 1 < Lf.nz) ? Lf.up[0][r] : 0.0;
          const double cy_ = (iy + 1 < Lf.ny) ? Lf.up[1][r] : 0.0;
          // the +x coupling of the last plane leaves the box: kept only when a rank above exists
          const double cx_ = (ix + 1 < Lf.nx) ? Lf.up[2][r] : (Lf.has_hi ? Lf.up[2][r] : 0.0);
          if (dz == 0 && iz + 1 < Lf.nz) d += 2.0 * cz_; else a0 += cz_;
          if (dy == 0 && iy + 1 < Lf.ny) d += 2.0 * cy_; else a1 += cy_;
          if (dx == 0 && ix + 1 < Lf.nx) d += 2.0 * cx_; else a2 += cx_;
        }
    dc[I] = d; u0[I] = a0; u1[I] = a1; u2[I] = a2;
    if (lo_c_coarse && Ix == 0) {
      // coupling of the first coarse plane to the rank below = sum over its children in fine plane 0
      double lc = 0.0;
      for (int dy = 0; dy < 2; ++dy)
        for (int dz = 0; dz < 2; ++dz) {
          const int iy = 2 * Iy + dy, iz = 2 * Iz + dz;
          if (iy < Lf.ny && iz < Lf.nz) lc += Lf.lo_c[(int64_t)iy * Lf.nz + iz];
        }
      lo_c_coarse[(int64_t)Iy * cz + Iz] = lc;
    }
  }
}

// Coarsest level: kMgCoarseSweeps damped-Jacobi sweeps from zero in ONE CTA (n <= kMgCoarsest... any n, strided)
__global__ void __launch_bounds__(kBlock)
k_mg_coarse_solve(MgLevel L, const double *__restrict__ rhs, double *x, double *t, double omega, int sweeps,
                  const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  for (int64_t r = threadIdx.x; r < L.n; r += kBlock) {
    const double d = L.diag[r];
    x[r] = d != 0.0 ? omega * rhs[r] / d : 0.0;
  }
  __syncthreads();
  double *a = x, *b = t;
  for (int s = 1; s < sweeps; ++s) {
    for (int64_t r = threadIdx.x; r < L.n; r += kBlock) {
      const double xr = a[r], d = L.diag[r];
      b[r] = d != 0.0 ? xr + omega * (rhs[r] - mg_apply(L, a, r, xr)) / d : xr;
    }
    __syncthreads();
    double *tmp = a; a = b; b = tmp;
  }
  if (a != x) {
    for (int64_t r = threadIdx.x; r < L.n; r += kBlock) x[r] = a[r];
  }
}

// wrap-around check of the fine stencil: a +z coupling out of the last cell of a z-line (or +y out of
// the last line of a plane) would mean the matrix is not a box stencil.  flag[0] |= 1.
__global__ void k_mg_check_box(MgLevel L, int *flag) {
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < L.n; r += (int64_t)gridDim.x * blockDim.x) {
    const int iz = (int)(r % L.nz), iy = (int)((r / L.nz) % L.ny);
    if (iz == L.nz - 1 && r + 1 < L.n && L.up[0][r] != 0.0) bad = true;
    if (iy == L.ny - 1 && r + L.nz < L.n && L.up[1][r] != 0.0) bad = true;
  }
  if (bad) *flag = 1;
}

// ---- CG kernels for a general preconditioner (z = M^-1 r materialised) --------------------------------
// IterativeSolvers.cg recurrence, same as pcg.cuh: rho = z.r ; u = z + beta u ; c = A u ; alpha = rho/u.c ;
// x += alpha u (deferred) ; r -= alpha c ; residual = ||r||.
__device__ __forceinline__ void mgpcg_finish_rz(PcgScal *s, double rz) {
  if (s->done) return;
  s->rho_prev = s->iter == 0 ? 1.0 : s->rho;
  s->rho = rz;
}
__device__ __forceinline__ void mgpcg_finish_r(PcgScal *s, double rr, double *hist) {
  if (s->done) return;
  s->alpha_prev = s->rho / s->uc;
  const double resid = sqrt(rr);
  s->resid = resid;
  if (s->iter < s->hist_cap) hist[s->iter] = resid;
  s->iter += 1;
  if (resid <= s->reltol) { s->done = 1; s->converged = 1; }
  else if (s->iter >= s->maxiter) s->done = 1;
}

__global__ void __launch_bounds__(kBlock)
k_mgpcg_rz(int64_t n, const double *__restrict__ r, const double *__restrict__ z, double *partials,
           unsigned int *ticket, PcgScal *scal, int finalize_mode) {
  if (scal->done) return;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += r[i] * z[i];
  s = block_sum(s);
  double t;
  if (last_block_sum1(s, partials, ticket, &t)) {
    scal->red[0] = t;
    if (finalize_mode == 1) mgpcg_finish_rz(scal, t);
  }
}

// x += alpha_prev * u (deferred) ; u = z + beta u
__global__ void __launch_bounds__(kBlock)
k_mgpcg_update_u(int64_t n, const double *__restrict__ z, double *__restrict__ u, double *__restrict__ x,
                 const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const bool first = scal->iter == 0;
  const double beta = first ? 0.0 : scal->rho / scal->rho_prev;
  const double ap = scal->alpha_prev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (first) u[i] = z[i];
    else {
      const double ui = u[i];
      x[i] += ap * ui;
      u[i] = z[i] + beta * ui;
    }
  }
}

// r -= alpha c ; r.r ; closes the iteration
__global__ void __launch_bounds__(kBlock)
k_mgpcg_update_r(int64_t n, const double *__restrict__ c, double *__restrict__ r, double *partials,
                 unsigned int *ticket, PcgScal *scal, double *hist, int finalize_mode) {
  if (scal->done) return;
  const double alpha = scal->rho / scal->uc;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i] - alpha * c[i];
    r[i] = ri;
    s += ri * ri;
  }
  s = block_sum(s);
  double t;
  if (last_block_sum1(s, partials, ticket, &t)) {
    scal->red[0] = t;
    if (finalize_mode == 1) mgpcg_finish_r(scal, t, hist);
  }
}

// single-thread finishers for the NCCL path
__global__ void k_fin_rz(PcgScal *s) { mgpcg_finish_rz(s, s->red[0]); }
__global__ void k_fin_r(PcgScal *s, double *hist) { mgpcg_finish_r(s, s->red[0], hist); }

}  // namespace fvb
