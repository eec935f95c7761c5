// coop.cuh -- one cooperative kernel = one whole step-doubling attempt of the transient integrator on a SMALL system.
//
// The reference's adaptive stepper (src/transient.jl:78-87, backwardeulertwostep!) takes one full backward-Euler step
// and two half steps and compares them: three warm-started linear solves of (A + D/dt) u+ = b + D u/dt
// (src/transient.jl:65-76 in its SPD form) plus one norm per attempt.  On the Theis problem (test/theis.jl:
// 15 650 unknowns, ~3300 solves of 5-20 CG iterations) every kernel of the launch-per-phase PCG runs for a couple of
// microseconds, so the solve is pure launch latency and host round trips.  Here the grid is resident (one CTA per
// SM at most, cooperative launch), the three phases of a CG iteration are separated by grid-wide barriers instead
// of kernel boundaries, all scalars are re-derived redundantly by every CTA from per-CTA partials (fixed order, so
// every CTA holds bit-identical values and takes the same branch), and the whole attempt -- up to three solves and
// ||onestep - twostep||_2 -- is ONE launch; the host reads back one double per attempt.
//
// Recurrence and stopping rule are those of pcg.cuh (IterativeSolvers.cg! 0.8.1 with Pl = Jacobi: tolerance relative
// to the initial residual of the warm start).  Both matrix formats are served: CSR rows or the symmetric-diagonal copy.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "dia.cuh"
#include "reduce.cuh"

namespace fvb {

namespace cg = cooperative_groups;

constexpr int kCoopMaxRows = 1 << 20;  // larger systems are bandwidth-bound: the launch-per-phase kernels serve them

struct CoopSolve {
  const double *b;   // unscaled right-hand side (forward) / adjoint forcing g
  const double *u;   // state the step starts from
  double *out;       // result
  double sigma;      // 1/dt
};

struct CoopJob {
  int nsolves;               // 1..3
  CoopSolve s[3];
  const double *na, *nb;     // ||na - nb||_2 -> result[0] after the solves (null: skipped)
  int adjoint;               // 0: (A + sD) u+ = b + s D u;  1: (A + sD) w = g + s u, u+ = D w, warm start w0 = u ./ D
  double rtol;
  long long maxiter;
  // matrix
  int n;
  const int *rowptr, *colidx;
  const double *vals;        // CSR (null when the diagonal copy is used)
  DiaDesc D;                 // diagonal copy (D.K = 0 when CSR is used)
  const double *diag, *Dvec; // diag(A); D = Ss*vol (null: identity)
  // work vectors (n doubles each) and reduction scratch (3 * gridDim doubles)
  double *x, *r, *p, *c, *dinv, *rhs;
  double *partials;
  // out: result[0] = ||na - nb||, result[1] = total CG iterations, result[2] = 1 if every solve converged
  double *result;
};

template <int K>
__device__ __forceinline__ double coop_row_dia(const DiaDesc &D, const double *__restrict__ x, int r) {
  double acc = 0.0;
#pragma unroll
  for (int k = K - 1; k >= 0; --k) {
    const double lo = D.U[k][r];
    if (lo != 0.0) {
      const int64_t il = r - D.off[k];
      acc = __dadd_rn(acc, __dmul_rn(lo, il >= 0 ? x[il] : 0.0));
    }
  }
  acc = __dadd_rn(acc, __dmul_rn(D.diag[r], x[r]));
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double up = D.U[k][D.off[k] + r];
    if (up != 0.0) {
      const int64_t iu = r + D.off[k];
      acc = __dadd_rn(acc, __dmul_rn(up, iu < D.nf ? x[iu] : 0.0));
    }
  }
  return acc;
}

__device__ __forceinline__ double coop_row(const CoopJob &J, const double *x, int r) {
  // x is written by other CTAs between grid barriers: plain (coherent) loads, no read-only cache path
  if (J.vals) {
    double acc = 0.0;
    for (int k = J.rowptr[r]; k < J.rowptr[r + 1]; ++k) acc = __dadd_rn(acc, __dmul_rn(J.vals[k], x[J.colidx[k]]));
    return acc;
  }
  switch (J.D.K) {
    case 1: return coop_row_dia<1>(J.D, x, r);
    case 2: return coop_row_dia<2>(J.D, x, r);
    case 3: return coop_row_dia<3>(J.D, x, r);
    default: return coop_row_dia<4>(J.D, x, r);
  }
}

// every CTA folds the per-CTA partials of slot `which` in CTA order: identical result everywhere
__device__ __forceinline__ double coop_total(const double *partials, int which) {
  __shared__ double bc;
  double s = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(&partials[which * gridDim.x + i]);
  s = block_sum(s);
  if (threadIdx.x == 0) bc = s;
  __syncthreads();
  const double t = bc;
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kBlock)
k_coop_attempt(CoopJob J) {
  cg::grid_group grid = cg::this_grid();
  const int n = J.n;
  const int tid0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  long long total_its = 0;
  int all_conv = 1;
  for (int q = 0; q < J.nsolves; ++q) {
    const CoopSolve S = J.s[q];
    const double sg = S.sigma;
    // ---- set-up: rhs, Jacobi, warm start -------------------------------------------------------------------------
    for (int i = tid0; i < n; i += stride) {
      const double Di = J.Dvec ? J.Dvec[i] : 1.0;
      const double ui = S.u[i];
      J.rhs[i] = J.adjoint ? S.b[i] + sg * ui : S.b[i] + sg * (Di * ui);
      J.x[i] = J.adjoint ? ui / Di : ui;
      J.dinv[i] = 1.0 / (J.diag[i] + sg * Di);
    }
    grid.sync();
    double s0 = 0.0, s1 = 0.0;
    for (int i = tid0; i < n; i += stride) {
      const double Di = J.Dvec ? J.Dvec[i] : 1.0;
      const double xi = J.x[i];
      const double ri = J.rhs[i] - (coop_row(J, J.x, i) + sg * Di * xi);
      J.r[i] = ri;
      s0 += J.dinv[i] * ri * ri;
      s1 += ri * ri;
    }
    s0 = block_sum(s0);
    s1 = block_sum(s1);
    if (threadIdx.x == 0) { J.partials[blockIdx.x] = s0; J.partials[gridDim.x + blockIdx.x] = s1; }
    grid.sync();
    double rho = coop_total(J.partials, 0);
    double resid = sqrt(coop_total(J.partials, 1));
    const double reltol = resid * J.rtol;
    double rho_prev = 1.0;
    long long it = 0;
    // ---- iterations -----------------------------------------------------------------------------------------------
    while (it < J.maxiter && resid > reltol) {
      const double beta = it == 0 ? 0.0 : rho / rho_prev;
      for (int i = tid0; i < n; i += stride) J.p[i] = J.dinv[i] * J.r[i] + (it == 0 ? 0.0 : beta * J.p[i]);
      grid.sync();
      double d = 0.0;
      for (int i = tid0; i < n; i += stride) {
        const double pi = J.p[i];
        const double ci = coop_row(J, J.p, i) + sg * (J.Dvec ? J.Dvec[i] : 1.0) * pi;
        J.c[i] = ci;
        d += pi * ci;
      }
      d = block_sum(d);
      if (threadIdx.x == 0) J.partials[2 * gridDim.x + blockIdx.x] = d;
      grid.sync();
      const double alpha = rho / coop_total(J.partials, 2);
      s0 = 0.0; s1 = 0.0;
      for (int i = tid0; i < n; i += stride) {
        J.x[i] += alpha * J.p[i];
        const double ri = J.r[i] - alpha * J.c[i];
        J.r[i] = ri;
        s0 += J.dinv[i] * ri * ri;
        s1 += ri * ri;
      }
      s0 = block_sum(s0);
      s1 = block_sum(s1);
      if (threadIdx.x == 0) { J.partials[blockIdx.x] = s0; J.partials[gridDim.x + blockIdx.x] = s1; }
      grid.sync();
      rho_prev = rho;
      rho = coop_total(J.partials, 0);
      resid = sqrt(coop_total(J.partials, 1));
      ++it;
    }
    total_its += it;
    if (!(resid <= reltol)) all_conv = 0;
    for (int i = tid0; i < n; i += stride) S.out[i] = J.adjoint ? (J.Dvec ? J.Dvec[i] : 1.0) * J.x[i] : J.x[i];
    grid.sync();  // out may be the next solve's starting state; work vectors are reused
  }
  double nrm = 0.0;
  if (J.na && J.nb) {
    double s = 0.0;
    for (int i = tid0; i < n; i += stride) {
      const double d = J.na[i] - J.nb[i];
      s += d * d;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) J.partials[2 * gridDim.x + blockIdx.x] = s;
    grid.sync();
    nrm = sqrt(coop_total(J.partials, 2));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    J.result[0] = nrm;
    J.result[1] = (double)total_its;
    J.result[2] = (double)all_conv;
  }
}

}  // namespace fvb
