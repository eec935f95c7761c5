// amg.cuh -- aggregation algebraic multigrid V-cycle on CSR rows: the preconditioner class the reference uses
// (AlgebraicMultigrid.ruge_stuben + aspreconditioner, src/FiniteVolume.jl:160) for matrices that are NOT
// box-structured -- fracture networks (examples/fractures/ex.jl:14), Dirichlet sets that leave the grid's
// diagonals (test/theis.jl) -- where mg.cuh's geometric aggregation does not apply.
//
// Set-up, all on the device and deterministic (no floating-point atomics):
//   * pairwise aggregation by handshaking: every unmatched row proposes to its strongest unmatched neighbour
//     (most negative a_ij, ties by a symmetric edge hash); mutual proposals become pairs; six rounds; rows left
//     over join the pair of their strongest paired neighbour (hubs), else stay singletons.  Applied twice per
//     level ("double pairwise") with the intermediate Galerkin matrix built and dropped.
//   * Galerkin coarse operator for piecewise-constant prolongation, A_c[I,J] = sum_{i in I, j in J} a_ij: one
//     thread per coarse row merges its member rows (members ascending, entries in row order => fixed summation
//     order), columns sorted ascending; count pass, scan, fill pass.
// Cycle: V(nu,nu) with damped Jacobi from a zero initial guess, coarse correction scaled by `oc`, coarsest level by
// a fixed number of sweeps in one CTA -- a fixed symmetric positive operator, hence legal inside CG.
#pragma once
#include "common.cuh"

namespace fvb {

constexpr int kAmgMaxLevels = 24;
constexpr int kAmgCoarsest = 512;     // stop coarsening at or below this many rows
constexpr int kAmgCoarseSweeps = 40;  // even
constexpr int kAmgMaxRow = 128;       // entries a coarse row may have (more: the hierarchy stops at that level)
constexpr int kAmgRounds = 6;         // handshake rounds per pairwise pass
constexpr double kAmgStall = 0.9;     // a level that keeps more than this fraction of its rows ends the hierarchy
constexpr int kAmgOneCta = 2048;      // coarsest levels up to this size are swept by a single CTA

struct AmgLevel {
  int n, nnz;
  const int *rowptr, *colidx;
  const double *vals;
  double *dinv;       // 1 / a_ii
  int *agg;           // row -> coarse row of the next level (null on the coarsest)
  int *memptr, *mem;  // coarse row -> its rows on this level, ascending
  int nc;
  double *x, *r, *t;  // iterate, right-hand side, scratch (level 0: r is the caller's residual)
};

__global__ void k_amg_dinv(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                           const double *__restrict__ vals, double *__restrict__ dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
    if (colidx[k] == i) d = vals[k];
  dinv[i] = d != 0.0 ? 1.0 / d : 0.0;
}

// ---- pairwise aggregation ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned amg_edge_hash(int a, int b) {
  unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
  unsigned h = lo * 0x9E3779B1u ^ (hi + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__global__ void k_amg_propose(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                              const double *__restrict__ vals, const int *__restrict__ match, int *__restrict__ prop) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int best = -1;
  double bw = 0.0;
  unsigned bh = 0u;
  if (match[i] < 0) {
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int j = colidx[k];
      if (j == i || j >= n || match[j] >= 0) continue;  // (j >= n: halo column of another rank -- never aggregated)
      const double w = -vals[k];
      if (w <= 0.0) continue;
      // strict total order on the EDGES (weight, then a symmetric hash of the endpoints): proposals then point along
      // locally dominant edges, which matches a constant fraction of the rows per round even when all weights are equal
      const unsigned hk = amg_edge_hash(i, j);
      if (best < 0 || w > bw || (w == bw && hk > bh)) { bw = w; bh = hk; best = j; }
    }
  }
  prop[i] = best;
}
__global__ void k_amg_accept(int n, const int *__restrict__ prop, int *__restrict__ match) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || match[i] >= 0) return;
  const int j = prop[i];
  if (j >= 0 && prop[j] == i) match[i] = j;
}
// Rows the handshakes left alone join the pair of their strongest paired neighbour (same edge order), so that hubs
// -- a fracture intersection whose many neighbours all prefer it -- do not stall the coarsening: attach[i] = that
// neighbour, or -1 (the row stays a singleton aggregate).
__global__ void k_amg_attach(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                             const double *__restrict__ vals, const int *__restrict__ match, int *__restrict__ attach) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int best = -1;
  double bw = 0.0;
  unsigned bh = 0u;
  if (match[i] < 0) {
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int j = colidx[k];
      if (j == i || j >= n || match[j] < 0) continue;
      const double w = -vals[k];
      if (w <= 0.0) continue;
      const unsigned hk = amg_edge_hash(i, j);
      if (best < 0 || w > bw || (w == bw && hk > bh)) { bw = w; bh = hk; best = j; }
    }
  }
  attach[i] = best;
}
__device__ __forceinline__ int amg_root(int i, const int *__restrict__ match, const int *__restrict__ attach) {
  const int m = match[i];
  if (m >= 0) return min(i, m);
  const int a = attach[i];
  return a >= 0 ? min(a, match[a]) : i;
}
__global__ void k_amg_roots(int n, const int *__restrict__ match, const int *__restrict__ attach, int *__restrict__ isroot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) isroot[i] = amg_root(i, match, attach) == i;
}
__global__ void k_amg_number(int n, const int *__restrict__ match, const int *__restrict__ attach,
                             const int *__restrict__ cid, int *__restrict__ agg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) agg[i] = cid[amg_root(i, match, attach)];
}
__global__ void k_amg_compose(int n, const int *__restrict__ a1, const int *__restrict__ a2, int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a2[a1[i]];
}

// ---- member lists (coarse row -> fine rows, ascending) ------------------------------------------------------------
__global__ void k_amg_count_members(int n, const int *__restrict__ agg, int *__restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[agg[i]], 1);
}
__global__ void k_amg_fill_members(int n, const int *__restrict__ agg, const int *__restrict__ memptr,
                                   int *__restrict__ cursor, int *__restrict__ mem) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mem[memptr[agg[i]] + atomicAdd(&cursor[agg[i]], 1)] = i;
}
__global__ void k_amg_sort_members(int nc, const int *__restrict__ memptr, int *__restrict__ mem) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= nc) return;
  const int lo = memptr[I], hi = memptr[I + 1];
  for (int a = lo + 1; a < hi; ++a) {
    const int v = mem[a];
    int p = a;
    while (p > lo && mem[p - 1] > v) { mem[p] = mem[p - 1]; --p; }
    mem[p] = v;
  }
}

// ---- Galerkin product for piecewise-constant P ----------------------------------------------------------------------
// WRITE = false: rowcnt[I] = number of distinct coarse columns; WRITE = true: columns ascending + summed values.
template <bool WRITE>
__global__ void __launch_bounds__(128)
k_amg_galerkin(int nc, const int *__restrict__ memptr, const int *__restrict__ mem, const int *__restrict__ rowptr,
               const int *__restrict__ colidx, const double *__restrict__ vals, const int *__restrict__ agg, int nfine,
               const int *__restrict__ crowptr, int *__restrict__ ccol, double *__restrict__ cval,
               int *__restrict__ rowcnt, int *__restrict__ overflow) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= nc) return;
  int cols[kAmgMaxRow];
  double vs[kAmgMaxRow];
  int m = 0;
  bool over = false;
  for (int q = memptr[I]; q < memptr[I + 1]; ++q) {
    const int i = mem[q];
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int j = colidx[k];
      if (j >= nfine) continue;  // couplings to another rank's rows stay out of the (block-local) hierarchy
      const int c = agg[j];
      int p = 0;
      while (p < m && cols[p] != c) ++p;
      if (p < m) { if (WRITE) vs[p] = __dadd_rn(vs[p], vals[k]); }
      else if (m < kAmgMaxRow) { cols[m] = c; if (WRITE) vs[m] = vals[k]; ++m; }
      else over = true;
    }
  }
  if (over) *overflow = 1;
  if constexpr (!WRITE) {
    rowcnt[I] = m;
  } else {
    // insertion sort by column (rows are short)
    for (int a = 1; a < m; ++a) {
      const int c = cols[a];
      const double v = vs[a];
      int p = a;
      while (p > 0 && cols[p - 1] > c) { cols[p] = cols[p - 1]; vs[p] = vs[p - 1]; --p; }
      cols[p] = c; vs[p] = v;
    }
    int w = crowptr[I];
    for (int a = 0; a < m; ++a) { ccol[w] = cols[a]; cval[w++] = vs[a]; }
  }
}

// ---- cycle kernels (one thread per row; coarse levels are small, the fine level is read ~2*nu+1 times per cycle) ----
__device__ __forceinline__ double amg_row(const AmgLevel &L, const double *__restrict__ x, int i) {
  double acc = 0.0;
  for (int k = L.rowptr[i]; k < L.rowptr[i + 1]; ++k) {
    const int j = L.colidx[k];
    if (j < L.n) acc += L.vals[k] * x[j];
  }
  return acc;
}
__global__ void __launch_bounds__(kBlock)
k_amg_smooth0(AmgLevel L, const double *__restrict__ r, double *__restrict__ x, double omega, const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L.n) x[i] = omega * L.dinv[i] * r[i];
}
__global__ void __launch_bounds__(kBlock)
k_amg_smooth(AmgLevel L, const double *__restrict__ r, const double *__restrict__ xin, double *__restrict__ xout, double omega,
             const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L.n) xout[i] = xin[i] + omega * L.dinv[i] * (r[i] - amg_row(L, xin, i));
}
// coarse right-hand side: r_c[I] = sum over the members of (r - A x), members ascending
__global__ void __launch_bounds__(kBlock)
k_amg_restrict(AmgLevel L, const double *__restrict__ r, const double *__restrict__ x, double *__restrict__ rc,
               const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= L.nc) return;
  double s = 0.0;
  for (int q = L.memptr[I]; q < L.memptr[I + 1]; ++q) {
    const int i = L.mem[q];
    s += r[i] - amg_row(L, x, i);
  }
  rc[I] = s;
}
__global__ void __launch_bounds__(kBlock)
k_amg_prolong(AmgLevel L, const double *__restrict__ xc, double *__restrict__ x, double oc, const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L.n) x[i] += oc * xc[L.agg[i]];
}
__global__ void __launch_bounds__(kBlock)
k_amg_coarse_solve(AmgLevel L, const double *__restrict__ rhs, double *x, double *t, double omega, int sweeps,
                   const PcgScal *__restrict__ scal) {
  if (scal->done) return;
  for (int i = threadIdx.x; i < L.n; i += kBlock) x[i] = omega * L.dinv[i] * rhs[i];
  __syncthreads();
  double *a = x, *b = t;
  for (int s = 1; s < sweeps; ++s) {
    for (int i = threadIdx.x; i < L.n; i += kBlock) b[i] = a[i] + omega * L.dinv[i] * (rhs[i] - amg_row(L, a, i));
    __syncthreads();
    double *tmp = a; a = b; b = tmp;
  }
  if (a != x)
    for (int i = threadIdx.x; i < L.n; i += kBlock) x[i] = a[i];
}

}  // namespace fvb
