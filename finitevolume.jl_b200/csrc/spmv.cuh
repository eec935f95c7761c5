// spmv.cuh -- y = A x (+ sigma * D .* x) on the CSR rows this rank owns.
// Stands in for mul!(y, A::SparseMatrixCSC, x) inside IterativeSolvers.cg (called from
// src/FiniteVolume.jl:161 and src/transient.jl:52,55) and for `b - A u` (src/transient.jl:197).
//
// Finite-volume rows are short (7 entries on a regular grid, 3-14 on fracture meshes), so
// "one warp per row" would idle 25 of 32 lanes.  Instead each CTA owns kRows consecutive
// rows and streams their entries -- a contiguous slice of vals/colidx -- with fully
// coalesced loads, multiplies by the gathered x (L1/L2 hits on a structured grid: the
// columns of neighbouring rows overlap), parks the products in shared memory, and then
// each thread folds the products of its own row in CSR order.  HBM traffic per row is the
// algorithmic 12*nnz_row + 4 + 8 + 8 bytes; the sum order is the row's column order, so
// the result is run-to-run deterministic (and equals a serial CSR gather without FMA).
//
// The optional epilogue fuses the CG quantities that would otherwise cost another pass:
// the partial dot product x.y (u.Au of the recurrence) reduced per block.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace fvb {

constexpr int kSpmvRows = 256;            // rows per CTA == threads per CTA
constexpr int kSpmvTile = 256 * 9;        // products parked per pass (18 KB of smem)

template <bool DOT>
__global__ void __launch_bounds__(kSpmvRows)
k_spmv(int nrows, const int *__restrict__ rowptr, const int *__restrict__ colidx,
       const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
       const double *__restrict__ Dvec, double sigma, double *__restrict__ partials,
       unsigned int *ticket, PcgScal *scal, int finalize_mode) {
  __shared__ double prod[kSpmvTile];
  __shared__ int rp[kSpmvRows + 1];
  if (DOT && scal->done) return;
  const int t = threadIdx.x;
  const int r0 = blockIdx.x * kSpmvRows;
  const int r = r0 + t;
  rp[t] = rowptr[min(r, nrows)];
  if (t == 0) rp[kSpmvRows] = rowptr[min(r0 + kSpmvRows, nrows)];
  __syncthreads();
  const int nz0 = rp[0], nz1 = rp[kSpmvRows];
  const int my_lo = rp[t], my_hi = rp[t + 1];
  double acc = 0.0;
  for (int base = nz0; base < nz1; base += kSpmvTile) {
    const int end = min(base + kSpmvTile, nz1);
#pragma unroll 3
    for (int k = base + t; k < end; k += kSpmvRows)
      prod[k - base] = __dmul_rn(vals[k], __ldg(&x[colidx[k]]));
    __syncthreads();
    const int lo = max(my_lo, base), hi = min(my_hi, end);
    for (int k = lo; k < hi; ++k) acc = __dadd_rn(acc, prod[k - base]);
    if (end < nz1) __syncthreads();
  }
  double contrib = 0.0;
  if (r < nrows) {
    const double xr = x[r];
    if (sigma != 0.0) acc += sigma * (Dvec ? Dvec[r] : 1.0) * xr;
    y[r] = acc;
    contrib = xr * acc;
  }
  if (DOT) {
    // u.Au: block partial, last block folds the partials in block order
    double s = block_sum(contrib);
    if (last_block_sum1(s, partials, ticket, &s)) {
      scal->red[0] = s;
      if (finalize_mode == 1) scal->uc = s;
    }
  }
}

// Gather the rows a peer needs into the contiguous send buffer (halo pack).
__global__ void k_pack(int64_t n, const int *__restrict__ rows, const double *__restrict__ x,
                       double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[rows[i]];
}

}  // namespace fvb
