// spmv.cuh -- y = A x (+ sigma * D .* x) on the CSR rows this rank owns.
// Stands in for mul!(y, A::SparseMatrixCSC, x) inside IterativeSolvers.cg (called from
// src/FiniteVolume.jl:161 and src/transient.jl:52,55) and for `b - A u` (src/transient.jl:197).
//
// Finite-volume rows are short (7 entries on a regular grid, 3-14 on fracture meshes), so
// "one warp per row" would idle 25 of 32 lanes.  Instead the kernel is a persistent,
// warp-specialised stream over tiles of kSpmvRows consecutive rows:
//   * one producer warp per CTA asks the TMA unit (cp.async.bulk -> UBLKCP) for the tile's
//     slice of rowptr, vals and colidx -- three contiguous 1-D copies that complete on an
//     mbarrier, two tiles in flight per CTA -- so the matrix stream costs no LSU/L1 work;
//   * eight consumer warps own one row per thread and walk it out of shared memory.  Since
//     thread t owns row r0+t, the gathers x[col] of a warp hit 32 consecutive addresses per
//     diagonal on a structured grid (two 128-byte lines, L1/L2 hits), and each row is summed
//     in its column order: run-to-run deterministic, equal to a serial CSR gather without FMA.
// HBM traffic per row is the algorithmic 12*nnz_row + 4 + 8 + 8 bytes.
// The DOT epilogue fuses the u.Au of the CG recurrence: one partial per CTA (grid is only
// 4 CTAs per SM, so the "last CTA folds the partials" ticket costs ~600 atomics, not 5e5).
#pragma once
#include "common.cuh"
#include "peer_base.cuh"
#include "reduce.cuh"
#include "tma.cuh"

namespace fvb {

constexpr int kSpmvRows = 256;            // rows per tile == consumer threads per CTA
constexpr int kSpmvThreads = kSpmvRows + 32;  // + one producer warp
constexpr int kSpmvTile = 2048;           // entries staged per tile (rows with more spill to global)
constexpr int kSpmvStages = 2;
constexpr int kSpmvCtasPerSm = 4;
constexpr int kCsrPad = 16;               // vals/colidx are allocated with this many spare entries
constexpr int kRowptrPad = kSpmvRows + 8; // rowptr is allocated with this many spare entries (= nnz)

struct SpmvSmem {
  alignas(128) double sv[kSpmvStages][kSpmvTile + 4];
  alignas(128) int sc[kSpmvStages][kSpmvTile + 4];
  alignas(128) int rp[kSpmvStages][kSpmvRows + 8];
  alignas(8) uint64_t full[kSpmvStages];
  alignas(8) uint64_t empty[kSpmvStages];
  double wsum[kSpmvRows / 32];
  int is_last;
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSpmvRows) : "memory"); }

template <bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, kSpmvCtasPerSm)
k_spmv(int nrows, const int *__restrict__ rowptr, const int *__restrict__ colidx,
       const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
       const double *__restrict__ Dvec, double sigma, double *__restrict__ partials,
       unsigned int *ticket, PcgScal *scal, int finalize_mode, PeerRed pr) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SpmvSmem &S = *reinterpret_cast<SpmvSmem *>(smem_raw);
  if (DOT && scal->done) return;
  const int t = threadIdx.x;
  const int ntiles = (nrows + kSpmvRows - 1) / kSpmvRows;
  if (t == 0) {
    for (int s = 0; s < kSpmvStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], kSpmvRows / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (t >= kSpmvRows) {
    // ---------------- producer warp: one elected lane drives the TMA unit ----------------
    if (t == kSpmvRows) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int stage = it % kSpmvStages;
        // slot reuse: wait until the consumers released its previous tile (use number it/stages - 1)
        if (it >= kSpmvStages) mbar_wait(&S.empty[stage], (uint32_t)((it / kSpmvStages - 1) & 1));
        const int r0 = tile * kSpmvRows;
        const int nz0 = __ldg(&rowptr[r0]);
        const int nz1 = __ldg(&rowptr[min(r0 + kSpmvRows, nrows)]);
        const int a0 = nz0 & ~3;
        const int cnt = (min(nz1 - a0, kSpmvTile) + 3) & ~3;
        mbar_expect_tx(&S.full[stage], (uint32_t)cnt * 12u + (kSpmvRows + 4) * 4u);
        bulk_g2s(S.rp[stage], rowptr + r0, (kSpmvRows + 4) * 4u, &S.full[stage]);
        if (cnt > 0) {
          bulk_g2s(S.sv[stage], vals + a0, (uint32_t)cnt * 8u, &S.full[stage]);
          bulk_g2s(S.sc[stage], colidx + a0, (uint32_t)cnt * 4u, &S.full[stage]);
        }
      }
    }
    return;
  }

  // ---------------- consumers: thread t owns row r0 + t of every tile ----------------
  double dot = 0.0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it % kSpmvStages;
    mbar_wait(&S.full[stage], (uint32_t)((it / kSpmvStages) & 1));
    const int r = tile * kSpmvRows + t;
    const int a0 = S.rp[stage][0] & ~3;
    const int my_lo = S.rp[stage][t], my_hi = S.rp[stage][t + 1];
    const int lim = a0 + kSpmvTile;                 // entries at or beyond lim were not staged
    const double *sv = S.sv[stage] - a0;
    const int *sc = S.sc[stage] - a0;
    double acc = 0.0;
    const int hi_s = min(my_hi, lim);
    for (int k = my_lo; k < hi_s; k += 8) {
      double p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k + j < hi_s) p[j] = __dmul_rn(sv[k + j], __ldg(&x[sc[k + j]]));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k + j < hi_s) acc = __dadd_rn(acc, p[j]);
    }
    for (int k = max(my_lo, lim); k < my_hi; ++k)   // oversize tiles: tail straight from global
      acc = __dadd_rn(acc, __dmul_rn(__ldg(&vals[k]), __ldg(&x[__ldg(&colidx[k])])));
    __syncwarp();
    if ((t & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.empty[stage])) : "memory");
    if (r < nrows) {
      const double xr = x[r];
      if (sigma != 0.0) acc += sigma * (Dvec ? Dvec[r] : 1.0) * xr;
      y[r] = acc;
      dot += xr * acc;
    }
  }
  if (DOT) {
    // u.Au: one partial per CTA; the CTA drawing the last ticket folds them in CTA order
    double s = warp_sum(dot);
    if ((t & 31) == 0) S.wsum[t >> 5] = s;
    consumer_sync();
    if (t == 0) {
      double tot = 0.0;
      for (int w = 0; w < kSpmvRows / 32; ++w) tot += S.wsum[w];
      partials[blockIdx.x] = tot;
      __threadfence();
      S.is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    consumer_sync();
    if (S.is_last) {
      __threadfence();
      double q = 0.0;
      for (unsigned int i = t; i < gridDim.x; i += kSpmvRows) q += __ldcg(&partials[i]);
      q = warp_sum(q);
      consumer_sync();
      if ((t & 31) == 0) S.wsum[t >> 5] = q;
      consumer_sync();
      if (t < 32) {
        double tot = 0.0;
        if (t == 0)
          for (int w = 0; w < kSpmvRows / 32; ++w) tot += S.wsum[w];
        const bool ok = finalize_mode != 2 || peer_allreduce_warp(pr, &tot, 1, scal);
        if (t == 0) {
          scal->red[0] = tot;
          if (ok && finalize_mode >= 1) scal->uc = tot;
        }
      }
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
