// dia_tma.cuh -- the symmetric-diagonal SpMV of dia.cuh as a persistent, warp-specialised
// TMA pipeline (same structure as the CSR kernel in spmv.cuh).
//
// On the diagonal format every operand of a tile of T consecutive rows is a CONTIGUOUS slice:
//     U_k[o_k + r0 .. +T)  upper entries          U_k[r0 .. +T)   lower entries
//     x[r0 +- o_k .. +T)   neighbour values       x[r0 .. +T), diag[r0 .. +T)
// so one elected producer lane can ask the TMA unit (cp.async.bulk -> UBLKCP) for all of them and
// the eight consumer warps only ever touch shared memory: no LSU/L1 wavefronts, no address
// arithmetic and no registers are spent on the streams, and two or three tiles per SM are in flight
// regardless of occupancy.  The plane-distance re-reads (lower diagonals, x[r +- n2*n3]) are served
// by L2 exactly as in k_spmv_dia: the grid is resident (1 CTA/SM) and tiles are dealt round-robin,
// so the sweep is a moving front thinner than one plane.
//
// What bounds this product on B200 is not HBM but the L2 slices (ncu: ~11.7 TB/s of L2->SM traffic,
// the LTS throughput cap, at 6 TB/s of DRAM traffic): every re-read of a neighbour's value or of a
// lower diagonal is an L2 transaction.  Hence the tile is long (T = 1024 rows) and every offset
// o_k <= T/2 is "near": its neighbours x[r +- o_k] and its lower entries come out of the tile's own
// centre slices, fetched once with a margin of o_k (z- AND y-neighbours on grids up to 512 points
// per axis: 76 instead of 88 bytes of L2 traffic per row).  Larger offsets get their own slices,
// started one element early when o_k is odd so that every copy stays 16-byte aligned.
//
// Tiles whose slices would leave the owned range [0, nf) -- the first and last plane of a slab,
// where columns live in the halo or do not exist -- are "edge" tiles: the producer completes the
// barrier without copying and the consumers take the per-row path of dia.cuh.  The host only
// picks this kernel when edge tiles are a small minority (fvb200.cu: launch_spmv).
//
// Row sums use the same order and the same separate multiply/add as k_spmv_dia and k_spmv, so
// y is bit-identical to both.
#pragma once
#include "dia.cuh"
#include "tma.cuh"

namespace fvb {

constexpr int kDiaTmaRows = 512;                        // consumer threads per CTA
constexpr int kDiaTmaR = 2;                             // rows per consumer thread and tile
constexpr int kDiaTmaTile = kDiaTmaRows * kDiaTmaR;     // T
constexpr int kDiaTmaThreads = kDiaTmaRows + 32;        // + producer warp
constexpr int kDiaTmaMaxStages = 3;
constexpr int kDiaTmaCtasPerSm = 1;
constexpr int kDiaNearMax = kDiaTmaTile / 2;            // offsets up to here are served by the centre slices
constexpr int kDiaTmaSmemMax = 227 * 1024;              // dynamic shared memory one CTA may opt in to
constexpr int kDiaBlockRows = 262144;                   // rows of one plane swept together when planes are larger

// Per-stage layout in doubles (filled on the host: dia_tma_layout).
struct DiaTmaLayout {
  int margin;        // M = largest near offset rounded up to even (>= 2)
  int xc;            // x[r0 - M .. r0 + T + M)
  int dg;            // diag[r0 .. r0 + T)                                  (unused when UNIT)
  int un[kDiaMaxOff];  // near k: U_k[r0 .. r0 + T + even(o_k))
  int xl[kDiaMaxOff];  // far k:  x[(r0 - o_k) & ~1 .. + T + 2)
  int xu[kDiaMaxOff];  // far k:  x[(r0 + o_k) & ~1 .. + T + 2)
  int ul[kDiaMaxOff];  // far k:  U_k[r0 .. r0 + T)
  int uu[kDiaMaxOff];  // far k:  U_k[(o_k + r0) & ~1 .. + T + 2)
  int stage_doubles; // size of one stage
  int stages;        // 3 if they fit into shared memory, else 2 (0: does not fit at all)
  int bar_off;       // byte offset of the mbarriers / reduction scratch behind the stages
  uint32_t tx_bytes; // bytes one interior tile brings in
  int64_t reach;     // interior tiles satisfy r0 >= reach and r0 + T + reach <= nf
  int blk_tpp;       // plane-blocked tile order (DiaTileOrder): tiles per plane-distance offset, 0 = off
  int blk_yb;        // tiles of one plane that form a block (a divisor of blk_tpp)
};

inline int dia_even_up(int64_t o) { return (int)((o + 1) & ~(int64_t)1); }
inline size_t dia_tma_tail_bytes() {
  return 2 * kDiaTmaMaxStages * sizeof(uint64_t) + (kDiaTmaRows / 32) * sizeof(double) + 16;
}

inline DiaTmaLayout dia_tma_layout(int K, const int64_t *off, bool unit) {
  DiaTmaLayout L = {};
  const int T = kDiaTmaTile;
  int M = 2;
  for (int k = 0; k < K; ++k)
    if (off[k] <= kDiaNearMax) M = std::max(M, dia_even_up(off[k]));
  L.margin = M;
  int p = 0;
  auto take = [&](int len) { int at = p; p += len; return at; };
  L.xc = take(T + 2 * M);
  if (!unit) L.dg = take(T);
  int64_t reach = M;
  for (int k = 0; k < K; ++k) {
    if (off[k] <= kDiaNearMax) {
      L.un[k] = take(T + dia_even_up(off[k]));
    } else {
      L.xl[k] = take(T + 2);
      L.xu[k] = take(T + 2);
      L.ul[k] = take(T);
      L.uu[k] = take(T + 2);
      reach = std::max<int64_t>(reach, off[k] + 2);
    }
  }
  L.stage_doubles = p;
  L.tx_bytes = (uint32_t)p * 8u;
  L.stages = 0;
  for (int st = kDiaTmaMaxStages; st >= 2 && !L.stages; --st)
    if ((size_t)st * p * 8 + dia_tma_tail_bytes() <= (size_t)kDiaTmaSmemMax) L.stages = st;
  L.bar_off = std::max(L.stages, 1) * p * 8;
  L.reach = reach;
  // Plane-blocked sweep for big planes.  The plane-distance re-reads (x[r +- o], the lower plane diagonal) are L2
  // hits only while two planes' worth of streams (2 * o rows * ~40 B) fit in L2: true at 512^2 rows per plane
  // (21 MB), not at 1024^2 (84 MB against a 2 x 63 MB L2).  Then the tiles of a plane are cut into blocks of
  // blk_yb tiles and the sweep runs through ALL planes of one block before moving to the next block, which
  // brings the re-use distance back to 2 * blk_yb tiles.
  L.blk_tpp = L.blk_yb = 0;
  if (K >= 2 && off[K - 1] > kDiaBlockRows && off[K - 1] % T == 0) {
    const int tpp = (int)(off[K - 1] / T);
    int yb = kDiaBlockRows / T;
    while (yb > 1 && tpp % yb) --yb;
    if (yb >= 32) { L.blk_tpp = tpp; L.blk_yb = yb; }
  }
  return L;
}
inline size_t dia_tma_smem_bytes(const DiaTmaLayout &L) { return (size_t)L.bar_off + dia_tma_tail_bytes(); }

__device__ __forceinline__ void dia_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kDiaTmaRows) : "memory"); }

// Tiles are dealt in the order "interior first, edge tiles last": the tiles at the two ends of the owned range
// are the only ones that read halo values, so a multi-GPU product can start on its interior while the neighbours'
// planes are still in flight and wait for them (FusedWait) only when it gets there.
// Big planes (L.blk_tpp > 0): the interior tiles are additionally taken block by block -- all planes of the
// first blk_yb tiles of a plane, then all planes of the next blk_yb tiles, ... (see dia_tma_layout).
struct DiaTileOrder {
  int first_int, n_int;  // interior tiles are first_int .. first_int + n_int - 1
  int tpp, yb, n_blk;    // blocked part: the first n_blk = (n_int / tpp) * tpp interior tiles (0: no blocking)
  __device__ __forceinline__ int operator()(int v) const {
    if (v < n_int) {
      if (v < n_blk) {
        const int np = n_blk / tpp;            // planes
        const int per_block = np * yb;
        const int b = v / per_block, rem = v - b * per_block;
        const int p = rem / yb, j = rem - p * yb;
        return first_int + p * tpp + b * yb + j;
      }
      return first_int + v;
    }
    const int e = v - n_int;
    return e < first_int ? e : e + n_int;
  }
};
__device__ __forceinline__ DiaTileOrder dia_tile_order(int nrows, const DiaTmaLayout &L) {
  constexpr int T = kDiaTmaTile;
  DiaTileOrder o;
  const int64_t first = (L.reach + T - 1) / T;
  const int64_t last = ((int64_t)nrows - T - L.reach) >= 0 ? ((int64_t)nrows - T - L.reach) / T : -1;
  o.first_int = (int)first;
  o.n_int = last >= first ? (int)(last - first + 1) : 0;
  if (o.n_int == 0) o.first_int = 0;
  o.tpp = L.blk_tpp; o.yb = L.blk_yb;
  o.n_blk = (L.blk_tpp > 0 && o.n_int >= 2 * L.blk_tpp) ? (o.n_int / L.blk_tpp) * L.blk_tpp : 0;
  return o;
}

// one row straight from global memory (edge tiles): identical arithmetic to k_spmv_dia
// (halo columns through L2: their values may have arrived from the neighbour GPU while this kernel was running)
template <int K, bool UNIT>
__device__ __forceinline__ double dia_row_direct(const DiaDesc &D, const double *__restrict__ x, int64_t r, double xr) {
  double acc = 0.0;
#pragma unroll
  for (int k = K - 1; k >= 0; --k) {
    const double lo = __ldg(&D.U[k][r]);
    if (lo != 0.0) {
      const int64_t il = r - D.off[k];
      const double xv = il >= 0 ? __ldg(&x[il]) : __ldcg(&x[D.xindex(D.row_start + il)]);
      acc = __dadd_rn(acc, __dmul_rn(lo, xv));
    }
  }
  acc = __dadd_rn(acc, UNIT ? xr : __dmul_rn(__ldg(&D.diag[r]), xr));
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double up = __ldg(&D.U[k][D.off[k] + r]);
    if (up != 0.0) {
      const int64_t iu = r + D.off[k];
      const double xv = iu < D.nf ? __ldg(&x[iu]) : __ldcg(&x[D.xindex(D.row_start + iu)]);
      acc = __dadd_rn(acc, __dmul_rn(up, xv));
    }
  }
  return acc;
}

template <bool DOT, int K, bool UNIT>
__global__ void __launch_bounds__(kDiaTmaThreads, kDiaTmaCtasPerSm)
k_spmv_dia_tma(int nrows, DiaDesc D, DiaTmaLayout L, const double *__restrict__ x, double *__restrict__ y,
               const double *__restrict__ Dvec, double sigma, double *__restrict__ partials, unsigned int *ticket,
               PcgScal *scal, int finalize_mode, PeerRed pr, FusedWait fw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (DOT && scal->done) return;
  constexpr int T = kDiaTmaTile;
  double *const stage0 = reinterpret_cast<double *>(smem_raw);
  uint64_t *const full = reinterpret_cast<uint64_t *>(smem_raw + L.bar_off);
  uint64_t *const empty = full + kDiaTmaMaxStages;
  double *const wsum = reinterpret_cast<double *>(empty + kDiaTmaMaxStages);
  const int nstages = L.stages, M = L.margin;
  int *const is_last = reinterpret_cast<int *>(wsum + kDiaTmaRows / 32);
  const int t = threadIdx.x;
  const int ntiles = (nrows + T - 1) / T;
  const DiaTileOrder order = dia_tile_order(nrows, L);
  if (t == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kDiaTmaRows / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (t >= kDiaTmaRows) {
    // ---------------- producer warp: one elected lane drives the TMA unit ----------------
    if (t == kDiaTmaRows) {
      int stage = 0, use = 0;  // use = how often this stage has been filled before
      for (int v = blockIdx.x; v < ntiles; v += gridDim.x) {
        const int tile = order(v);
        if (use > 0) mbar_wait(&empty[stage], (uint32_t)((use - 1) & 1));
        const int cur = stage;
        if (++stage == nstages) { stage = 0; ++use; }
        const int64_t r0 = (int64_t)tile * T;
        const bool interior = r0 >= L.reach && r0 + T + L.reach <= (int64_t)nrows;
        if (!interior) {
          mbar_expect_tx(&full[cur], 0u);  // nothing to copy: the consumers read global memory
          continue;
        }
        double *const sm = stage0 + (size_t)cur * L.stage_doubles;
        mbar_expect_tx(&full[cur], L.tx_bytes);
        bulk_g2s(sm + L.xc, x + (r0 - M), (uint32_t)(T + 2 * M) * 8u, &full[cur]);
        if (!UNIT) bulk_g2s(sm + L.dg, D.diag + r0, T * 8u, &full[cur]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int64_t o = D.off[k];
          if (o <= kDiaNearMax) {
            bulk_g2s(sm + L.un[k], D.U[k] + r0, (uint32_t)(T + (int)((o + 1) & ~(int64_t)1)) * 8u, &full[cur]);
          } else {
            const int64_t sh = o & 1;
            bulk_g2s(sm + L.xl[k], x + (r0 - o - sh), (T + 2) * 8u, &full[cur]);
            bulk_g2s(sm + L.xu[k], x + (r0 + o - sh), (T + 2) * 8u, &full[cur]);
            bulk_g2s(sm + L.ul[k], D.U[k] + r0, T * 8u, &full[cur]);
            bulk_g2s(sm + L.uu[k], D.U[k] + (r0 + o - sh), (T + 2) * 8u, &full[cur]);
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers: thread t owns rows r0 + t, r0 + kDiaTmaRows + t of every tile ----------------
  double dot = 0.0;
  int nxt = 0, use = 0;
  bool halo_here = fw.n == 0;
  for (int v = blockIdx.x; v < ntiles; v += gridDim.x) {
    const int tile = order(v);
    const int stage = nxt;
    const uint32_t parity = (uint32_t)(use & 1);
    if (++nxt == nstages) { nxt = 0; ++use; }
    const int64_t r0 = (int64_t)tile * T;
    const bool interior = r0 >= L.reach && r0 + T + L.reach <= (int64_t)nrows;
    mbar_wait(&full[stage], parity);
    double acc[kDiaTmaR], xr[kDiaTmaR];
    if (interior) {
      const double *const sm = stage0 + (size_t)stage * L.stage_doubles;
#pragma unroll
      for (int j = 0; j < kDiaTmaR; ++j) {
        const int i = j * kDiaTmaRows + t;
        xr[j] = sm[L.xc + M + i];
        double a = 0.0;
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {  // most negative column first
          const int o = (int)D.off[k];
          double lo, xv;
          if (D.off[k] <= kDiaNearMax) { lo = sm[L.un[k] + i]; xv = sm[L.xc + M + i - o]; }
          else { lo = sm[L.ul[k] + i]; xv = sm[L.xl[k] + i + (o & 1)]; }
          if (lo != 0.0) a = __dadd_rn(a, __dmul_rn(lo, xv));
        }
        a = __dadd_rn(a, UNIT ? xr[j] : __dmul_rn(sm[L.dg + i], xr[j]));
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int o = (int)D.off[k];
          double up, xv;
          if (D.off[k] <= kDiaNearMax) { up = sm[L.un[k] + i + o]; xv = sm[L.xc + M + i + o]; }
          else { up = sm[L.uu[k] + i + (o & 1)]; xv = sm[L.xu[k] + i + (o & 1)]; }
          if (up != 0.0) a = __dadd_rn(a, __dmul_rn(up, xv));
        }
        acc[j] = a;
      }
    } else {
      if (!halo_here) {  // first edge tile of this CTA: the neighbours' planes must have landed (CTA-uniform branch)
        if (t < fw.n && !wait_flag(fw.flag[t], fw.seq)) { *fw.error = 1; scal->done = 1; scal->converged = 0; }
        dia_consumer_sync();
        halo_here = true;
      }
#pragma unroll
      for (int j = 0; j < kDiaTmaR; ++j) {
        const int64_t r = r0 + j * kDiaTmaRows + t;
        xr[j] = 0.0;
        acc[j] = 0.0;
        if (r < nrows) {
          xr[j] = x[r];
          acc[j] = dia_row_direct<K, UNIT>(D, x, r, xr[j]);
        }
      }
    }
    // the stage can be refilled as soon as every lane of this warp has read it
    __syncwarp();
    if ((t & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
#pragma unroll
    for (int j = 0; j < kDiaTmaR; ++j) {
      const int64_t r = r0 + j * kDiaTmaRows + t;
      if (r < nrows) {
        double a = acc[j];
        if (!UNIT && sigma != 0.0) a += sigma * (Dvec ? Dvec[r] : 1.0) * xr[j];
        y[r] = a;
        dot += xr[j] * a;
      }
    }
  }
  if (DOT) {
    // u.Au: one partial per CTA; the CTA drawing the last ticket folds them in CTA order
    double s = warp_sum(dot);
    if ((t & 31) == 0) wsum[t >> 5] = s;
    dia_consumer_sync();
    if (t == 0) {
      double tot = 0.0;
      for (int w = 0; w < kDiaTmaRows / 32; ++w) tot += wsum[w];
      partials[blockIdx.x] = tot;
      __threadfence();
      *is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    dia_consumer_sync();
    if (*is_last) {
      __threadfence();
      double q = 0.0;
      for (unsigned int i = t; i < gridDim.x; i += kDiaTmaRows) q += __ldcg(&partials[i]);
      q = warp_sum(q);
      dia_consumer_sync();
      if ((t & 31) == 0) wsum[t >> 5] = q;
      dia_consumer_sync();
      if (t < 32) {
        double tot = 0.0;
        if (t == 0)
          for (int w = 0; w < kDiaTmaRows / 32; ++w) tot += wsum[w];
        const bool ok = finalize_mode != 2 || peer_allreduce_warp(pr, &tot, 1, scal);
        if (t == 0) {
          scal->red[0] = tot;
          if (ok && finalize_mode >= 1) scal->uc = tot;
        }
      }
    }
  }
}

}  // namespace fvb
