// box.cuh -- assembly straight into the symmetric-diagonal format for problems whose face list is the one
// `regulargrid` emits (src/grid.jl:56-110): nodes z-fastest, every node pushing its +x, +y, +z faces.
//
// For such a list everything the general path (assemble.cuh) discovers by counting, scanning and sorting is
// known in closed form: the faces incident to node (i1,i2,i3) are, IN ASCENDING FACE INDEX,
//     -x, -y, -z   (emitted earlier by the nodes n - n2*n3, n - n3, n - 1)      and   +x, +y, +z  (its own),
// which is exactly the order in which sparse! left-folds the node's diagonal and assembleb adds the Dirichlet
// fluxes (src/FiniteVolume.jl:94-137); each off-diagonal entry has a single contribution -c.  So one thread per
// free row reads its six conductances and writes the row: no adjacency, no atomics on values, no per-row sort,
// no CSR (12*nnz + 4*Nf bytes never written; built lazily by k_box_csr_* only if a caller asks for A).
// The values are bit-identical to the general path's (same folds, __dadd_rn/__dmul_rn), which the tests check.
//
// Two sources of the per-face conductance:
//   * FaceFromArray: c_i = cface[i] computed from the caller's conductivities/areasoverlengths arrays by
//     k_face_conductance, addressed through the closed-form face index (the drop-in call: the neighbor list is
//     only VERIFIED against the closed form, k_box_check, never used for addressing);
//   * FaceImplicit: no face array exists at all -- aol from the grid spacing exactly as regulargrid computes
//     it, face K as the (log-)mean of the two node values exactly as nodehycos2neighborhycos does
//     (src/grid.jl:14-33), c = [exp](k) * aol.  This is what lets 1024^3 (3.2e9 faces) run on 1-2 GPUs.
//
// Requirements checked before this path is taken (else the general path runs): the face list equals the
// closed form for the owned x-planes; the map node -> free row shifts uniformly across every face joining two
// free nodes (so all entries sit on the diagonals 1, n3, n2*n3 of the free-index space -- true for Dirichlet
// sets made of leading/trailing nodes such as the left/right planes of examples/box_model/ex.jl:27-37); the
// x-planes just outside the owned range are entirely free (halo) or entirely Dirichlet.
#pragma once
#include "common.cuh"
#include "grid.cuh"

namespace fvb {

struct BoxDesc {
  GridDesc G;            // n1,n2,n3, spacings, owned planes p_lo..p_hi, first emitting plane e_lo
  int lo_kind, hi_kind;  // plane below / above the owned range: 0 none, 1 entirely free (halo), 2 entirely Dirichlet
  long long n_own;       // owned nodes
  long long nd_owned;    // Dirichlet nodes among them
};

// ---- verification of the caller's neighbor list + the uniform-shift condition ----------------------------------
// One thread per emitting node (halo plane included).  flag[0] = 1: the list is not regulargrid's;
// flag[1] = 1: two free owned nodes joined by a face have different node->row shifts.
__global__ void __launch_bounds__(kBlock)
k_box_check(BoxDesc B, const longlong2 *__restrict__ nb, const int *__restrict__ nodemap, int *__restrict__ flag) {
  const GridDesc &G = B.G;
  const long long plane = G.n2 * G.n3;
  const long long nodes = (G.p_hi - G.e_lo + 1) * plane;
  const bool halo_plane = G.e_lo < G.p_lo;
  const long long base_owned = faces_before(G, G.p_lo, 1, 1);
  bool bad_list = false, bad_shift = false;
  // (node counts of a rank fit 32 bits: 32-bit divisions, several times cheaper than 64-bit ones)
  const unsigned plane32 = (unsigned)plane, n332 = (unsigned)G.n3;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nodes; t += (long long)gridDim.x * blockDim.x) {
    const unsigned q1 = (unsigned)t / plane32, rem32 = (unsigned)t - q1 * plane32, q2 = rem32 / n332;
    const long long i1 = G.e_lo + q1;
    const long long rem = rem32;
    const long long i2 = q2 + 1, i3 = rem32 - q2 * n332 + 1;
    const long long lin = i3 + G.n3 * (i2 - 1) + plane * (i1 - 1);
    if (i1 < G.p_lo) {
      const longlong2 p = nb[rem];
      bad_list |= p.x != lin || p.y != lin + plane;
      continue;
    }
    const long long tl = (i1 - G.p_lo) * plane + rem;  // local node index
    const int m = nodemap[tl];
    const long long sh = tl - m;                       // (meaningful when m >= 0)
    long long j = (halo_plane ? plane : 0) + faces_before(G, i1, i2, i3) - base_owned;
    if (i1 < G.n1) {
      const longlong2 p = nb[j++];
      bad_list |= p.x != lin || p.y != lin + plane;
      if (i1 < G.p_hi && m >= 0) { const int mo = nodemap[tl + plane]; bad_shift |= mo >= 0 && (tl + plane - mo) != sh; }
    }
    if (i2 < G.n2) {
      const longlong2 p = nb[j++];
      bad_list |= p.x != lin || p.y != lin + G.n3;
      if (m >= 0) { const int mo = nodemap[tl + G.n3]; bad_shift |= mo >= 0 && (tl + G.n3 - mo) != sh; }
    }
    if (i3 < G.n3) {
      const longlong2 p = nb[j++];
      bad_list |= p.x != lin || p.y != lin + 1;
      if (m >= 0) { const int mo = nodemap[tl + 1]; bad_shift |= mo >= 0 && (tl + 1 - mo) != sh; }
    }
  }
  if (bad_list) flag[0] = 1;
  if (bad_shift) flag[1] = 1;
}
// the same shift condition when no neighbor list exists (implicit grid)
__global__ void __launch_bounds__(kBlock)
k_box_check_shift(BoxDesc B, const int *__restrict__ nodemap, int *__restrict__ flag) {
  const GridDesc &G = B.G;
  const long long plane = G.n2 * G.n3;
  bool bad = false;
  for (long long tl = (long long)blockIdx.x * blockDim.x + threadIdx.x; tl < B.n_own; tl += (long long)gridDim.x * blockDim.x) {
    const int m = nodemap[tl];
    if (m < 0) continue;
    const long long sh = tl - m;
    const long long i1 = G.p_lo + tl / plane, rem = tl % plane;
    const long long i2 = rem / G.n3 + 1, i3 = rem % G.n3 + 1;
    if (i1 < G.p_hi) { const int mo = nodemap[tl + plane]; bad |= mo >= 0 && (tl + plane - mo) != sh; }
    if (i2 < G.n2) { const int mo = nodemap[tl + G.n3]; bad |= mo >= 0 && (tl + G.n3 - mo) != sh; }
    if (i3 < G.n3) { const int mo = nodemap[tl + 1]; bad |= mo >= 0 && (tl + 1 - mo) != sh; }
  }
  if (bad) flag[1] = 1;
}

// ---- where a face's conductance comes from ------------------------------------------------------------------------
// dir: 0 -x, 1 -y, 2 -z, 3 +x, 4 +y, 5 +z of node (i1,i2,i3) with in-plane index rem and local index tl.
struct FaceFromArray {
  const double *cface;     // per-face conductance in the order of the (slab's) face list
  __device__ __forceinline__ double operator()(const BoxDesc &B, int dir, long long i1, long long i2, long long i3,
                                               long long rem, long long tl) const {
    return cface[index(B, dir, i1, i2, i3, rem, tl)];
  }
  // position of that face in the (slab's) face list
  __device__ __forceinline__ long long index(const BoxDesc &B, int dir, long long i1, long long i2, long long i3,
                                             long long rem, long long /*tl*/) const {
    const GridDesc &G = B.G;
    const long long plane = G.n2 * G.n3;
    const long long shift = (G.e_lo < G.p_lo ? plane : 0) - faces_before(G, G.p_lo, 1, 1);
    const long long hx = i1 < G.n1, hy = i2 < G.n2;
    long long j;
    switch (dir) {
      case 0: j = i1 == G.p_lo ? rem : faces_before(G, i1 - 1, i2, i3) + shift; break;
      case 1: j = faces_before(G, i1, i2 - 1, i3) + hx + shift; break;
      case 2: j = faces_before(G, i1, i2, i3 - 1) + hx + hy + shift; break;
      case 3: j = faces_before(G, i1, i2, i3) + shift; break;
      case 4: j = faces_before(G, i1, i2, i3) + hx + shift; break;
      default: j = faces_before(G, i1, i2, i3) + hx + hy + shift; break;
    }
    return j;
  }
};

struct FaceImplicit {
  const double *nodek;     // node values of planes k_plane_lo .. (node order), covering p_lo-1 .. p_hi+1 where they exist
  long long kofs;          // nodek index of local node 0  (= node_lo - first node held)
  int logmean, logk;       // nodehycos2neighborhycos(..., logmean); logtransformconductivity
  __device__ __forceinline__ double operator()(const BoxDesc &B, int dir, long long i1, long long i2, long long i3,
                                               long long /*rem*/, long long tl) const {
    const GridDesc &G = B.G;
    const long long plane = G.n2 * G.n3;
    const double wx = (i1 == 1 || i1 == G.n1) ? G.dx * 0.5 : G.dx;
    const double wy = (i2 == 1 || i2 == G.n2) ? G.dy * 0.5 : G.dy;
    const double wz = (i3 == 1 || i3 == G.n3) ? G.dz * 0.5 : G.dz;
    long long to;
    double aol;
    switch (dir) {  // the widths across the face normal are those of either endpoint (src/grid.jl:91-105)
      case 0: to = tl - plane; aol = __ddiv_rn(__dmul_rn(wy, wz), G.dx); break;
      case 1: to = tl - G.n3; aol = __ddiv_rn(__dmul_rn(wx, wz), G.dy); break;
      case 2: to = tl - 1; aol = __ddiv_rn(__dmul_rn(wx, wy), G.dz); break;
      case 3: to = tl + plane; aol = __ddiv_rn(__dmul_rn(wy, wz), G.dx); break;
      case 4: to = tl + G.n3; aol = __ddiv_rn(__dmul_rn(wx, wz), G.dy); break;
      default: to = tl + 1; aol = __ddiv_rn(__dmul_rn(wx, wy), G.dz); break;
    }
    const double ka = nodek[kofs + tl], kb = nodek[kofs + to];
    double k = logmean ? __dmul_rn(0.5, __dadd_rn(ka, kb)) : sqrt(__dmul_rn(ka, kb));
    if (logk) k = exp(k);
    return __dmul_rn(k, aol);
  }
};

// Dirichlet slot of a node outside the owned range (sorted table of the whole problem; last duplicate wins)
__device__ __forceinline__ int box_offrank_slot(long long node0, const int64_t *__restrict__ dsorted,
                                                const int *__restrict__ dslot, int64_t nd) {
  int64_t lo = 0, hi = nd;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (dsorted[mid] < node0) lo = mid + 1; else hi = mid;
  }
  return (lo < nd && dsorted[lo] == node0) ? dslot[lo] : -1;
}

// ---- the rows ----------------------------------------------------------------------------------------------------------
// One thread per owned node; free nodes write their row: U_k upper entries (0 where there is none), the
// lower entries whose partner row lives on the rank below, diag, b, and a 6-bit presence mask (bit d set:
// the entry towards direction d is stored -- explicit zeros included -- which is what the lazy CSR needs).
// o1 = n3, o2 = n2*n3 are the diagonal offsets 1, o1, o2 (U0, U1, U2).
template <class FaceC>
__global__ void __launch_bounds__(kBlock)
k_box_values(BoxDesc B, FaceC fc, const int *__restrict__ nodemap, const double *__restrict__ sources,
             const double *__restrict__ dheads, const int64_t *__restrict__ dsorted,
             const int *__restrict__ dsorted_slot, int64_t nd_sorted, double *__restrict__ U0,
             double *__restrict__ U1, double *__restrict__ U2, double *__restrict__ diag, double *__restrict__ b,
             uint8_t *__restrict__ mask, unsigned long long *__restrict__ nnz_total) {
  const GridDesc &G = B.G;
  const long long plane = G.n2 * G.n3, o1 = G.n3, o2 = plane;
  const long long node_lo0 = (G.p_lo - 1) * plane;  // 0-based global index of local node 0
  unsigned long long my_nnz = 0;
  for (long long tl = (long long)blockIdx.x * blockDim.x + threadIdx.x; tl < B.n_own; tl += (long long)gridDim.x * blockDim.x) {
    const int m = nodemap[tl];
    if (m < 0) continue;
    const long long r = m;
    const unsigned q1 = (unsigned)tl / (unsigned)plane, rem32 = (unsigned)tl - q1 * (unsigned)plane, q2 = rem32 / (unsigned)G.n3;
    const long long i1 = G.p_lo + q1, rem = rem32;
    const long long i2 = q2 + 1, i3 = rem32 - q2 * (unsigned)G.n3 + 1;
    double dv = 0.0, bv = sources ? sources[tl] : 0.0;
    bool first = true;
    unsigned mk = 0;
    // returns the stored off-diagonal value (or 0 with *present = false)
    auto visit = [&](int dir, long long to, int off_kind) -> double {
      const double c = fc(B, dir, i1, i2, i3, rem, tl);
      if (first) { dv = c; first = false; } else dv = __dadd_rn(dv, c);
      if (off_kind == 0) {                       // neighbour owned
        const int mo = nodemap[to];
        if (mo >= 0) { mk |= 1u << dir; return -c; }
        bv = __dadd_rn(bv, __dmul_rn(c, dheads[-1 - mo]));
        return 0.0;
      }
      if (off_kind == 1) { mk |= 1u << dir; return -c; }  // free node of the neighbouring rank
      const int slot = box_offrank_slot(node_lo0 + to, dsorted, dsorted_slot, nd_sorted);
      if (slot >= 0) bv = __dadd_rn(bv, __dmul_rn(c, dheads[slot]));
      return 0.0;
    };
    // ascending face index: -x, -y, -z, +x, +y, +z
    if (i1 > 1) {
      const bool own = i1 > G.p_lo;
      const double v = visit(0, tl - plane, own ? 0 : B.lo_kind);
      if (!own && (mk & 1u)) U2[r] = v;          // lower entry whose partner row is owned by the rank below
    }
    if (i2 > 1) visit(1, tl - o1, 0);
    if (i3 > 1) visit(2, tl - 1, 0);
    double ux = 0.0, uy = 0.0, uz = 0.0;
    if (i1 < G.n1) ux = visit(3, tl + plane, i1 < G.p_hi ? 0 : B.hi_kind);
    if (i2 < G.n2) uy = visit(4, tl + o1, 0);
    if (i3 < G.n3) uz = visit(5, tl + 1, 0);
    U0[1 + r] = uz;
    U1[o1 + r] = uy;
    U2[o2 + r] = ux;
    diag[r] = dv;
    b[r] = bv;
    mask[r] = (uint8_t)mk;
    my_nnz += 1u + __popc(mk);
  }
  // stored entries of this rank (integer sum: order-independent)
  for (int o = 16; o > 0; o >>= 1) my_nnz += __shfl_down_sync(0xffffffffu, my_nnz, o);
  if ((threadIdx.x & 31) == 0 && my_nnz) atomicAdd(nnz_total, my_nnz);
}

// ---- lazy CSR from the diagonals (fvb_get_csr, forced CSR format) -------------------------------------------------
__global__ void k_box_csr_count(int n, const uint8_t *__restrict__ mask, int *__restrict__ cnt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) cnt[r] = 1 + __popc((unsigned)mask[r]);
}
// columns ascending: r-o2, r-o1, r-1, r, r+1, r+o1, r+o2; off-rank columns are halo slots
// [n, n+nlo) (rank below, plane order) and [n+nlo, n+nlo+nhi) (rank above).
__global__ void k_box_csr_fill(int n, const uint8_t *__restrict__ mask, const int *__restrict__ rowptr, long long o1,
                               long long o2, long long nlo, const double *__restrict__ U0,
                               const double *__restrict__ U1, const double *__restrict__ U2,
                               const double *__restrict__ diag, int *__restrict__ colidx, double *__restrict__ vals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const unsigned mk = mask[r];
  int w = rowptr[r];
  if (mk & 1u) { colidx[w] = r >= o2 ? (int)(r - o2) : (int)(n + (r - o2 + nlo)); vals[w++] = U2[r]; }
  if (mk & 2u) { colidx[w] = (int)(r - o1); vals[w++] = U1[r]; }
  if (mk & 4u) { colidx[w] = r - 1; vals[w++] = U0[r]; }
  colidx[w] = r; vals[w++] = diag[r];
  if (mk & 32u) { colidx[w] = r + 1; vals[w++] = U0[1 + r]; }
  if (mk & 16u) { colidx[w] = (int)(r + o1); vals[w++] = U1[o1 + r]; }
  if (mk & 8u) { colidx[w] = r + o2 < n ? (int)(r + o2) : (int)(n + nlo + (r + o2 - n)); vals[w++] = U2[o2 + r]; }
}

}  // namespace fvb
