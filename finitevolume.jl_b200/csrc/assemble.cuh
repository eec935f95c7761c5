// assemble.cuh -- deterministic face list -> CSR build.
//
// What the reference does serially (src/FiniteVolume.jl:75-139): push up to four COO
// triples per face in face order, then SparseArrays.sparse(I,J,V,Nf,Nf,+) counting-sorts
// them and left-folds duplicates in push order, keeping explicit zeros, columns ascending.
//
// What happens here instead (no COO is ever materialised, no floating-point atomics):
//   1. nodemap: Dirichlet slots by atomicMax (last duplicate wins, :23), free-row ranks by
//      a prefix scan (:36-42).
//   2. per-face conductance c_i = [exp](k[meta(i)]) * aol[i]   (:83/:96), one rounding each.
//   3. node -> incident-face adjacency: integer-atomic counts, scan, integer-atomic fill
//      (slot order inside a row is arbitrary at this point).
//   4. per row: sort the row's entries by (global column, face index).  From here on
//      everything is a pure function of the inputs: the column set, and for every stored
//      entry the left fold of its contributions in ascending face order -- exactly the
//      order sparse! folds them in, because a row's triples are pushed in face order.
//   5. count distinct columns, scan -> rowptr, then one thread per row writes columns and
//      folded values, diag(A) and b (= sources + sum c*head_D in face order, :113-137).
// All folds use __dadd_rn/__dmul_rn so the compiler cannot contract them into FMAs
// (Julia does not), which is what makes values bit-identical to the CPU oracle.
#pragma once
#include "common.cuh"

namespace fvb {

enum { ERR_BAD_NODE = 0, ERR_SRC_ON_DIRICHLET = 1, ERR_BAD_META = 2, ERR_COUNT = 4 };

// ---- 1. node bookkeeping ---------------------------------------------------------------
// dslot[i] = largest k with dnodes[k] == node_lo+i, else -1.  (getnodei2dirichleti :20-30)
__global__ void k_mark_dirichlet(const int64_t *__restrict__ dnodes, int64_t nd, int64_t n_nodes,
                                 int64_t node_lo, int64_t node_hi, const double *__restrict__ sources,
                                 int *__restrict__ dslot, int *__restrict__ err) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nd) return;
  int64_t node = dnodes[k] - 1;
  if (node < 0 || node >= n_nodes) { atomicMin(&err[ERR_BAD_NODE], (int)min((int64_t)INT_MAX - 1, k)); return; }
  if (node < node_lo || node >= node_hi) return;
  atomicMax(&dslot[node - node_lo], (int)k);
  if (sources && sources[node - node_lo] != 0.0) atomicMin(&err[ERR_SRC_ON_DIRICHLET], (int)k);
}

// flag[i] = 1 for free nodes (input of the scan).
__global__ void k_free_flags(const int *__restrict__ dslot, int64_t n, int *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = dslot[i] < 0;
}

// nodemap[i] = free rank (>=0) or -1-slot; row2node[rank] = i.   (getfreenodes :32-44)
__global__ void k_finish_nodemap(const int *__restrict__ dslot, const int *__restrict__ rank, int64_t n,
                                 int *__restrict__ nodemap, int *__restrict__ row2node) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = dslot[i];
  if (s < 0) {
    int r = rank[i];
    nodemap[i] = r;
    row2node[r] = (int)i;
  } else {
    nodemap[i] = -1 - s;
  }
}

// ---- 2. per-face conductance -------------------------------------------------------------
__global__ void k_face_conductance(int64_t nf, const double *__restrict__ cond, int64_t n_cond,
                                   const int64_t *__restrict__ meta, const double *__restrict__ aol,
                                   int logk, double *__restrict__ cface, int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf) return;
  int64_t m = meta ? meta[i] - 1 : i;
  if (m < 0 || m >= n_cond) { atomicMin(&err[ERR_BAD_META], (int)min((int64_t)INT_MAX - 1, i)); return; }
  double k = cond[m];
  if (logk) k = exp(k);
  cface[i] = __dmul_rn(k, aol[i]);
}

// ---- endpoint resolution -------------------------------------------------------------------
// Kind of a face endpoint as seen from this rank.
struct NodeRef {
  int kind;      // 0 owned free, 1 Dirichlet, 2 off-rank free, 3 invalid
  int32_t local; // kind 0: local row; kind 1: dirichlet slot
  int64_t glob;  // kind 2: 0-based global free index
};

struct Resolver {
  const int *nodemap;
  int64_t n_nodes, node_lo, node_hi;
  const int64_t *dsorted;  // ascending distinct Dirichlet nodes (0-based)
  const int *dsorted_slot; // their slots (last duplicate wins)
  int64_t nd_sorted;

  __device__ __forceinline__ NodeRef operator()(int64_t node) const {
    NodeRef r;
    r.local = 0; r.glob = 0;
    if (node < 0 || node >= n_nodes) { r.kind = 3; return r; }
    if (node >= node_lo && node < node_hi) {
      int m = nodemap[node - node_lo];
      if (m >= 0) { r.kind = 0; r.local = m; } else { r.kind = 1; r.local = -1 - m; }
      return r;
    }
    // off-rank: #Dirichlet nodes below `node` by binary search (lower_bound)
    int64_t lo = 0, hi = nd_sorted;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (dsorted[mid] < node) lo = mid + 1; else hi = mid;
    }
    if (lo < nd_sorted && dsorted[lo] == node) { r.kind = 1; r.local = dsorted_slot[lo]; }
    else { r.kind = 2; r.glob = node - lo; }
    return r;
  }
};

// ---- 3. adjacency -----------------------------------------------------------------------------
// pass 0: count entries per owned row and off-rank references.
// pass 1: fill (positions by integer atomics; order fixed later by the per-row sort).
template <int PASS>
__global__ void k_adjacency(int64_t nfaces, const int64_t *__restrict__ nb, Resolver res,
                            int *__restrict__ cnt_or_cursor, const int *__restrict__ adjptr,
                            int *__restrict__ adj_face, int *__restrict__ adj_col,
                            unsigned long long *__restrict__ n_offrank, int64_t *__restrict__ offrank_refs,
                            const int64_t *__restrict__ halo_glob, int64_t n_halo, int nf_local,
                            int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfaces) return;
  // 16-byte vector load of the interleaved Pair{Int64,Int64}
  longlong2 p = reinterpret_cast<const longlong2 *>(nb)[i];
  NodeRef a = res(p.x - 1), b = res(p.y - 1);
  if (a.kind == 3 || b.kind == 3) { atomicMin(&err[ERR_BAD_NODE], (int)min((int64_t)INT_MAX - 1, i)); return; }
  auto encode = [&](const NodeRef &o) -> int {
    if (o.kind == 0) return o.local;
    if (o.kind == 1) return -1 - o.local;
    // off-rank free column: position in the sorted halo list
    int64_t lo = 0, hi = n_halo;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (halo_glob[mid] < o.glob) lo = mid + 1; else hi = mid;
    }
    return nf_local + (int)lo;
  };
  auto emit = [&](const NodeRef &self, const NodeRef &other) {
    if (PASS == 0) {
      atomicAdd(&cnt_or_cursor[self.local], 1);
      if (other.kind == 2) {
        unsigned long long s = atomicAdd(n_offrank, 1ULL);
        if (offrank_refs) offrank_refs[s] = other.glob;
      }
    } else {
      int pos = adjptr[self.local] + atomicAdd(&cnt_or_cursor[self.local], 1);
      adj_face[pos] = (int)i;
      adj_col[pos] = encode(other);
    }
  };
  if (a.kind == 0 && b.kind == 0 && a.local == b.local) {
    emit(a, b);  // self loop: one entry, handled as +c,-c,+c,-c in the fold
  } else {
    if (a.kind == 0) emit(a, b);
    if (b.kind == 0) emit(b, a);
  }
}

// ---- 4./5. per-row work ------------------------------------------------------------------------
constexpr int kMaxDeg = 32;  // rows up to this degree are sorted in registers/local memory

struct ColKey {
  int nf_local;
  int64_t row_start;
  const int64_t *halo_glob;
  // global column used for ordering; Dirichlet entries (negative) sort first
  __device__ __forceinline__ int64_t operator()(int c) const {
    if (c < 0) return -1;
    return c < nf_local ? row_start + c : halo_glob[c - nf_local];
  }
};

// Sort each row's entries by (column key, face), write them back, count stored entries.
__global__ void k_row_structure(int nf_local, const int *__restrict__ adjptr, int *__restrict__ adj_face,
                                int *__restrict__ adj_col, ColKey key, int *__restrict__ row_nnz) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nf_local) return;
  const int lo = adjptr[r], d = adjptr[r + 1] - lo;
  const int64_t self = key.row_start + r;
  int nnz = d > 0 ? 1 : 0;  // the diagonal
  if (d <= kMaxDeg) {
    int f[kMaxDeg];
    int c[kMaxDeg];
    int64_t k[kMaxDeg];
    for (int j = 0; j < d; ++j) {
      int fj = adj_face[lo + j], cj = adj_col[lo + j];
      int64_t kj = key(cj);
      int p = j;
      while (p > 0 && (k[p - 1] > kj || (k[p - 1] == kj && f[p - 1] > fj))) {
        k[p] = k[p - 1]; f[p] = f[p - 1]; c[p] = c[p - 1];
        --p;
      }
      k[p] = kj; f[p] = fj; c[p] = cj;
    }
    int64_t prev = -1;
    for (int j = 0; j < d; ++j) {
      adj_face[lo + j] = f[j];
      adj_col[lo + j] = c[j];
      if (k[j] >= 0 && k[j] != self && k[j] != prev) ++nnz;
      prev = k[j];
    }
  } else {
    // long rows: in-place insertion sort in global memory (rare: irregular hubs)
    for (int j = 1; j < d; ++j) {
      int fj = adj_face[lo + j], cj = adj_col[lo + j];
      int64_t kj = key(cj);
      int p = j;
      while (p > 0) {
        int fq = adj_face[lo + p - 1], cq = adj_col[lo + p - 1];
        int64_t kq = key(cq);
        if (kq > kj || (kq == kj && fq > fj)) {
          adj_face[lo + p] = fq; adj_col[lo + p] = cq;
          --p;
        } else break;
      }
      adj_face[lo + p] = fj; adj_col[lo + p] = cj;
    }
    int64_t prev = -1;
    for (int j = 0; j < d; ++j) {
      int64_t kj = key(adj_col[lo + j]);
      if (kj >= 0 && kj != self && kj != prev) ++nnz;
      prev = kj;
    }
  }
  row_nnz[r] = nnz;
}

// One thread per row: columns (optional) + folded values + diag + b.
// Off-diagonal (r,c): fold of -c_i over the faces joining r and c, ascending face index.
// Diagonal: fold of +c_i over ALL incident faces in ascending face index (self loops add
// +c,-c,+c,-c, the push order of :84-87).  b: sources[node], then += c_i*head in face order.
__global__ void k_row_values(int nf_local, const int *__restrict__ adjptr, const int *__restrict__ adj_face,
                             const int *__restrict__ adj_col, ColKey key, const double *__restrict__ cface,
                             const double *__restrict__ sources, const double *__restrict__ dheads,
                             const int *__restrict__ row2node, const int *__restrict__ rowptr,
                             int *__restrict__ colidx, double *__restrict__ vals, double *__restrict__ diag,
                             double *__restrict__ b, int write_cols) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nf_local) return;
  const int lo = adjptr[r], d = adjptr[r + 1] - lo;
  double bv = sources[row2node[r]];
  if (d == 0) { b[r] = bv; diag[r] = 0.0; return; }

  // -- diagonal and b: visit entries in ascending face order.  The segment is sorted by
  //    (column, face); repeatedly pick the smallest face index above the last one taken.
  //    Faces are distinct within a row except that no face appears twice (self loops are
  //    stored once), so a strict "greater than last" scan visits each entry exactly once.
  double dv = 0.0;
  bool first = true;
  int last = -1;
  for (int t = 0; t < d; ++t) {
    int best = INT_MAX, bj = -1;
    for (int j = 0; j < d; ++j) {
      int fj = adj_face[lo + j];
      if (fj > last && fj < best) { best = fj; bj = j; }
    }
    last = best;
    const int cj = adj_col[lo + bj];
    const double c = cface[best];
    if (first) { dv = c; first = false; } else dv = __dadd_rn(dv, c);
    if (cj == r) {  // self loop
      dv = __dadd_rn(dv, -c); dv = __dadd_rn(dv, c); dv = __dadd_rn(dv, -c);
    } else if (cj < 0) {
      bv = __dadd_rn(bv, __dmul_rn(c, dheads[-1 - cj]));
    }
  }
  diag[r] = dv;
  b[r] = bv;

  // -- stored entries in ascending column order, diagonal spliced in at its place
  int w = rowptr[r];
  const int64_t self = key.row_start + r;
  bool diag_done = false;
  int j = 0;
  while (j < d && adj_col[lo + j] < 0) ++j;  // Dirichlet entries only feed diag and b
  while (j < d) {
    const int cj = adj_col[lo + j];
    if (cj == r) { ++j; continue; }
    const int64_t kj = key(cj);
    if (!diag_done && kj > self) {
      if (write_cols) colidx[w] = r;
      vals[w++] = dv;
      diag_done = true;
    }
    double v = -cface[adj_face[lo + j]];
    ++j;
    while (j < d && adj_col[lo + j] == cj) { v = __dadd_rn(v, -cface[adj_face[lo + j]]); ++j; }
    if (write_cols) colidx[w] = cj;
    vals[w++] = v;
  }
  if (!diag_done) {
    if (write_cols) colidx[w] = r;
    vals[w] = dv;
  }
}

// rowptr[n+1 .. n+pad] = rowptr[n]: tiles that run past the last row see empty rows.
__global__ void k_fill_tail(int *rowptr, int n, int pad) {
  const int v = rowptr[n];
  for (int i = threadIdx.x; i < pad; i += blockDim.x) rowptr[n + 1 + i] = v;
}

// ---- adjoint gradient gather (src/transientadjointutils.jl:22-32) -------------------------
// endpoints of every face as (local row | -1-dirichlet slot | INT_MIN for anything not owned)
__global__ void k_face_endpoints(int64_t nfaces, const int64_t *__restrict__ nb, Resolver res,
                                 int *__restrict__ e1, int *__restrict__ e2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfaces) return;
  longlong2 p = reinterpret_cast<const longlong2 *>(nb)[i];
  NodeRef a = res(p.x - 1), b = res(p.y - 1);
  e1[i] = a.kind == 0 ? a.local : (a.kind == 1 ? -1 - a.local : INT_MIN);
  e2[i] = b.kind == 0 ? b.local : (b.kind == 1 ? -1 - b.local : INT_MIN);
}

// One thread per face; every accumulator has a single writer.
__global__ void k_gradient_faces(int64_t nfaces, const int *__restrict__ e1, const int *__restrict__ e2,
                                 const double *__restrict__ cface, const double *__restrict__ aol, int logk,
                                 const double *__restrict__ u, const double *__restrict__ lam,
                                 const double *__restrict__ Dvec, const double *__restrict__ dheads, double w,
                                 double *__restrict__ gface, double *__restrict__ gdh) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfaces) return;
  const int a = e1[i], b = e2[i];
  if (a == INT_MIN || b == INT_MIN) return;
  const double c = cface[i];
  const double dc = logk ? c : aol[i];
  if (a >= 0 && b >= 0) {
    if (a == b) return;
    const double la = lam[a] / (Dvec ? Dvec[a] : 1.0), lb = lam[b] / (Dvec ? Dvec[b] : 1.0);
    gface[i] += w * (-dc * (u[a] - u[b]) * (la - lb));
  } else if (a >= 0) {
    const double la = lam[a] / (Dvec ? Dvec[a] : 1.0);
    gface[i] += w * (dc * (dheads[-1 - b] - u[a]) * la);
    gdh[i] += w * (c * la);
  } else if (b >= 0) {
    const double lb = lam[b] / (Dvec ? Dvec[b] : 1.0);
    gface[i] += w * (dc * (dheads[-1 - a] - u[b]) * lb);
    gdh[i] += w * (c * lb);
  }
}

__global__ void k_gradient_rows(int64_t n, const double *__restrict__ lam, const double *__restrict__ Dvec, double w,
                                double *__restrict__ gsrc) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) gsrc[r] += w * lam[r] / (Dvec ? Dvec[r] : 1.0);
}

// which Dirichlet slot a face's head-gradient belongs to (-1: none)
__global__ void k_gradient_dslots(int64_t nfaces, const int *__restrict__ e1, const int *__restrict__ e2,
                                  int64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfaces) return;
  const int a = e1[i], b = e2[i];
  int64_t s = -1;
  if (a != INT_MIN && b != INT_MIN) {
    if (a >= 0 && b < 0) s = -1 - b;
    else if (b >= 0 && a < 0) s = -1 - a;
  }
  out[i] = s < 0 ? -1 : s + 1;  // 1-based on the wire
}

// ---- extraction to the Julia layout ---------------------------------------------------------------
__global__ void k_export_ptr(const int *__restrict__ rowptr, int64_t n1, int64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) out[i] = (int64_t)rowptr[i] + 1;
}
__global__ void k_export_cols(const int *__restrict__ colidx, int64_t nnz, ColKey key, int64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) out[i] = key(colidx[i]) + 1;
}
__global__ void k_export_nodemap(const int *__restrict__ nodemap, int64_t n, int64_t row_start,
                                 uint8_t *__restrict__ freenode, int64_t *__restrict__ n2f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int m = nodemap[i];
  if (freenode) freenode[i] = m >= 0;
  if (n2f) n2f[i] = m >= 0 ? row_start + m + 1 : -1;
}

// head[node] = x[row] on free nodes, prescribed head on Dirichlet nodes
// (freenodes2nodes, src/FiniteVolume.jl:141-155).
__global__ void k_scatter_heads(const int *__restrict__ nodemap, int64_t n, const double *__restrict__ x,
                                const double *__restrict__ dheads, double *__restrict__ head) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int m = nodemap[i];
  head[i] = m >= 0 ? x[m] : dheads[-1 - m];
}

}  // namespace fvb
