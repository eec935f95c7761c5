// common.cuh -- shared declarations of libfvb200 (handle layout, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/fvb200.h"

namespace fvb {

extern thread_local std::string g_last_error;
int set_error(int code, const std::string &msg);

#define FVB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return fvb::set_error(e__ == cudaErrorMemoryAllocation ? FVB_ERR_OOM : FVB_ERR_CUDA, \
                            std::string(#expr) + ": " + cudaGetErrorString(e__));        \
  } while (0)

#define FVB_TRY(expr)              \
  do {                             \
    int s__ = (expr);              \
    if (s__ != FVB_OK) return s__; \
  } while (0)

constexpr int kBlock = 256;

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device-resident scalars of the PCG recurrence (IterativeSolvers.cg 0.8.1 names:
// rho = c.r, uc = u.Au, resid = ||r||).  Lives in device memory; a copy is polled by the
// host every few iterations through pinned memory.
struct PcgScal {
  double rho, rho_prev, uc, resid, resid0, reltol, tol;
  double alpha_prev;   // step length of the last closed iteration; its x += alpha*u is applied lazily
  double red[4];       // local partial sums awaiting the (optional) all-reduce
  long long iter, maxiter, hist_cap;
  int done, converged;
};

class Arena;      // arena.h
struct BoxDesc;   // box.cuh
struct Comm;      // nccl_dyn.h
struct PeerState; // fvb200.cu (peer.cuh tables)
struct MgState;   // fvb200.cu (mg.cuh hierarchy)
struct AmgState;  // fvb200.cu (amg.cuh hierarchy)

}  // namespace fvb

struct fvb_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  int num_sms = 148;
  fvb::Arena *arena = nullptr;  // device-memory arena (arena.h); null: cudaMallocAsync pool (FVB_ARENA=0)

  // ---- partition -------------------------------------------------------------------
  int64_t n_nodes = 0;          // global N
  int64_t node_lo = 0, node_hi = 0;  // 0-based [lo, hi)
  int64_t n_own_nodes = 0;
  int64_t nf_local = 0, nf_global = 0, row_start = 0;  // row_start: 0-based global free idx
  int64_t n_halo = 0, nnz = 0;
  int64_t n_faces = 0, n_dirichlet = 0, n_adj = 0;
  bool assembled = false;

  // ---- retained device arrays ---------------------------------------------------------
  int32_t *nodemap = nullptr;   // [n_own_nodes] >=0 local free row, <0: -1-dirichlet slot
  int32_t *row2node = nullptr;  // [nf_local] local node index of each row
  double *sources = nullptr;    // [n_own_nodes]
  double *dheads = nullptr;     // [n_dirichlet]
  double *aol = nullptr;        // [n_faces]
  int64_t *meta = nullptr;      // [n_faces] or null
  double *cface = nullptr;      // [n_faces] per-face conductance
  int32_t *adjptr = nullptr;    // [nf_local+1]
  int32_t *adj_face = nullptr;  // [n_adj] sorted per row by (column key, face)
  int32_t *adj_col = nullptr;   // [n_adj] local col / nf_local+halo / -1-dirichlet slot
  int64_t *halo_glob = nullptr; // [n_halo] 0-based global free idx, ascending
  int32_t *rowptr = nullptr;    // [nf_local+1]
  int32_t *colidx = nullptr;    // [nnz]
  double *vals = nullptr;       // [nnz]
  double *b = nullptr, *diag = nullptr;  // [nf_local]
  std::vector<int64_t> halo_host;

  // ---- solver workspace ---------------------------------------------------------------
  double *x = nullptr, *r = nullptr, *u = nullptr, *c = nullptr, *dinv = nullptr;  // u: nf_local+n_halo
  double *rhs = nullptr;        // transient right-hand side
  double *Dvec = nullptr;       // Ss*vol per free row, or null (identity)
  double *slots[FVB_NSLOT] = {};
  double *partials = nullptr;   // [2 * max grid] block partial sums
  unsigned int *ticket = nullptr;
  fvb::PcgScal *scal = nullptr;       // device
  fvb::PcgScal *scal_host = nullptr;  // pinned, 2 entries
  double *hist = nullptr;
  int64_t hist_cap = 0;
  double *xio = nullptr, *yio = nullptr;  // staging for fvb_spmv host vectors

  // ---- multi-GPU ------------------------------------------------------------------------
  int nranks = 1, rank = 0;
  fvb::Comm *comm = nullptr;
  std::vector<int> peers;
  std::vector<int64_t> send_counts, recv_counts;
  std::vector<int64_t> send_first;  // per plan peer: first row of its send list when that list is a contiguous run
  bool send_contig = false;         // every peer's send list is one ascending run (slab partitions: a plane)
  bool fused_halo = true;           // FVB_FUSED_HALO_OFF unset (read once at fvb_create: getenv is too slow for the solve loop)
  int32_t *send_rows = nullptr;
  double *sendbuf = nullptr;
  int64_t n_send = 0;
  bool halo_ready = false;
  fvb::PeerState *peer = nullptr;  // NVLink peer-memory exchange state (null: NCCL path)
  int64_t u_cap = 0;               // capacity of u (plain cudaMalloc when nranks > 1: IPC-exportable)

  // index-free diagonal copy of A (dia.cuh), built when the pattern allows
  int fmt_request = 0;          // 0 auto, 1 CSR only, 2 diagonal with per-thread loads only, 3 diagonal TMA forced
  bool last_dia_tma = false;    // the last diagonal SpMV launch used the TMA pipeline (dia_tma.cuh)
  bool dia_on = false;
  int dia_K = 0;
  int64_t dia_off[4] = {};
  double *dia_U[4] = {};
  int64_t dia_lo0 = 0, dia_nlo = 0, dia_hi0 = 0, dia_nhi = 0;

  // closed-form assembly of regulargrid-ordered problems straight into the diagonal copy (box.cuh)
  int box_request = 0;          // 0 automatic, 1 never (FVB_BOX=0): always the general adjacency/CSR path
  bool box = false;             // current problem assembled by the box path: no adjacency, CSR built lazily
  bool box_implicit = false;    // ... from node conductivities and the grid spacing only (no face arrays at all)
  fvb::BoxDesc *boxd = nullptr;
  uint8_t *boxmask = nullptr;   // [nf_local] which of the six off-diagonal entries each row stores
  double *nodek = nullptr;      // implicit: node conductivities, nodes nodek_lo .. (planes p_lo-1 .. p_hi+1)
  int64_t nodek_n = 0, nodek_ofs = 0;  // entries held; index of local node 0
  int box_logmean = 0;
  int64_t *d_dsorted = nullptr; // sorted Dirichlet table of the whole problem (slab ranks; retained in box mode)
  int32_t *d_dsorted_slot = nullptr;
  int64_t nd_sorted = 0;

  // symmetric Jacobi scaling of that copy (dia.cuh: k_dia_scale; pcg.cuh: SC kernels)
  int scale_request = 0;        // 0 auto, 1 never
  int scale_state = 0;          // 0 not built for the current values, 1 ready, 2 not applicable
  double *dia_S[4] = {};        // D^-1/2 U_k D^-1/2, same layout as dia_U
  double *sinv = nullptr;       // [nf_local] diag^-1/2
  bool last_solve_scaled = false;

  // adjoint gradient accumulators (fvb_gradient_*)
  int logk = 0;
  int32_t *g_e1 = nullptr, *g_e2 = nullptr;   // per-face endpoints: local row >= 0, -1-slot, or INT_MIN (off-rank)
  double *g_face = nullptr, *g_dh = nullptr, *g_src = nullptr;

  // preconditioner: 0 Jacobi, 1 aggregation multigrid (mg.cuh)
  int precond_request = 0;
  int mg_nu = 2;
  double mg_omega = 0.8, mg_oc = 1.5;
  fvb::MgState *mg = nullptr;
  fvb::AmgState *amg = nullptr;   // aggregation AMG on CSR rows for matrices mg.cuh does not cover (amg.cuh)

  // in-situ SpMV launch timing (fvb_set_profiling)
  int prof_stride = 0;
  int64_t prof_seen = 0;
  int prof_count = 0;
  cudaEvent_t prof_ev[128] = {};

  fvb_timings tm = {};
};
