// fvb200.cu -- the C ABI of libfvb200.so (include/fvb200.h) over the sm_100a kernels in
// assemble.cuh / spmv.cuh / pcg.cuh.  Host code here only allocates, orders launches and
// moves data; there is no CPU implementation of any part of the path.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdlib>

#include "amg.cuh"
#include "arena.h"
#include "assemble.cuh"
#include "box.cuh"
#include "common.cuh"
#include "coop.cuh"
#include "dia.cuh"
#include "dia_tma.cuh"
#include "grid.cuh"
#include "host_util.h"
#include "nccl_dyn.h"
#include "pcg.cuh"
#include "peer.cuh"
#include "scan.cuh"
#include "spmv.cuh"

namespace fvb {
thread_local std::string g_last_error;
int set_error(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}
}  // namespace fvb

namespace fvb {
struct PeerState {
  bool active = false;
  PeerMail *mail = nullptr;             // this rank's mailbox (cudaMalloc, exported)
  void *opened_u[kMaxRanks] = {};       // cudaIpcOpenMemHandle results (to close)
  void *opened_mail[kMaxRanks] = {};
  PeerTable tab = {};
  PeerTable *d_tab = nullptr;           // device copy of tab for the in-kernel all-reduce (peer_base.cuh)
  bool fused = true;                    // reducing kernels finish their sums across the ranks themselves
  HaloPlanDev push = {}, wait = {};
  unsigned long long red_seq = 0, halo_seq = 0;
  std::vector<uint8_t> last_blobs;      // what the open mappings correspond to
};
struct AmgState {
  bool ready = false;
  int nlev = 0;
  AmgLevel lev[kAmgMaxLevels];
  std::vector<void *> owned;  // every device array of the hierarchy (freed together)
};
struct MgState {
  bool ready = false;
  int nlev = 0;
  MgLevel lev[kMgMaxLevels];
  double *own[kMgMaxLevels][8] = {};  // diag, up0..2, x, r, t, lo_c owned by level l (level 0 owns only x, t)
  bool dist = false;                  // V-cycle spans the ranks (halo planes exchanged per level)
  int lo_rank = -1, hi_rank = -1;
};
}  // namespace fvb

using namespace fvb;

#define FVB_NCCL(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t r__ = (expr);                                                                  \
    if (r__ != ncclSuccess)                                                                     \
      return set_error(FVB_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(r__));   \
  } while (0)

namespace {

// Device memory of a handle comes from its arena (arena.h: best-fit blocks inside a few big cudaMalloc
// chunks, so a repeated assemble -> solve sequence reuses the very same blocks without any driver call).
// FVB_ARENA=0 selects the stream-ordered pool instead (cudaMallocAsync, release threshold = keep
// everything), whose defragmentation stalls of up to a second per step are the reason the arena exists.
void *arena_chunk_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void arena_chunk_free(void *p) { cudaFree(p); }

template <typename T>
int dalloc(fvb_handle h, T **p, int64_t n) {
  *p = nullptr;
  size_t bytes = sizeof(T) * (size_t)std::max<int64_t>(n, 1);
  if (h->arena) {
    *p = static_cast<T *>(h->arena->alloc(bytes));
    if (!*p)
      return set_error(FVB_ERR_OOM, "out of device memory: " + std::to_string(bytes) + " bytes requested, arena holds " +
                                        std::to_string(h->arena->reserved()) + " (" + std::to_string(h->arena->in_use()) + " in use)");
    return FVB_OK;
  }
  cudaError_t e = cudaMallocAsync((void **)p, bytes, h->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(FVB_ERR_OOM, "cudaMallocAsync of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
  }
  return FVB_OK;
}
template <typename T>
void dfree(fvb_handle h, T *&p) {
  if (p) {
    if (h->arena) h->arena->free(p);
    else cudaFreeAsync(p, h->stream);
  }
  p = nullptr;
}

// every copy goes through the handle's own stream: pool memory is ordered on it, and the handle's
// stream is non-blocking (it does not synchronise with the legacy default stream)
cudaError_t memcpy_sync(cudaStream_t st, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}

int ensure_csr(fvb_handle h);  // box problems: CSR image on first demand (defined with the box code below)
void amg_free(fvb_handle h);   // aggregation-AMG hierarchy (defined with the multigrid code below)

int grid_for(int64_t n) { return std::max(1, cdiv(n, kBlock)); }
int vgrid(fvb_handle h, int64_t n) { return std::max(1, std::min(cdiv(n, kBlock), h->num_sms * 8)); }

void free_problem(fvb_handle h) {
  dfree(h, h->nodemap); dfree(h, h->row2node); dfree(h, h->sources); dfree(h, h->dheads); dfree(h, h->aol);
  dfree(h, h->meta); dfree(h, h->cface); dfree(h, h->adjptr); dfree(h, h->adj_face); dfree(h, h->adj_col);
  dfree(h, h->halo_glob); dfree(h, h->rowptr); dfree(h, h->colidx); dfree(h, h->vals); dfree(h, h->b); dfree(h, h->diag);
  dfree(h, h->x); dfree(h, h->r); dfree(h, h->c); dfree(h, h->dinv); dfree(h, h->rhs); dfree(h, h->Dvec);
  if (h->nranks == 1) { dfree(h, h->u); h->u_cap = 0; }  // multi-rank: u is IPC-exported, kept and reused
  if (h->peer) h->peer->active = false;                  // plan changes: peers must be re-imported
  for (auto &s : h->slots) dfree(h, s);
  dfree(h, h->partials); dfree(h, h->hist); dfree(h, h->xio); dfree(h, h->yio);
  dfree(h, h->send_rows); dfree(h, h->sendbuf);
  dfree(h, h->g_e1); dfree(h, h->g_e2); dfree(h, h->g_face); dfree(h, h->g_dh); dfree(h, h->g_src);
  for (auto &u : h->dia_U) dfree(h, u);
  for (auto &u : h->dia_S) dfree(h, u);
  dfree(h, h->sinv);
  dfree(h, h->boxmask); dfree(h, h->nodek); dfree(h, h->d_dsorted); dfree(h, h->d_dsorted_slot);
  delete h->boxd;
  h->boxd = nullptr;
  h->box = false; h->box_implicit = false; h->nodek_n = 0; h->nd_sorted = 0;
  h->scale_state = 0;
  h->dia_on = false;
  h->dia_K = 0;
  if (h->mg) {
    for (auto &lv : h->mg->own) for (auto &p : lv) dfree(h, p);
    h->mg->ready = false;
    h->mg->nlev = 0;
  }
  amg_free(h);
  h->hist_cap = 0;
  h->assembled = false;
  h->halo_ready = false;
  h->peers.clear(); h->send_counts.clear(); h->recv_counts.clear(); h->send_first.clear(); h->send_contig = false;
  h->halo_host.clear();
  h->n_send = 0;
}

int check_handle(fvb_handle h, bool need_assembled) {
  if (!h) return set_error(FVB_ERR_BAD_INPUT, "null handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  if (need_assembled && !h->assembled) return set_error(FVB_ERR_STATE, "fvb_assemble has not succeeded on this handle");
  return FVB_OK;
}

int ensure_workspace(fvb_handle h) {
  const int64_t n = h->nf_local;
  if (!h->x) {
    FVB_TRY(dalloc(h, &h->x, n)); FVB_TRY(dalloc(h, &h->r, n)); FVB_TRY(dalloc(h, &h->c, n));
    // dinv (unscaled recurrence, multigrid PCG) and rhs (transient step) are taken on first use: the scaled
    // steady solve needs neither, and at 1024^3 on one GPU each is 8.6 GB
    if (h->nranks == 1) {
      FVB_TRY(dalloc(h, &h->u, n + h->n_halo));
    } else if (h->u_cap < std::max<int64_t>(n + h->n_halo, 1)) {  // (a rank without rows or halo still exports a vector)
      if (h->u) FVB_CUDA(cudaFree(h->u));
      h->u = nullptr;
      h->u_cap = std::max<int64_t>(n + h->n_halo, 1);
      FVB_CUDA(cudaMalloc((void **)&h->u, sizeof(double) * (size_t)h->u_cap));
    }
    FVB_TRY(dalloc(h, &h->partials, 2 * (int64_t)std::max(cdiv(n, kSpmvRows), h->num_sms * 8) + 2));
    FVB_CUDA(cudaMemsetAsync(h->u, 0, sizeof(double) * (size_t)std::max<int64_t>(n + h->n_halo, 1), h->stream));
  }
  return FVB_OK;
}

int ensure_hist(fvb_handle h, int64_t cap) {
  cap = std::max<int64_t>(std::min<int64_t>(cap, 1 << 24), 1);
  if (cap > h->hist_cap) {
    dfree(h, h->hist);
    FVB_TRY(dalloc(h, &h->hist, cap));
    h->hist_cap = cap;
  }
  return FVB_OK;
}

// ---- multi-rank plumbing -----------------------------------------------------------------
// pushed: the producer of `vec` already stored the boundary rows into the neighbours' halo slots under a freshly
//         drawn sequence number (fused_push); defer: the consumer kernel waits for the neighbours' flags itself
//         (dia_tma.cuh: FusedWait) -- filled here instead of launching k_halo_wait.
int halo_exchange(fvb_handle h, double *vec, bool pushed = false, FusedWait *defer = nullptr) {
  if (h->nranks == 1) return FVB_OK;
  if (!h->halo_ready) return set_error(FVB_ERR_STATE, "fvb_set_halo_plan has not been called on this rank");
  if (h->peer && h->peer->active && vec == h->u) {
    // NVLink peer stores + flags (peer.cuh)
    PeerState &P = *h->peer;
    const unsigned long long seq = pushed ? P.halo_seq : ++P.halo_seq;
    if (!pushed && P.push.npeers > 0) {
      const int g = std::max(1, std::min(cdiv(h->n_send, kBlock), h->num_sms * 2));
      k_halo_push<<<g, kBlock, 0, h->stream>>>(P.tab, P.push, h->send_rows, vec, seq, h->ticket + 1);
      h->tm.kernel_launches++;
    }
    if (P.wait.npeers > 0) {
      if (defer && P.wait.npeers <= 2) {
        defer->n = P.wait.npeers;
        for (int k = 0; k < P.wait.npeers; ++k) defer->flag[k] = &P.mail->hseq[P.wait.peer[k]];
        defer->seq = seq;
        defer->error = &P.mail->error;
      } else {
        k_halo_wait<<<1, 32, 0, h->stream>>>(P.mail, P.wait, seq, h->scal, pushed ? 1 : 0);
        h->tm.kernel_launches++;
      }
    }
    return FVB_OK;
  }
  if (pushed) return set_error(FVB_ERR_STATE, "internal: fused halo push without the peer-memory path");
  if (h->n_send > 0) {
    k_pack<<<grid_for(h->n_send), kBlock, 0, h->stream>>>(h->n_send, h->send_rows, vec, h->sendbuf);
    h->tm.kernel_launches++;
  }
  NcclApi &N = nccl();
  FVB_NCCL(N.GroupStart());
  int64_t so = 0, ro = 0;
  for (size_t p = 0; p < h->peers.size(); ++p) {
    if (h->send_counts[p] > 0)
      FVB_NCCL(N.Send(h->sendbuf + so, (size_t)h->send_counts[p], ncclDouble, h->peers[p], h->comm->comm, h->stream));
    if (h->recv_counts[p] > 0)
      FVB_NCCL(N.Recv(vec + h->nf_local + ro, (size_t)h->recv_counts[p], ncclDouble, h->peers[p], h->comm->comm, h->stream));
    so += h->send_counts[p];
    ro += h->recv_counts[p];
  }
  FVB_NCCL(N.GroupEnd());
  return FVB_OK;
}

// How a reducing kernel of the Jacobi-PCG closes its sums (pcg.cuh: finalize_mode): 1 single rank,
// 2 in-kernel all-reduce over peer memory (the sequence number is drawn here), 0 separate all-reduce.
int red_mode(fvb_handle h, PeerRed *pr) {
  *pr = PeerRed{nullptr, 0ull};
  if (h->nranks == 1) return 1;
  if (h->peer && h->peer->active && h->peer->fused && h->peer->d_tab) {
    pr->tab = h->peer->d_tab;
    pr->seq = ++h->peer->red_seq;
    return 2;
  }
  return 0;
}

// Descriptor for pushing the boundary rows of h->u from the kernel that writes them (pcg.cuh: k_update_u): possible
// when the peer-memory path is up and every neighbour receives one contiguous run of rows (slab partitions).
// Draws the halo sequence number; the product that follows must be launched with pushed = true.
FusedPush fused_push(fvb_handle h) {
  FusedPush fp = {};
  if (h->nranks == 1 || !h->halo_ready || !(h->peer && h->peer->active) || !h->send_contig) return fp;
  PeerState &P = *h->peer;
  if (P.push.npeers < 1 || P.push.npeers > 2 || !h->fused_halo) return fp;
  int k = 0;
  for (size_t p = 0; p < h->peers.size(); ++p) {
    if (h->send_counts[p] <= 0) continue;
    const int peer = h->peers[p];
    fp.dst[k] = P.tab.u[peer] + P.push.dst_off[k];
    fp.begin[k] = h->send_first[p];
    fp.count[k] = h->send_counts[p];
    fp.flag[k] = &P.tab.mail[peer]->hseq[h->rank];
    ++k;
  }
  fp.npeers = k;
  fp.seq = ++P.halo_seq;
  return fp;
}

// Sum scal->red[0..count) over the ranks and advance the recurrence (mode: FIN_*).
int allreduce_fin(fvb_handle h, int count, int mode) {
  if (h->nranks == 1) return FVB_OK;
  double *red = h->scal->red;
  if (h->peer && h->peer->active) {
    PeerState &P = *h->peer;
    k_allreduce_fin<<<1, 32, 0, h->stream>>>(P.tab, ++P.red_seq, red, count, mode, h->scal, h->hist);
    h->tm.kernel_launches++;
    return FVB_OK;
  }
  FVB_NCCL(nccl().AllReduce(red, red, (size_t)count, ncclDouble, ncclSum, h->comm->comm, h->stream));
  if (mode == FIN_INIT) k_fin_init<<<1, 1, 0, h->stream>>>(h->scal);
  else if (mode == FIN_UC) k_fin_uc<<<1, 1, 0, h->stream>>>(h->scal);
  else if (mode == FIN_ITER) k_fin_iter<<<1, 1, 0, h->stream>>>(h->scal, h->hist);
  else if (mode == FIN_RZ) k_fin_rz<<<1, 1, 0, h->stream>>>(h->scal);
  else if (mode == FIN_R) k_fin_r<<<1, 1, 0, h->stream>>>(h->scal, h->hist);
  if (mode != FIN_NONE) h->tm.kernel_launches++;
  return FVB_OK;
}

// The TMA-staged diagonal kernel (dia_tma.cuh) needs more than 48 KB of dynamic shared memory: opt in
// once per instantiation and device (again only if a larger layout shows up).
template <bool DOT, int K, bool UNIT>
int launch_dia_tma(fvb_handle h, int n, const DiaDesc &D, const DiaTmaLayout &L, const double *vec, double *out,
                   double sigma, int fin, PeerRed pr, FusedWait fw) {
  static int opted[64] = {};
  const int smem = (int)dia_tma_smem_bytes(L);
  const int dev = h->device & 63;
  if (opted[dev] < smem) {
    FVB_CUDA(cudaFuncSetAttribute(k_spmv_dia_tma<DOT, K, UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    opted[dev] = smem;
  }
  const int ntiles = cdiv(n, kDiaTmaTile);
  const int grid = std::max(1, std::min(ntiles, h->num_sms * kDiaTmaCtasPerSm));
  k_spmv_dia_tma<DOT, K, UNIT><<<grid, kDiaTmaThreads, smem, h->stream>>>(n, D, L, vec, out, h->Dvec, sigma, h->partials,
                                                                        h->ticket, h->scal, fin, pr, fw);
  return FVB_OK;
}

// Which diagonal kernel serves this launch: the TMA pipeline when (almost) all tiles are interior
// tiles and every slice is 16-byte aligned, else the per-thread-load kernel of dia.cuh.
bool use_dia_tma(fvb_handle h, const double *vec, const DiaTmaLayout &L, bool scaled) {
  if (h->fmt_request == 2) return false;
  const int64_t n = h->nf_local, T = kDiaTmaTile;
  auto misaligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if ((vec && misaligned(vec)) || misaligned(h->diag)) return false;
  for (int k = 0; k < h->dia_K; ++k)
    if (misaligned(scaled ? h->dia_S[k] : h->dia_U[k])) return false;
  if (L.stages == 0) return false;  // one stage pair does not fit into shared memory
  if (h->fmt_request == 3) return true;
  const int64_t ntiles = (n + T - 1) / T;
  const int64_t first = (L.reach + T - 1) / T, last = (n - T - L.reach) >= 0 ? (n - T - L.reach) / T : -1;
  const int64_t interior = std::max<int64_t>(0, last - first + 1);
  return interior * 10 >= ntiles * 8 && ntiles >= 4 * h->num_sms;
}

// c = (A + sigma*D) vec, vec has halo room.  With dot: also u.Au into scal (CG use).
// scaled: c = A^ vec with A^ = D^-1/2 A D^-1/2 (build_scaled must have succeeded; sigma = 0).
// fuse: let the kernel finish u.Au across the ranks itself when the peer-memory path allows (the caller
// must then skip allreduce_fin: *fin_out reports the finalize mode used).
int launch_spmv(fvb_handle h, double *vec, double *out, double sigma, bool dot, bool scaled = false, bool fuse = false,
                int *fin_out = nullptr, bool pushed = false) {
  if (h->box && h->fmt_request == 1) FVB_TRY(ensure_csr(h));  // forced CSR on a problem assembled without one
  const int n = (int)h->nf_local;
  // which kernel serves this product decides how the halo arrives: the TMA diagonal pipeline waits for the
  // neighbours' planes itself, right before its edge tiles; every other kernel gets a k_halo_wait launch
  const bool use_dia = h->dia_on && h->fmt_request != 1 && n > 0;
  DiaDesc D = {};
  DiaTmaLayout L = {};
  bool tma = false;
  if (use_dia) {
    D.K = h->dia_K;
    for (int k = 0; k < kDiaMaxOff; ++k) { D.off[k] = h->dia_off[k]; D.U[k] = scaled ? h->dia_S[k] : h->dia_U[k]; }
    D.diag = h->diag; D.row_start = h->row_start; D.nf = h->nf_local;
    D.lo0 = h->dia_lo0; D.nlo = h->dia_nlo; D.hi0 = h->dia_hi0; D.nhi = h->dia_nhi;
    L = dia_tma_layout(D.K, D.off, scaled);
    tma = use_dia_tma(h, vec, L, scaled);
  }
  FusedWait fw = {};
  FVB_TRY(halo_exchange(h, vec, pushed, (tma && h->fused_halo) ? &fw : nullptr));
  PeerRed pr = {nullptr, 0ull};
  int fin = h->nranks == 1 ? 1 : 0;
  if (dot && fuse) fin = red_mode(h, &pr);
  if (fin_out) *fin_out = fin;
  // A rank that owns no free row (e.g. a slab made of a Dirichlet plane only) still takes part in the
  // reduction of u.Au with a zero contribution: the kernel runs with one CTA and no tiles.
  if (n == 0 && !dot) {
    if (h->nranks > 1 && h->peer && h->peer->active && vec == h->u) FVB_TRY(allreduce_fin(h, 0, FIN_NONE));  // see below
    return FVB_OK;
  }
  const int grid = std::max(1, std::min(cdiv(n, kSpmvRows), h->num_sms * kSpmvCtasPerSm));
  const size_t smem = sizeof(SpmvSmem);
  int sample = -1;
  if (dot && h->prof_stride > 0 && h->prof_count < 64 && (h->prof_seen++ % h->prof_stride) == 0) {
    sample = h->prof_count++;
    cudaEventRecord(h->prof_ev[2 * sample], h->stream);
  }
  if (use_dia) {
    const int dg = std::min(cdiv(n, kBlock * kDiaRowsPerThread), h->num_sms * kDiaCtasPerSm);
#define FVB_DIA_LAUNCH(DOTV, KV, UV)                                                                                \
  if (tma) FVB_TRY((launch_dia_tma<DOTV, KV, UV>(h, n, D, L, vec, out, sigma, fin, pr, fw)));                         \
  else k_spmv_dia<DOTV, KV, UV><<<dg, kBlock, 0, h->stream>>>(n, D, vec, out, h->Dvec, sigma, h->partials, h->ticket, \
                                                              h->scal, fin, pr)
#define FVB_DIA_K(DOTV, UV)                                   \
  switch (D.K) {                                              \
    case 1: FVB_DIA_LAUNCH(DOTV, 1, UV); break;               \
    case 2: FVB_DIA_LAUNCH(DOTV, 2, UV); break;               \
    case 3: FVB_DIA_LAUNCH(DOTV, 3, UV); break;               \
    default: FVB_DIA_LAUNCH(DOTV, 4, UV); break;              \
  }
    if (scaled) { if (dot) { FVB_DIA_K(true, true) } else { FVB_DIA_K(false, true) } }
    else if (dot) { FVB_DIA_K(true, false) } else { FVB_DIA_K(false, false) }
    h->last_dia_tma = tma;
#undef FVB_DIA_K
#undef FVB_DIA_LAUNCH
  } else if (scaled)
    return set_error(FVB_ERR_STATE, "scaled SpMV requested without the diagonal format");
  else if (h->box && !h->rowptr)
    return set_error(FVB_ERR_STATE, "internal: CSR product requested before ensure_csr");
  else if (dot)
    k_spmv<true><<<grid, kSpmvThreads, smem, h->stream>>>(n, h->rowptr, h->colidx, h->vals, vec, out, h->Dvec, sigma,
                                                     h->partials, h->ticket, h->scal, fin, pr);
  else
    k_spmv<false><<<grid, kSpmvThreads, smem, h->stream>>>(n, h->rowptr, h->colidx, h->vals, vec, out, h->Dvec, sigma,
                                                      h->partials, h->ticket, h->scal, fin, pr);
  if (sample >= 0) cudaEventRecord(h->prof_ev[2 * sample + 1], h->stream);
  h->tm.kernel_launches++;
  // Peer-memory halo slots have no consumer->producer acknowledgement of their own: inside the CG loop the
  // cross-rank reduction that follows every product orders "all ranks have read halo k" before "anybody pushes
  // halo k+1".  A product that is NOT followed by a reduction (fvb_spmv, fvb_time_spmv, the warm-start residual)
  // gets that ordering from an empty all-reduce, i.e. a barrier across the ranks on the device.
  if (!dot && h->nranks > 1 && h->peer && h->peer->active && vec == h->u) FVB_TRY(allreduce_fin(h, 0, FIN_NONE));
  return FVB_OK;
}

// Try to mirror A onto <= kDiaMaxOff symmetric diagonals (dia.cuh).  structure=true: detect the
// pattern and allocate; false: only refresh the values (fvb_update_values).
int build_dia(fvb_handle h, bool structure) {
  cudaStream_t st = h->stream;
  const int n = (int)h->nf_local;
  h->scale_state = 0;  // the values change: the scaled copy (build_scaled) is stale
  if (structure) {
    for (auto &u : h->dia_U) dfree(h, u);
    for (auto &u : h->dia_S) dfree(h, u);
    dfree(h, h->sinv);
    h->dia_on = false;
    h->dia_K = 0;
    if (n < 2 || h->nnz == 0) return FVB_OK;
    FVB_CUDA(cudaStreamSynchronize(st));  // the sample reads below go through the blocking default stream
    // candidate offsets: the union over a few sample rows (first, quartiles, last)
    std::vector<int64_t> offs;
    std::vector<int> rp(2), cols;
    const int samples[5] = {0, n / 4, n / 2, (int)((int64_t)3 * n / 4), n - 1};
    for (int sr : samples) {
      FVB_CUDA(memcpy_sync(h->stream, rp.data(), h->rowptr + sr, 2 * sizeof(int), cudaMemcpyDeviceToHost));
      int len = rp[1] - rp[0];
      if (len > 64) return FVB_OK;
      cols.resize((size_t)std::max(len, 1));
      if (len) FVB_CUDA(memcpy_sync(h->stream, cols.data(), h->colidx + rp[0], sizeof(int) * (size_t)len, cudaMemcpyDeviceToHost));
      for (int k = 0; k < len; ++k) {
        int64_t g = cols[(size_t)k] < n ? h->row_start + cols[(size_t)k] : h->halo_host[(size_t)(cols[(size_t)k] - n)];
        int64_t d = g - (h->row_start + sr);
        if (d < 0) d = -d;
        if (d != 0 && std::find(offs.begin(), offs.end(), d) == offs.end()) offs.push_back(d);
      }
    }
    std::sort(offs.begin(), offs.end());
    const int K = (int)offs.size();
    if (K == 0 || K > kDiaMaxOff) return FVB_OK;
    if ((double)h->nnz < 0.5 * (2.0 * K + 1.0) * n) return FVB_OK;  // mostly empty diagonals: CSR is smaller
    // halo columns must form at most one contiguous run below and one above the owned rows
    int64_t lo0 = 0, nlo = 0, hi0 = 0, nhi = 0;
    for (int64_t g : h->halo_host) {
      if (g < h->row_start) {
        if (nlo == 0) lo0 = g;
        if (g != lo0 + nlo) return FVB_OK;
        ++nlo;
      } else {
        if (nhi == 0) hi0 = g;
        if (g != hi0 + nhi) return FVB_OK;
        ++nhi;
      }
    }
    int64_t o[4] = {0, 0, 0, 0};
    for (int k = 0; k < K; ++k) o[k] = offs[(size_t)k];
    int *d_flag = nullptr;
    FVB_TRY(dalloc(h, &d_flag, 1));
    cudaMemsetAsync(d_flag, 0, sizeof(int), st);
    k_dia_check<<<grid_for(n), kBlock, 0, st>>>(n, h->rowptr, h->colidx, n, h->row_start, h->halo_glob, K, o[0], o[1],
                                                o[2], o[3], d_flag);
    h->tm.kernel_launches++;
    int flag = 1;
    cudaError_t e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dfree(h, d_flag);
    if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
    if (flag) return FVB_OK;
    for (int k = 0; k < K; ++k) {
      if (dalloc(h, &h->dia_U[k], (int64_t)n + o[k]) != FVB_OK) {  // not enough memory: stay on CSR
        for (auto &u : h->dia_U) dfree(h, u);
        return FVB_OK;
      }
      h->dia_off[k] = o[k];
    }
    h->dia_K = K;
    h->dia_lo0 = lo0; h->dia_nlo = nlo; h->dia_hi0 = hi0; h->dia_nhi = nhi;
    h->dia_on = true;
    // Room for the Jacobi-scaled copy (build_scaled) is taken here, in the same place of every
    // assembly, so that repeated assemble/solve cycles reach a steady allocation pattern in the
    // stream-ordered pool; if it does not fit, build_scaled tries again or stays unscaled.
    if (h->scale_request != 1) {
      bool ok = dalloc(h, &h->sinv, n) == FVB_OK;
      for (int k = 0; k < K && ok; ++k) ok = dalloc(h, &h->dia_S[k], (int64_t)n + o[k]) == FVB_OK;
      if (!ok) {
        for (auto &u : h->dia_S) dfree(h, u);
        dfree(h, h->sinv);
      }
    }
  }
  if (!h->dia_on) return FVB_OK;
  for (int k = 0; k < h->dia_K; ++k)
    FVB_CUDA(cudaMemsetAsync(h->dia_U[k], 0, sizeof(double) * (size_t)(n + h->dia_off[k]), st));
  k_dia_fill<<<grid_for(n), kBlock, 0, st>>>(n, h->rowptr, h->colidx, h->vals, n, h->row_start, h->halo_glob, h->dia_K,
                                             h->dia_off[0], h->dia_off[1], h->dia_off[2], h->dia_off[3], h->dia_U[0],
                                             h->dia_U[1], h->dia_U[2], h->dia_U[3]);
  h->tm.kernel_launches++;
  return FVB_OK;
}

// Symmetric Jacobi scaling of the diagonal copy: S_k = D^-1/2 U_k D^-1/2, sinv = diag^-1/2 (dia.cuh).
// Built lazily by the first cold-started steady Jacobi solve after the values changed.  Uses h->u as
// scratch (its halo slots receive the neighbours' s through the ordinary halo exchange).  All ranks
// must run the same recurrence, so the outcome is agreed over the communicator: scale_state = 1
// only if every rank has the diagonal format and a strictly positive diagonal.
int build_scaled(fvb_handle h) {
  if (h->scale_state != 0) return FVB_OK;
  cudaStream_t st = h->stream;
  const int64_t n = h->nf_local;
  int mine = (h->dia_on && n > 0) ? 1 : 0;
  int *d_flag = nullptr;
  FVB_TRY(dalloc(h, &d_flag, 1));
  FVB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  if (mine) {
    bool ok = true;
    if (!h->sinv) ok = dalloc(h, &h->sinv, n) == FVB_OK;
    for (int k = 0; k < h->dia_K && ok; ++k)
      if (!h->dia_S[k]) ok = dalloc(h, &h->dia_S[k], n + h->dia_off[k]) == FVB_OK;
    if (!ok) {  // not enough memory for the second copy: stay on the unscaled recurrence
      for (auto &u : h->dia_S) dfree(h, u);
      dfree(h, h->sinv);
      mine = 0;
    }
  }
  if (mine) {
    k_make_sinv<<<vgrid(h, n), kBlock, 0, st>>>(n, h->diag, h->u, h->sinv, d_flag);
    h->tm.kernel_launches++;
  }
  int flag = 1;
  if (h->nranks > 1 && !(h->comm && h->comm->comm)) mine = 0;
  else if (h->nranks > 1) {
    // flag <- max over ranks of (bad diagonal | no diagonal format)
    if (!mine) { flag = 1; FVB_CUDA(cudaMemcpyAsync(d_flag, &flag, sizeof(int), cudaMemcpyHostToDevice, st)); }
    FVB_NCCL(nccl().AllReduce(d_flag, d_flag, 1, ncclInt, ncclMax, h->comm->comm, st));
  }
  cudaError_t e = memcpy_sync(st, &flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
  dfree(h, d_flag);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  if (!mine || flag) { h->scale_state = 2; return FVB_OK; }
  FVB_TRY(halo_exchange(h, h->u));
  DiaDesc D;
  D.K = h->dia_K;
  for (int k = 0; k < kDiaMaxOff; ++k) { D.off[k] = h->dia_off[k]; D.U[k] = h->dia_U[k]; }
  D.diag = h->diag; D.row_start = h->row_start; D.nf = h->nf_local;
  D.lo0 = h->dia_lo0; D.nlo = h->dia_nlo; D.hi0 = h->dia_hi0; D.nhi = h->dia_nhi;
  const int g = vgrid(h, n);
  switch (D.K) {
    case 1: k_dia_scale<1><<<g, kBlock, 0, st>>>((int)n, D, h->u, h->dia_S[0], h->dia_S[1], h->dia_S[2], h->dia_S[3]); break;
    case 2: k_dia_scale<2><<<g, kBlock, 0, st>>>((int)n, D, h->u, h->dia_S[0], h->dia_S[1], h->dia_S[2], h->dia_S[3]); break;
    case 3: k_dia_scale<3><<<g, kBlock, 0, st>>>((int)n, D, h->u, h->dia_S[0], h->dia_S[1], h->dia_S[2], h->dia_S[3]); break;
    default: k_dia_scale<4><<<g, kBlock, 0, st>>>((int)n, D, h->u, h->dia_S[0], h->dia_S[1], h->dia_S[2], h->dia_S[3]); break;
  }
  h->tm.kernel_launches++;
  h->scale_state = 1;
  return FVB_OK;
}

// ---- closed-form assembly of regulargrid-ordered problems (box.cuh) -------------------------------------------------
// Number of faces regulargrid lists for the x-planes p_lo..p_hi (plus the +x faces of plane p_lo-1).
int64_t box_face_count(const GridDesc &G) {
  const int64_t plane = G.n2 * G.n3, pfull = plane + (G.n2 - 1) * G.n3 + G.n2 * (G.n3 - 1);
  int64_t F = G.e_lo < G.p_lo ? plane : 0;
  const int64_t full = std::max<int64_t>(0, std::min<int64_t>(G.p_hi, G.n1 - 1) - G.p_lo + 1);
  F += full * pfull;
  if (G.p_hi == G.n1) F += pfull - plane;
  return F;
}

// Rows of the current box problem: (re)compute U, diag, b, the entry masks and nnz from the retained inputs.
int box_fill(fvb_handle h) {
  cudaStream_t st = h->stream;
  const BoxDesc &B = *h->boxd;
  const int64_t n = h->nf_local;
  h->scale_state = 0;  // values change: the Jacobi-scaled copy is stale
  unsigned long long *d_nnz = nullptr;
  FVB_TRY(dalloc(h, &d_nnz, 1));
  FVB_CUDA(cudaMemsetAsync(d_nnz, 0, sizeof(unsigned long long), st));
  for (int k = 0; k < 3; ++k)  // lower entries without an owned partner row: zero unless the rank below owns it
    FVB_CUDA(cudaMemsetAsync(h->dia_U[k], 0, sizeof(double) * (size_t)h->dia_off[k], st));
  const int g = std::max(1, std::min(cdiv(B.n_own, kBlock), h->num_sms * 16));
  if (h->box_implicit) {
    FaceImplicit fc{h->nodek, h->nodek_ofs, h->box_logmean, h->logk};
    k_box_values<<<g, kBlock, 0, st>>>(B, fc, h->nodemap, h->sources, h->dheads, h->d_dsorted, h->d_dsorted_slot,
                                       h->nd_sorted, h->dia_U[0], h->dia_U[1], h->dia_U[2], h->diag, h->b, h->boxmask, d_nnz);
  } else {
    FaceFromArray fc{h->cface};
    k_box_values<<<g, kBlock, 0, st>>>(B, fc, h->nodemap, h->sources, h->dheads, h->d_dsorted, h->d_dsorted_slot,
                                       h->nd_sorted, h->dia_U[0], h->dia_U[1], h->dia_U[2], h->diag, h->b, h->boxmask, d_nnz);
  }
  h->tm.kernel_launches++;
  unsigned long long nnz = 0;
  cudaError_t e = memcpy_sync(st, &nnz, d_nnz, sizeof(nnz), cudaMemcpyDeviceToHost);
  dfree(h, d_nnz);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, std::string("box assembly: ") + cudaGetErrorString(e));
  h->nnz = (int64_t)nnz;
  (void)n;
  return FVB_OK;
}

// Plane just outside the owned range: 0 none, 1 entirely free, 2 entirely Dirichlet, -1 mixed.
// `nodes` = ascending distinct 0-based Dirichlet nodes of the whole problem.
int box_plane_kind(const std::vector<int64_t> &nodes, int64_t first, int64_t plane) {
  const int64_t c = std::lower_bound(nodes.begin(), nodes.end(), first + plane) - std::lower_bound(nodes.begin(), nodes.end(), first);
  return c == 0 ? 1 : (c == plane ? 2 : -1);
}

// Allocate the diagonal copy for a verified box problem, install halo runs and fill the rows.
// On FVB_OK with *ok == false nothing was kept (not enough memory for this path's arrays is an error, not a fallback).
int box_install(fvb_handle h, const BoxDesc &B) {
  cudaStream_t st = h->stream;
  const int64_t n = h->nf_local, plane = B.G.n2 * B.G.n3;
  const int64_t o[3] = {1, B.G.n3, plane};
  delete h->boxd;
  h->boxd = new BoxDesc(B);
  for (int k = 0; k < 3; ++k) {
    FVB_TRY(dalloc(h, &h->dia_U[k], n + o[k]));
    h->dia_off[k] = o[k];
  }
  h->dia_off[3] = 0;
  h->dia_K = 3;
  FVB_TRY(dalloc(h, &h->diag, n));
  FVB_TRY(dalloc(h, &h->b, n));
  FVB_TRY(dalloc(h, &h->boxmask, n));
  // halo: whole planes of the neighbouring ranks, contiguous in the global free numbering
  h->dia_lo0 = h->dia_hi0 = 0; h->dia_nlo = h->dia_nhi = 0;
  h->halo_host.clear();
  if (B.lo_kind == 1) { h->dia_lo0 = h->row_start - plane; h->dia_nlo = plane; }
  if (B.hi_kind == 1) { h->dia_hi0 = h->row_start + n; h->dia_nhi = plane; }
  h->n_halo = h->dia_nlo + h->dia_nhi;
  if (h->n_halo) {
    h->halo_host.resize((size_t)h->n_halo);
    for (int64_t i = 0; i < h->dia_nlo; ++i) h->halo_host[(size_t)i] = h->dia_lo0 + i;
    for (int64_t i = 0; i < h->dia_nhi; ++i) h->halo_host[(size_t)(h->dia_nlo + i)] = h->dia_hi0 + i;
    FVB_TRY(dalloc(h, &h->halo_glob, h->n_halo));
    FVB_CUDA(memcpy_sync(st, h->halo_glob, h->halo_host.data(), sizeof(int64_t) * (size_t)h->n_halo, cudaMemcpyHostToDevice));
  }
  h->box = true;
  h->dia_on = true;
  FVB_TRY(box_fill(h));
  if (h->scale_request != 1) {  // room for the Jacobi-scaled copy, taken at the same place of every assembly
    bool ok = dalloc(h, &h->sinv, n) == FVB_OK;
    for (int k = 0; k < 3 && ok; ++k) ok = dalloc(h, &h->dia_S[k], n + o[k]) == FVB_OK;
    if (!ok) {
      for (auto &u : h->dia_S) dfree(h, u);
      dfree(h, h->sinv);
    }
  }
  return FVB_OK;
}

// Does the face list of this handle's problem equal regulargrid's for the owned planes?  Reads a few faces to infer
// (n2*n3, n3), compares the whole list and the node->row shifts on the device.  `dnodes_sorted`: Dirichlet table
// of the whole problem (slab ranks only; empty for an unpartitioned problem, where no plane lies outside).
int box_detect(fvb_handle h, const int64_t *d_nb, const std::vector<int64_t> &dnodes_sorted, BoxDesc *out, bool *ok) {
  *ok = false;
  cudaStream_t st = h->stream;
  const int64_t N = h->n_nodes, lo = h->node_lo, hi = h->node_hi, F = h->n_faces;
  if (F < 3 || hi - lo < 8 || h->nf_local < 2) return FVB_OK;
  int64_t p0[2];
  FVB_CUDA(memcpy_sync(st, p0, d_nb, sizeof(p0), cudaMemcpyDeviceToHost));
  const int64_t plane = p0[1] - p0[0];
  if (plane < 4 || N % plane || lo % plane || hi % plane || N / plane < 2) return FVB_OK;
  GridDesc G = {};
  G.n1 = N / plane;
  G.p_lo = lo / plane + 1;
  G.p_hi = hi / plane;
  G.e_lo = G.p_lo > 1 ? G.p_lo - 1 : G.p_lo;
  const bool halo = G.e_lo < G.p_lo;
  if (p0[0] != (halo ? lo - plane + 1 : 1)) return FVB_OK;
  const int64_t idx0 = halo ? plane : 0;
  if (idx0 + 3 > F) return FVB_OK;
  int64_t f[6];
  FVB_CUDA(memcpy_sync(st, f, d_nb + 2 * idx0, sizeof(f), cudaMemcpyDeviceToHost));
  const int64_t lin = lo + 1;
  int k = 0;
  if (G.p_lo < G.n1) {
    if (f[0] != lin || f[1] != lin + plane) return FVB_OK;
    k = 1;
  }
  if (f[2 * k] != lin || f[2 * k + 2] != lin || f[2 * k + 3] != lin + 1) return FVB_OK;
  const int64_t n3 = f[2 * k + 1] - lin;
  if (n3 < 2 || n3 >= plane || plane % n3 || plane / n3 < 2) return FVB_OK;
  G.n3 = n3;
  G.n2 = plane / n3;
  if (box_face_count(G) != F) return FVB_OK;
  BoxDesc B = {};
  B.G = G;
  B.n_own = hi - lo;
  B.nd_owned = (hi - lo) - h->nf_local;
  B.lo_kind = B.hi_kind = 0;
  if (G.p_lo > 1) {
    B.lo_kind = box_plane_kind(dnodes_sorted, lo - plane, plane);
    if (B.lo_kind < 0 || (B.lo_kind == 1 && box_plane_kind(dnodes_sorted, lo, plane) != 1)) return FVB_OK;
  }
  if (G.p_hi < G.n1) {
    B.hi_kind = box_plane_kind(dnodes_sorted, hi, plane);
    if (B.hi_kind < 0 || (B.hi_kind == 1 && box_plane_kind(dnodes_sorted, hi - plane, plane) != 1)) return FVB_OK;
  }
  int *d_flag = nullptr;
  FVB_TRY(dalloc(h, &d_flag, 2));
  FVB_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), st));
  const int64_t nodes = (G.p_hi - G.e_lo + 1) * plane;
  k_box_check<<<std::max(1, std::min(cdiv(nodes, kBlock), h->num_sms * 16)), kBlock, 0, st>>>(
      B, reinterpret_cast<const longlong2 *>(d_nb), h->nodemap, d_flag);
  h->tm.kernel_launches++;
  int flag[2] = {1, 1};
  cudaError_t e = memcpy_sync(st, flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost);
  dfree(h, d_flag);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, std::string("box check: ") + cudaGetErrorString(e));
  if (flag[0] || flag[1]) return FVB_OK;
  *out = B;
  *ok = true;
  return FVB_OK;
}

// CSR arrays of a box problem, built on first demand (fvb_get_csr, forced CSR format).
int ensure_csr(fvb_handle h) {
  if (h->rowptr || !h->box) return FVB_OK;
  cudaStream_t st = h->stream;
  const int n = (int)h->nf_local;
  if (h->nnz >= INT_MAX - 1)
    return set_error(FVB_ERR_BAD_INPUT, "the CSR image of this rank's rows exceeds 32-bit local indices (" +
                                            std::to_string(h->nnz) + " entries); use more ranks to fetch A");
  int *d_cnt = nullptr, *d_scratch = nullptr;
  FVB_TRY(dalloc(h, &d_cnt, (int64_t)n + 2));
  FVB_TRY(dalloc(h, &d_scratch, scan_scratch_ints((int64_t)n + 1)));
  int st_code = FVB_OK;
  do {
    if ((st_code = dalloc(h, &h->rowptr, (int64_t)n + 1 + kRowptrPad)) != FVB_OK) break;
    if (n) k_box_csr_count<<<grid_for(n), kBlock, 0, st>>>(n, h->boxmask, d_cnt);
    exclusive_scan(d_cnt, n, h->rowptr, d_scratch, st, &h->tm.kernel_launches);
    k_fill_tail<<<1, kBlock, 0, st>>>(h->rowptr, n, kRowptrPad);
    if ((st_code = dalloc(h, &h->colidx, h->nnz + kCsrPad)) != FVB_OK) break;
    if ((st_code = dalloc(h, &h->vals, h->nnz + kCsrPad)) != FVB_OK) break;
    cudaMemsetAsync(h->colidx + h->nnz, 0, sizeof(int) * kCsrPad, st);
    cudaMemsetAsync(h->vals + h->nnz, 0, sizeof(double) * kCsrPad, st);
    if (n) k_box_csr_fill<<<grid_for(n), kBlock, 0, st>>>(n, h->boxmask, h->rowptr, h->dia_off[1], h->dia_off[2], h->dia_nlo,
                                                         h->dia_U[0], h->dia_U[1], h->dia_U[2], h->diag, h->colidx, h->vals);
    h->tm.kernel_launches += 3;
  } while (0);
  cudaError_t e = cudaStreamSynchronize(st);
  dfree(h, d_cnt); dfree(h, d_scratch);
  if (st_code != FVB_OK) { dfree(h, h->rowptr); dfree(h, h->colidx); dfree(h, h->vals); return st_code; }
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, std::string("lazy CSR build: ") + cudaGetErrorString(e));
  return FVB_OK;
}

// ---- aggregation multigrid (mg.cuh) ------------------------------------------------------------------------
int mg_grid(fvb_handle h, int64_t n) { return std::max(1, std::min(cdiv(n, kBlock), h->num_sms * kMgCtasPerSm)); }

// Build (structure=true) or refresh the Galerkin hierarchy from the diagonal copy of A.
// Returns FVB_OK with mg->ready=false when the matrix does not qualify.
int mg_setup(fvb_handle h, bool structure) {
  cudaStream_t st = h->stream;
  if (!h->mg) h->mg = new MgState();
  MgState &M = *h->mg;
  if (structure) {
    for (auto &lv : M.own) for (auto &p : lv) dfree(h, p);
    M.ready = false;
    M.nlev = 0;
    M.dist = false;
    // (a lambda: a rank whose matrix does not qualify must still reach the agreement below -- returning from
    //  mg_setup here would leave the other ranks alone in a collective)
    auto build = [&]() -> int {
    if (!h->dia_on || h->dia_K != 3 || h->dia_off[0] != 1) return FVB_OK;
    const int64_t n = h->nf_local, nz = h->dia_off[1], nynz = h->dia_off[2];
    if (nz < 2 || nynz % nz != 0 || n % nynz != 0 || nynz / nz < 2 || n / nynz < 1) return FVB_OK;
    MgLevel &L0 = M.lev[0];
    L0 = MgLevel{};
    L0.nz = (int)nz; L0.ny = (int)(nynz / nz); L0.nx = (int)(n / nynz); L0.n = n;
    L0.diag = h->diag;
    for (int k = 0; k < 3; ++k) L0.up[k] = h->dia_U[k] + h->dia_off[k];
    // slab neighbours: exactly one plane of halo columns below and/or above, owned by rank-1 / rank+1
    if (h->nranks > 1) {
      const bool lo = h->dia_nlo == nynz && h->rank > 0, hi = h->dia_nhi == nynz && h->rank + 1 < h->nranks;
      const bool clean = (h->dia_nlo == 0 || lo) && (h->dia_nhi == 0 || hi);
      if (clean && (lo || hi) && h->comm && h->comm->comm) {
        M.dist = true;
        M.lo_rank = lo ? h->rank - 1 : -1;
        M.hi_rank = hi ? h->rank + 1 : -1;
        L0.lo_c = lo ? h->dia_U[2] : nullptr;  // U_2[r], r < o_2: lower entries whose partner row is not owned
        L0.has_hi = hi ? 1 : 0;
      }
    }
    int *d_flag = nullptr;
    FVB_TRY(dalloc(h, &d_flag, 1));
    cudaMemsetAsync(d_flag, 0, sizeof(int), st);
    k_mg_check_box<<<mg_grid(h, n), kBlock, 0, st>>>(L0, d_flag);
    h->tm.kernel_launches++;
    int flag = 1;
    cudaError_t e = memcpy_sync(st, &flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
    dfree(h, d_flag);
    if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
    if (flag) return FVB_OK;
    FVB_TRY(dalloc(h, &M.own[0][4], n + 2 * nynz));  // x (= z, the preconditioned residual) + halo planes
    FVB_TRY(dalloc(h, &M.own[0][6], n + 2 * nynz));  // t
    FVB_CUDA(cudaMemsetAsync(M.own[0][4], 0, sizeof(double) * (size_t)(n + 2 * nynz), st));
    FVB_CUDA(cudaMemsetAsync(M.own[0][6], 0, sizeof(double) * (size_t)(n + 2 * nynz), st));
    L0.x = M.own[0][4]; L0.r = nullptr; L0.t = M.own[0][6];
    int l = 0;
    // Stop rule: single GPU -- local size; distributed -- the (y,z) plane only, so that every rank
    // builds the same number of levels whatever its number of x-planes.
    auto go_on = [&](const MgLevel &F) {
      return M.dist ? ((int64_t)F.ny * F.nz > 64 && F.ny >= 2 && F.nz >= 2) : (F.n > kMgCoarsest);
    };
    while (go_on(M.lev[l]) && l + 1 < kMgMaxLevels) {
      const MgLevel &F = M.lev[l];
      MgLevel &C = M.lev[l + 1];
      C = MgLevel{};
      C.nx = (F.nx + 1) / 2; C.ny = (F.ny + 1) / 2; C.nz = (F.nz + 1) / 2;
      C.n = (int64_t)C.nx * C.ny * C.nz;
      const int64_t pl = (int64_t)C.ny * C.nz;
      for (int a = 0; a < 4; ++a) FVB_TRY(dalloc(h, &M.own[l + 1][a], C.n));
      FVB_TRY(dalloc(h, &M.own[l + 1][4], C.n + 2 * pl));
      FVB_TRY(dalloc(h, &M.own[l + 1][5], C.n));
      FVB_TRY(dalloc(h, &M.own[l + 1][6], C.n + 2 * pl));
      FVB_CUDA(cudaMemsetAsync(M.own[l + 1][4], 0, sizeof(double) * (size_t)(C.n + 2 * pl), st));
      FVB_CUDA(cudaMemsetAsync(M.own[l + 1][6], 0, sizeof(double) * (size_t)(C.n + 2 * pl), st));
      if (F.lo_c) FVB_TRY(dalloc(h, &M.own[l + 1][7], pl));
      C.diag = M.own[l + 1][0];
      for (int k = 0; k < 3; ++k) C.up[k] = M.own[l + 1][1 + k];
      C.x = M.own[l + 1][4]; C.r = M.own[l + 1][5]; C.t = M.own[l + 1][6];
      C.lo_c = F.lo_c ? M.own[l + 1][7] : nullptr;
      C.has_hi = F.has_hi;
      ++l;
    }
    M.nlev = l + 1;
    return FVB_OK;
    };
    FVB_TRY(build());
  }
  if (structure && h->nranks > 1 && h->comm && h->comm->comm) {
    // All ranks must take the same path through the solver (its collectives are matched call by
    // call): agree on the weakest capability -- 0 none, 1 block-local V-cycle, 2 distributed.
    int mine = M.nlev == 0 ? 0 : (M.dist ? 2 : 1), all = 0;
    int *d_v = nullptr;
    FVB_TRY(dalloc(h, &d_v, 1));
    FVB_CUDA(cudaMemcpyAsync(d_v, &mine, sizeof(int), cudaMemcpyHostToDevice, st));
    FVB_NCCL(nccl().AllReduce(d_v, d_v, 1, ncclInt, ncclMin, h->comm->comm, st));
    cudaError_t e = memcpy_sync(st, &all, d_v, sizeof(int), cudaMemcpyDeviceToHost);
    dfree(h, d_v);
    if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
    if (all == 0) { M.nlev = 0; return FVB_OK; }
    if (all == 1 && M.dist) {
      // somebody cannot exchange planes: everybody preconditions with its own diagonal block
      M.dist = false;
      for (int l = 0; l < M.nlev; ++l) { M.lev[l].lo_c = nullptr; M.lev[l].has_hi = 0; }
    }
  }
  if (M.nlev == 0) return FVB_OK;
  for (int l = 0; l + 1 < M.nlev; ++l) {
    const MgLevel &C = M.lev[l + 1];
    k_mg_coarsen<<<grid_for(C.n), kBlock, 0, st>>>(M.lev[l], C.nx, C.ny, C.nz, M.own[l + 1][0], M.own[l + 1][1],
                                                   M.own[l + 1][2], M.own[l + 1][3], M.own[l + 1][7]);
    h->tm.kernel_launches++;
  }
  M.ready = true;
  return FVB_OK;
}

// Exchange the boundary planes of one level's iterate with the slab neighbours (NCCL; the planes are
// contiguous: first/last ny*nz entries out, halo slots [n, n+pl) and [n+pl, n+2pl) in).
int mg_halo(fvb_handle h, const MgLevel &L, double *v) {
  MgState &M = *h->mg;
  if (!M.dist) return FVB_OK;
  const size_t pl = (size_t)L.ny * L.nz;
  NcclApi &N = nccl();
  FVB_NCCL(N.GroupStart());
  if (M.lo_rank >= 0) {
    FVB_NCCL(N.Send(v, pl, ncclDouble, M.lo_rank, h->comm->comm, h->stream));
    FVB_NCCL(N.Recv(v + L.n, pl, ncclDouble, M.lo_rank, h->comm->comm, h->stream));
  }
  if (M.hi_rank >= 0) {
    FVB_NCCL(N.Send(v + L.n - pl, pl, ncclDouble, M.hi_rank, h->comm->comm, h->stream));
    FVB_NCCL(N.Recv(v + L.n + pl, pl, ncclDouble, M.hi_rank, h->comm->comm, h->stream));
  }
  FVB_NCCL(N.GroupEnd());
  return FVB_OK;
}

// z = M^-1 r : one V-cycle.  r is only read; the result lands in lev[0].x.
int mg_vcycle(fvb_handle h, const double *r) {
  MgState &M = *h->mg;
  cudaStream_t st = h->stream;
  const int nu = h->mg_nu;
  const double om = h->mg_omega, oc = h->mg_oc;
  M.lev[0].r = const_cast<double *>(r);
  // going down
  for (int l = 0; l + 1 < M.nlev; ++l) {
    MgLevel &L = M.lev[l];
    const int g = mg_grid(h, L.n);
    double *cur = L.t, *oth = L.x;  // 2*nu-1 swaps in total: start in t to finish in x
    k_mg_smooth0<<<g, kBlock, 0, st>>>(L, L.r, cur, om, h->scal);
    for (int s = 1; s < nu; ++s) {
      FVB_TRY(mg_halo(h, L, cur));
      k_mg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
      std::swap(cur, oth);
    }
    const MgLevel &C = M.lev[l + 1];
    FVB_TRY(mg_halo(h, L, cur));
    k_mg_restrict<<<mg_grid(h, C.n), kBlock, 0, st>>>(L, L.r, cur, C.nx, C.ny, C.nz, C.r, h->scal);
    h->tm.kernel_launches += nu + 1;
  }
  {
    MgLevel &L = M.lev[M.nlev - 1];
    if (!M.dist) {
      // (with a single level the matrix is tiny and the sweeps are the whole preconditioner)
      k_mg_coarse_solve<<<1, kBlock, 0, st>>>(L, L.r, L.x, L.t, om, kMgCoarseSweeps, h->scal);
      h->tm.kernel_launches++;
    } else {
      // the coarsest problem is spread over the ranks: the same fixed number of sweeps, planes exchanged
      const int g = mg_grid(h, L.n);
      double *cur = L.t, *oth = L.x;  // kMgCoarseSweeps is even: an odd number of swaps ends in x
      k_mg_smooth0<<<g, kBlock, 0, st>>>(L, L.r, cur, om, h->scal);
      for (int s = 1; s < kMgCoarseSweeps; ++s) {
        FVB_TRY(mg_halo(h, L, cur));
        k_mg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
        std::swap(cur, oth);
      }
      h->tm.kernel_launches += kMgCoarseSweeps;
    }
  }
  // going up
  for (int l = M.nlev - 2; l >= 0; --l) {
    MgLevel &L = M.lev[l];
    const int g = mg_grid(h, L.n);
    const MgLevel &C = M.lev[l + 1];
    // where the pre-smoothed iterate lives: t after an even number of swaps, x after an odd one
    double *cur = ((nu - 1) % 2 == 0) ? L.t : L.x;
    double *oth = cur == L.t ? L.x : L.t;
    k_mg_prolong<<<g, kBlock, 0, st>>>(L, C.ny, C.nz, C.x, cur, oc, h->scal);
    for (int s = 0; s < nu; ++s) {
      FVB_TRY(mg_halo(h, L, cur));
      k_mg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
      std::swap(cur, oth);
    }
    h->tm.kernel_launches += nu + 1;
    // cur == L.x by construction
  }
  return FVB_OK;
}

// ---- aggregation AMG on CSR rows (amg.cuh) ---------------------------------------------------------------------------
void amg_free(fvb_handle h) {
  if (!h->amg) return;
  for (void *p : h->amg->owned) {
    unsigned char *q = static_cast<unsigned char *>(p);
    dfree(h, q);
  }
  h->amg->owned.clear();
  h->amg->ready = false;
  h->amg->nlev = 0;
}

template <typename T>
int amg_alloc(fvb_handle h, T **p, int64_t n) {
  FVB_TRY(dalloc(h, p, n));
  h->amg->owned.push_back(*p);
  return FVB_OK;
}

// One pairwise-aggregation pass on a CSR matrix: agg[n] (row -> aggregate) and the number of aggregates.
int amg_pairwise(fvb_handle h, int n, const int *rowptr, const int *colidx, const double *vals, int *agg, int *nc_out,
                 int *d_match, int *d_prop, int *d_flag, int *d_scratch) {
  cudaStream_t st = h->stream;
  const int g = grid_for(n);
  FVB_CUDA(cudaMemsetAsync(d_match, 0xFF, sizeof(int) * (size_t)n, st));
  for (int round = 0; round < kAmgRounds; ++round) {
    k_amg_propose<<<g, kBlock, 0, st>>>(n, rowptr, colidx, vals, d_match, d_prop);
    k_amg_accept<<<g, kBlock, 0, st>>>(n, d_prop, d_match);
  }
  k_amg_attach<<<g, kBlock, 0, st>>>(n, rowptr, colidx, vals, d_match, d_prop);  // d_prop is free again: attach[]
  k_amg_roots<<<g, kBlock, 0, st>>>(n, d_match, d_prop, d_flag);
  exclusive_scan(d_flag, n, d_flag, d_scratch, st, &h->tm.kernel_launches);
  k_amg_number<<<g, kBlock, 0, st>>>(n, d_match, d_prop, d_flag, agg);
  h->tm.kernel_launches += 2 * kAmgRounds + 3;
  FVB_CUDA(memcpy_sync(st, nc_out, d_flag + n, sizeof(int), cudaMemcpyDeviceToHost));
  return FVB_OK;
}

// Galerkin coarse matrix of `fine` for the aggregation agg (nc aggregates); the arrays are owned by the hierarchy
// when keep is set, else by the caller (tmp_*: freed by the caller).  *ok = false: a coarse row overflowed.
int amg_galerkin(fvb_handle h, int n, const int *rowptr, const int *colidx, const double *vals, const int *agg, int nc,
                 int **memptr_out, int **mem_out, int **crowptr_out, int **ccol_out, double **cval_out, int *cnnz_out,
                 bool keep, int *d_scratch, bool *ok) {
  cudaStream_t st = h->stream;
  *ok = false;
  int *memptr = nullptr, *mem = nullptr, *cursor = nullptr, *crowptr = nullptr, *ccol = nullptr, *d_over = nullptr;
  double *cval = nullptr;
  auto A = [&](auto **p, int64_t cnt) { return keep ? amg_alloc(h, p, cnt) : dalloc(h, p, cnt); };
  FVB_TRY(A(&memptr, (int64_t)nc + 2));
  FVB_TRY(A(&mem, n));
  FVB_TRY(dalloc(h, &cursor, (int64_t)nc + 2));
  FVB_TRY(dalloc(h, &d_over, 1));
  FVB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)nc + 2), st));
  FVB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), st));
  k_amg_count_members<<<grid_for(n), kBlock, 0, st>>>(n, agg, cursor);
  exclusive_scan(cursor, nc, memptr, d_scratch, st, &h->tm.kernel_launches);
  FVB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)nc + 2), st));
  k_amg_fill_members<<<grid_for(n), kBlock, 0, st>>>(n, agg, memptr, cursor, mem);
  k_amg_sort_members<<<grid_for(nc), kBlock, 0, st>>>(nc, memptr, mem);
  FVB_TRY(A(&crowptr, (int64_t)nc + 2));
  const int gg = std::max(1, cdiv(nc, 128));
  k_amg_galerkin<false><<<gg, 128, 0, st>>>(nc, memptr, mem, rowptr, colidx, vals, agg, n, nullptr, nullptr, nullptr, cursor, d_over);
  exclusive_scan(cursor, nc, crowptr, d_scratch, st, &h->tm.kernel_launches);
  int cnnz = 0, over = 0;
  FVB_CUDA(cudaMemcpyAsync(&cnnz, crowptr + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
  FVB_CUDA(memcpy_sync(st, &over, d_over, sizeof(int), cudaMemcpyDeviceToHost));
  h->tm.kernel_launches += 4;
  if (!over) {
    FVB_TRY(A(&ccol, std::max(cnnz, 1)));
    FVB_TRY(A(&cval, std::max(cnnz, 1)));
    k_amg_galerkin<true><<<gg, 128, 0, st>>>(nc, memptr, mem, rowptr, colidx, vals, agg, n, crowptr, ccol, cval, nullptr, d_over);
    h->tm.kernel_launches++;
    *ok = true;
  }
  dfree(h, cursor); dfree(h, d_over);
  *memptr_out = memptr; *mem_out = mem; *crowptr_out = crowptr; *ccol_out = ccol; *cval_out = cval; *cnnz_out = cnnz;
  return FVB_OK;
}

// Build the hierarchy from the CSR rows of this rank (off-rank couplings are left out: block-local).
int amg_setup(fvb_handle h) {
  if (!h->amg) h->amg = new AmgState();
  amg_free(h);
  AmgState &M = *h->amg;
  if (!h->rowptr || h->nf_local < 2 || h->nnz >= INT_MAX - 1) return FVB_OK;
  cudaStream_t st = h->stream;
  AmgLevel *L = &M.lev[0];
  *L = AmgLevel{};
  L->n = (int)h->nf_local; L->nnz = (int)h->nnz;
  L->rowptr = h->rowptr; L->colidx = h->colidx; L->vals = h->vals;
  int nlev = 1;
  int *d_match = nullptr, *d_prop = nullptr, *d_flag = nullptr, *d_scratch = nullptr, *d_a1 = nullptr, *d_a2 = nullptr;
  const int64_t n0 = L->n;
  FVB_TRY(dalloc(h, &d_match, n0 + 2)); FVB_TRY(dalloc(h, &d_prop, n0 + 2)); FVB_TRY(dalloc(h, &d_flag, n0 + 2));
  FVB_TRY(dalloc(h, &d_scratch, scan_scratch_ints(n0 + 2))); FVB_TRY(dalloc(h, &d_a1, n0 + 2)); FVB_TRY(dalloc(h, &d_a2, n0 + 2));
  int status = FVB_OK;
  auto fin = [&](int s) {
    dfree(h, d_match); dfree(h, d_prop); dfree(h, d_flag); dfree(h, d_scratch); dfree(h, d_a1); dfree(h, d_a2);
    return s;
  };
#define AMG_TRY(expr) do { status = (expr); if (status != FVB_OK) { amg_free(h); return fin(status); } } while (0)
  while (true) {
    L = &M.lev[nlev - 1];
    const int n = L->n;
    AMG_TRY(amg_alloc(h, &L->dinv, n));
    AMG_TRY(amg_alloc(h, &L->x, n));
    AMG_TRY(amg_alloc(h, &L->t, n));
    if (nlev > 1) AMG_TRY(amg_alloc(h, &L->r, n));
    k_amg_dinv<<<grid_for(n), kBlock, 0, st>>>(n, L->rowptr, L->colidx, L->vals, L->dinv);
    h->tm.kernel_launches++;
    if (n <= kAmgCoarsest || nlev >= kAmgMaxLevels) break;
    // double pairwise aggregation: pair the rows, pair the pairs on the intermediate Galerkin matrix
    int n1 = 0, n2 = 0;
    AMG_TRY(amg_pairwise(h, n, L->rowptr, L->colidx, L->vals, d_a1, &n1, d_match, d_prop, d_flag, d_scratch));
    int *m1p = nullptr, *m1 = nullptr, *r1 = nullptr, *c1 = nullptr;
    double *v1 = nullptr;
    int nnz1 = 0;
    bool ok = false;
    AMG_TRY(amg_galerkin(h, n, L->rowptr, L->colidx, L->vals, d_a1, n1, &m1p, &m1, &r1, &c1, &v1, &nnz1, false, d_scratch, &ok));
    bool stop = !ok;
    if (ok) {
      status = amg_pairwise(h, n1, r1, c1, v1, d_a2, &n2, d_match, d_prop, d_flag, d_scratch);
      if (status == FVB_OK) stop = n2 > (int)(kAmgStall * n) || n2 < 1;
    }
    dfree(h, m1p); dfree(h, m1); dfree(h, r1); dfree(h, c1); dfree(h, v1);
    if (status != FVB_OK) { amg_free(h); return fin(status); }
    if (stop) break;
    AMG_TRY(amg_alloc(h, &L->agg, n));
    k_amg_compose<<<grid_for(n), kBlock, 0, st>>>(n, d_a1, d_a2, L->agg);
    h->tm.kernel_launches++;
    int *crow = nullptr, *ccol = nullptr;
    double *cval = nullptr;
    int cnnz = 0;
    AMG_TRY(amg_galerkin(h, n, L->rowptr, L->colidx, L->vals, L->agg, n2, &L->memptr, &L->mem, &crow, &ccol, &cval, &cnnz, true, d_scratch, &ok));
    if (!ok) { L->agg = nullptr; break; }  // a coarse row too long for the merge buffers: this level is the coarsest
    L->nc = n2;
    AmgLevel &C = M.lev[nlev];
    C = AmgLevel{};
    C.n = n2; C.nnz = cnnz; C.rowptr = crow; C.colidx = ccol; C.vals = cval;
    ++nlev;
  }
#undef AMG_TRY
  M.nlev = nlev;
  M.ready = true;
  FVB_CUDA(cudaStreamSynchronize(st));
  return fin(FVB_OK);
}

// z = M^-1 r : one V(nu,nu) cycle from a zero initial guess; the result lands in lev[0].x.
int amg_vcycle(fvb_handle h, const double *r) {
  AmgState &M = *h->amg;
  cudaStream_t st = h->stream;
  const int nu = h->mg_nu;
  const double om = h->mg_omega, oc = h->mg_oc;
  M.lev[0].r = const_cast<double *>(r);
  for (int l = 0; l + 1 < M.nlev; ++l) {
    AmgLevel &L = M.lev[l];
    const int g = grid_for(L.n);
    double *cur = L.t, *oth = L.x;  // 2*nu-1 swaps in total: start in t to finish in x
    k_amg_smooth0<<<g, kBlock, 0, st>>>(L, L.r, cur, om, h->scal);
    for (int s = 1; s < nu; ++s) {
      k_amg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
      std::swap(cur, oth);
    }
    k_amg_restrict<<<grid_for(L.nc), kBlock, 0, st>>>(L, L.r, cur, M.lev[l + 1].r, h->scal);
    h->tm.kernel_launches += nu + 1;
  }
  {
    AmgLevel &L = M.lev[M.nlev - 1];
    if (L.n <= kAmgOneCta) {
      k_amg_coarse_solve<<<1, kBlock, 0, st>>>(L, L.r, L.x, L.t, om, kAmgCoarseSweeps, h->scal);
      h->tm.kernel_launches++;
    } else {
      // the aggregation stalled before the level got small: the same fixed number of sweeps, one launch each
      const int g = grid_for(L.n);
      double *cur = L.t, *oth = L.x;  // kAmgCoarseSweeps is even: an odd number of swaps ends in x
      k_amg_smooth0<<<g, kBlock, 0, st>>>(L, L.r, cur, om, h->scal);
      for (int s = 1; s < kAmgCoarseSweeps; ++s) {
        k_amg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
        std::swap(cur, oth);
      }
      h->tm.kernel_launches += kAmgCoarseSweeps;
    }
  }
  for (int l = M.nlev - 2; l >= 0; --l) {
    AmgLevel &L = M.lev[l];
    const int g = grid_for(L.n);
    double *cur = ((nu - 1) % 2 == 0) ? L.t : L.x;
    double *oth = cur == L.t ? L.x : L.t;
    k_amg_prolong<<<g, kBlock, 0, st>>>(L, M.lev[l + 1].x, cur, oc, h->scal);
    for (int s = 0; s < nu; ++s) {
      k_amg_smooth<<<g, kBlock, 0, st>>>(L, L.r, cur, oth, om, h->scal);
      std::swap(cur, oth);
    }
    h->tm.kernel_launches += nu + 1;
  }
  return FVB_OK;
}

// Which hierarchy serves precond kind 1: the geometric one of mg.cuh when the matrix is box-structured, else the
// algebraic one on the CSR rows (single rank).
bool use_geometric_mg(fvb_handle h) { return h->precond_request == 1 && h->mg && h->mg->ready && h->fmt_request != 1; }
bool use_amg(fvb_handle h) {
  return h->precond_request == 1 && !use_geometric_mg(h) && h->nranks == 1 && h->amg && h->amg->ready;
}
int precond_setup(fvb_handle h) {
  FVB_TRY(mg_setup(h, true));
  if (!(h->mg && h->mg->ready) && h->nranks == 1) {
    FVB_TRY(ensure_csr(h));
    if (h->rowptr) FVB_TRY(amg_setup(h));
  }
  return FVB_OK;
}

// CG preconditioned by the V-cycle, steady operator only.  x0 (if any) already in h->x.
int pcg_mg_run(fvb_handle h, const double *rhs, bool have_x0, double rtol, int64_t maxiter, int64_t *iters,
               int *converged);

// Jacobi-PCG on (A + sigma*D) x = rhs.  x0 (if any) already sits in h->x.  Result in h->x.
// Cold-started steady solves on the diagonal format run the symmetrically scaled recurrence
// (pcg.cuh, SC = true): same iterates in exact arithmetic, 112 instead of 128 bytes per row.
int pcg_run(fvb_handle h, const double *rhs, bool have_x0, double sigma, double rtol, int64_t maxiter,
            int64_t *iters, int *converged) {
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  const int vg = vgrid(h, n);
  PeerRed pr;
  int fin;
  FVB_TRY(ensure_hist(h, maxiter));
  h->prof_seen = 0;
  h->prof_count = 0;
  bool sc = false;
  if (sigma == 0.0 && !have_x0 && h->scale_request != 1 && h->fmt_request != 1) {  // rank-uniform conditions
    FVB_TRY(build_scaled(h));
    sc = h->scale_state == 1;
  }
  h->last_solve_scaled = sc;
  k_set_scal<<<1, 1, 0, st>>>(h->scal, rtol, (long long)maxiter, (long long)h->hist_cap);
  h->tm.kernel_launches++;
  if (sc) {
    fin = red_mode(h, &pr);
    k_pcg_init<true><<<vg, kBlock, 0, st>>>(n, rhs, nullptr, 0, h->sinv, h->x, h->r, h->partials, h->ticket,
                                            h->scal, fin, pr);
  } else {
    if (!h->dinv) FVB_TRY(dalloc(h, &h->dinv, n));
    k_make_dinv<<<vg, kBlock, 0, st>>>(n, h->diag, h->Dvec, sigma, h->dinv);
    h->tm.kernel_launches++;
    if (have_x0) {
      FVB_CUDA(cudaMemcpyAsync(h->u, h->x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
      FVB_TRY(launch_spmv(h, h->u, h->c, sigma, false));
    }
    fin = red_mode(h, &pr);
    k_pcg_init<false><<<vg, kBlock, 0, st>>>(n, rhs, h->c, have_x0 ? 1 : 0, h->dinv, h->x, h->r, h->partials,
                                             h->ticket, h->scal, fin, pr);
  }
  h->tm.kernel_launches++;
  if (!fin) FVB_TRY(allreduce_fin(h, 2, FIN_INIT));
  // Enqueue iterations in batches; poll the device scalars one batch behind so the GPU
  // never waits for the host.
  int64_t enq = 0;
  int batch = n < (1 << 20) ? 4 : 8, slot = 0;  // small systems are launch-bound: fewer iterations enqueued past convergence
  bool have_prev = false;
  bool stop = false;
  while (!stop) {
    int64_t todo = std::min<int64_t>(batch, maxiter - enq);
    for (int64_t it = 0; it < todo; ++it) {
      const FusedPush fp = fused_push(h);  // slab ranks over peer memory: the halo push rides on k_update_u
      if (fp.npeers > 0) {
        if (sc) k_update_u<true, true><<<vg, kBlock, 0, st>>>(n, nullptr, h->r, h->u, h->x, h->scal, fp, h->ticket + 1);
        else k_update_u<false, true><<<vg, kBlock, 0, st>>>(n, h->dinv, h->r, h->u, h->x, h->scal, fp, h->ticket + 1);
      } else {
        if (sc) k_update_u<true, false><<<vg, kBlock, 0, st>>>(n, nullptr, h->r, h->u, h->x, h->scal, fp, h->ticket + 1);
        else k_update_u<false, false><<<vg, kBlock, 0, st>>>(n, h->dinv, h->r, h->u, h->x, h->scal, fp, h->ticket + 1);
      }
      h->tm.kernel_launches++;
      FVB_TRY(launch_spmv(h, h->u, h->c, sigma, true, sc, true, &fin, fp.npeers > 0));
      if (!fin) FVB_TRY(allreduce_fin(h, 1, FIN_UC));
      fin = red_mode(h, &pr);
      if (sc) k_update_xr<true><<<vg, kBlock, 0, st>>>(n, h->c, h->diag, h->r, h->partials, h->ticket, h->scal, h->hist, fin, pr);
      else k_update_xr<false><<<vg, kBlock, 0, st>>>(n, h->c, h->dinv, h->r, h->partials, h->ticket, h->scal, h->hist, fin, pr);
      h->tm.kernel_launches++;
      if (!fin) FVB_TRY(allreduce_fin(h, 2, FIN_ITER));
    }
    enq += todo;
    FVB_CUDA(cudaMemcpyAsync(&h->scal_host[slot], h->scal, sizeof(PcgScal), cudaMemcpyDeviceToHost, st));
    FVB_CUDA(cudaEventRecord(h->ev[6 + slot], st));
    if (have_prev) {
      FVB_CUDA(cudaEventSynchronize(h->ev[6 + (slot ^ 1)]));
      if (h->scal_host[slot ^ 1].done) stop = true;
    }
    have_prev = true;
    slot ^= 1;
    if (enq >= maxiter) stop = true;
    if (batch < 32) batch *= 2;
  }
  if (sc) k_finish_x_scaled<<<vg, kBlock, 0, st>>>(n, h->u, h->sinv, h->x, h->scal);
  else k_finish_x<<<vg, kBlock, 0, st>>>(n, h->u, h->x, h->scal);
  h->tm.kernel_launches++;
  FVB_CUDA(cudaStreamSynchronize(st));
  FVB_CUDA(memcpy_sync(h->stream, &h->scal_host[0], h->scal, sizeof(PcgScal), cudaMemcpyDeviceToHost));
  if (iters) *iters = h->scal_host[0].iter;
  if (converged) *converged = h->scal_host[0].converged;
  if (h->peer && h->peer->active) {
    int perr = 0;
    FVB_CUDA(memcpy_sync(st, &perr, &h->peer->mail->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (perr) return set_error(FVB_ERR_NCCL, "peer-memory exchange timed out waiting for another rank");
  }
  h->tm.spmv_ms_total = 0;
  h->tm.spmv_samples = 0;
  for (int k = 0; k < h->prof_count; ++k) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->prof_ev[2 * k], h->prof_ev[2 * k + 1]) == cudaSuccess) {
      h->tm.spmv_ms_total += ms;
      h->tm.spmv_samples++;
    }
  }
  FVB_CUDA(cudaGetLastError());
  return FVB_OK;
}

int pcg_mg_run(fvb_handle h, const double *rhs, bool have_x0, double rtol, int64_t maxiter, int64_t *iters,
               int *converged) {
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  const int fin = h->nranks == 1 ? 1 : 0;
  const int vg = vgrid(h, n);
  const bool use_amg = !(h->mg && h->mg->ready && h->fmt_request != 1);  // else the geometric hierarchy of mg.cuh
  FVB_TRY(ensure_hist(h, maxiter));
  h->last_solve_scaled = false;
  h->prof_seen = 0;
  h->prof_count = 0;
  if (!h->dinv) FVB_TRY(dalloc(h, &h->dinv, n));
  k_set_scal<<<1, 1, 0, st>>>(h->scal, rtol, (long long)maxiter, (long long)h->hist_cap);
  k_make_dinv<<<vg, kBlock, 0, st>>>(n, h->diag, nullptr, 0.0, h->dinv);
  h->tm.kernel_launches += 2;
  if (have_x0) {
    FVB_CUDA(cudaMemcpyAsync(h->u, h->x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    FVB_TRY(launch_spmv(h, h->u, h->c, 0.0, false));
  }
  k_pcg_init<false><<<vg, kBlock, 0, st>>>(n, rhs, h->c, have_x0 ? 1 : 0, h->dinv, h->x, h->r, h->partials, h->ticket,
                                           h->scal, fin, PeerRed{nullptr, 0ull});
  h->tm.kernel_launches++;
  if (!fin) FVB_TRY(allreduce_fin(h, 2, FIN_INIT));
  int64_t enq = 0;
  int batch = 2, slot = 0;
  bool have_prev = false, stop = false;
  while (!stop) {
    int64_t todo = std::min<int64_t>(batch, maxiter - enq);
    for (int64_t it = 0; it < todo; ++it) {
      if (use_amg) FVB_TRY(amg_vcycle(h, h->r));
      else FVB_TRY(mg_vcycle(h, h->r));
      const double *z = use_amg ? h->amg->lev[0].x : h->mg->lev[0].x;
      k_mgpcg_rz<<<vg, kBlock, 0, st>>>(n, h->r, z, h->partials, h->ticket, h->scal, fin);
      h->tm.kernel_launches++;
      if (!fin) FVB_TRY(allreduce_fin(h, 1, FIN_RZ));
      k_mgpcg_update_u<<<vg, kBlock, 0, st>>>(n, z, h->u, h->x, h->scal);
      h->tm.kernel_launches++;
      FVB_TRY(launch_spmv(h, h->u, h->c, 0.0, true));
      if (!fin) FVB_TRY(allreduce_fin(h, 1, FIN_UC));
      k_mgpcg_update_r<<<vg, kBlock, 0, st>>>(n, h->c, h->r, h->partials, h->ticket, h->scal, h->hist, fin);
      h->tm.kernel_launches++;
      if (!fin) FVB_TRY(allreduce_fin(h, 1, FIN_R));
    }
    enq += todo;
    FVB_CUDA(cudaMemcpyAsync(&h->scal_host[slot], h->scal, sizeof(PcgScal), cudaMemcpyDeviceToHost, st));
    FVB_CUDA(cudaEventRecord(h->ev[6 + slot], st));
    if (have_prev) {
      FVB_CUDA(cudaEventSynchronize(h->ev[6 + (slot ^ 1)]));
      if (h->scal_host[slot ^ 1].done) stop = true;
    }
    have_prev = true;
    slot ^= 1;
    if (enq >= maxiter) stop = true;
  }
  k_finish_x<<<vg, kBlock, 0, st>>>(n, h->u, h->x, h->scal);
  h->tm.kernel_launches++;
  FVB_CUDA(cudaStreamSynchronize(st));
  FVB_CUDA(memcpy_sync(h->stream, &h->scal_host[0], h->scal, sizeof(PcgScal), cudaMemcpyDeviceToHost));
  if (iters) *iters = h->scal_host[0].iter;
  if (converged) *converged = h->scal_host[0].converged;
  if (h->peer && h->peer->active) {
    int perr = 0;
    FVB_CUDA(memcpy_sync(st, &perr, &h->peer->mail->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (perr) return set_error(FVB_ERR_NCCL, "peer-memory exchange timed out waiting for another rank");
  }
  h->tm.spmv_ms_total = 0;
  h->tm.spmv_samples = 0;
  for (int k = 0; k < h->prof_count; ++k) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->prof_ev[2 * k], h->prof_ev[2 * k + 1]) == cudaSuccess) {
      h->tm.spmv_ms_total += ms;
      h->tm.spmv_samples++;
    }
  }
  FVB_CUDA(cudaGetLastError());
  return FVB_OK;
}

// Push/wait tables of the peer-memory halo exchange from the installed halo plan (P.tab already holds every
// rank's vector and mailbox as seen from this process); shared by fvb_peer_import (CUDA IPC mappings) and the
// single-process multi-device front end (multi_impl.h: plain peer access).
int peer_finish_plan(fvb_handle h, const int64_t *send_dst_index) {
  PeerState &P = *h->peer;
  P.push = HaloPlanDev{};
  P.wait = HaloPlanDev{};
  long long so = 0;
  for (size_t p = 0; p < h->peers.size(); ++p) {
    if (h->send_counts[p] > 0) {
      const int k = P.push.npeers++;
      P.push.peer[k] = h->peers[p];
      P.push.send_begin[k] = so;
      P.push.dst_off[k] = send_dst_index[p];
      P.push.send_begin[k + 1] = so + h->send_counts[p];
    }
    so += h->send_counts[p];
    if (h->recv_counts[p] > 0) P.wait.peer[P.wait.npeers++] = h->peers[p];
  }
  // send_rows is laid out peer after peer in plan order; push.send_begin indexes it directly
  // only if peers without sends contribute nothing, which holds because their count is 0.
  if (!P.d_tab) FVB_CUDA(cudaMalloc((void **)&P.d_tab, sizeof(PeerTable)));
  FVB_CUDA(memcpy_sync(h->stream, P.d_tab, &P.tab, sizeof(PeerTable), cudaMemcpyHostToDevice));
  if (const char *env = getenv("FVB_FUSED_ALLREDUCE")) P.fused = atoi(env) != 0;
  P.active = true;
  return FVB_OK;
}

bool valid_slot(int s) { return s >= 0 && s < FVB_NSLOT; }
int ensure_slot(fvb_handle h, int s) {
  if (!valid_slot(s)) return set_error(FVB_ERR_BAD_INPUT, "vector slot out of range");
  if (!h->slots[s]) {
    FVB_TRY(dalloc(h, &h->slots[s], h->nf_local));
    FVB_CUDA(cudaMemsetAsync(h->slots[s], 0, sizeof(double) * (size_t)std::max<int64_t>(h->nf_local, 1), h->stream));
  }
  return FVB_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int fvb_version(void) { return 100; }
const char *fvb_last_error(void) { return g_last_error.c_str(); }

int fvb_device_count(int *count) {
  FVB_CUDA(cudaGetDeviceCount(count));
  return FVB_OK;
}

int fvb_create(int device, fvb_handle *out) {
  if (!out) return set_error(FVB_ERR_BAD_INPUT, "null out pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(FVB_ERR_CUDA, "no CUDA device available (libfvb200 has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return set_error(FVB_ERR_BAD_INPUT, "device index out of range");
  FVB_CUDA(cudaSetDevice(device));
  fvb_handle h = new fvb_handle_s();
  h->device = device;
  {
    const char *env = getenv("FVB_ARENA");
    if (!env || atoi(env) != 0) h->arena = new Arena(arena_chunk_alloc, arena_chunk_free);
  }
  if (const char *env = getenv("FVB_PCG_SCALING")) h->scale_request = atoi(env) == 0 ? 1 : 0;  // A/B measurements
  if (const char *env = getenv("FVB_BOX")) h->box_request = atoi(env) == 0 ? 1 : 0;  // A/B: general path only
  h->fused_halo = getenv("FVB_FUSED_HALO_OFF") == nullptr;                            // A/B: five-launch halo exchange
  if (const char *env = getenv("FVB_SPMV_FORMAT")) {  // initial fvb_set_spmv_format value
    const int f = atoi(env);
    if (f >= 0 && f <= 3) h->fmt_request = f;
  }
  FVB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto &ev : h->ev) FVB_CUDA(cudaEventCreate(&ev));
  FVB_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  {
    cudaMemPool_t pool;
    FVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    FVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  FVB_TRY(dalloc(h, &h->scal, 1));
  FVB_TRY(dalloc(h, &h->ticket, 4));
  FVB_CUDA(cudaMemsetAsync(h->ticket, 0, 4 * sizeof(unsigned int), h->stream));
  FVB_CUDA(cudaMemsetAsync(h->scal, 0, sizeof(PcgScal), h->stream));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  FVB_CUDA(cudaMallocHost((void **)&h->scal_host, 2 * sizeof(PcgScal)));
  FVB_CUDA(cudaFuncSetAttribute(k_spmv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpmvSmem)));
  FVB_CUDA(cudaFuncSetAttribute(k_spmv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpmvSmem)));
  *out = h;
  return FVB_OK;
}

int fvb_destroy(fvb_handle h) {
  if (!h) return FVB_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_problem(h);
  delete h->mg;
  delete h->amg;
  if (h->peer) {
    for (int r = 0; r < kMaxRanks; ++r) {
      if (h->peer->opened_u[r]) cudaIpcCloseMemHandle(h->peer->opened_u[r]);
      if (h->peer->opened_mail[r]) cudaIpcCloseMemHandle(h->peer->opened_mail[r]);
    }
    if (h->peer->mail) cudaFree(h->peer->mail);
    if (h->peer->d_tab) cudaFree(h->peer->d_tab);
    delete h->peer;
  }
  if (h->nranks > 1 && h->u) cudaFree(h->u);
  if (h->comm) {
    if (h->comm->comm) nccl().CommDestroy(h->comm->comm);
    delete h->comm;
  }
  dfree(h, h->scal); dfree(h, h->ticket);
  delete h->arena;  // gives every chunk back (cudaFree synchronises the device)
  h->arena = nullptr;
  if (h->scal_host) cudaFreeHost(h->scal_host);
  for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : h->prof_ev) if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return FVB_OK;
}

int fvb_comm_unique_id(uint8_t id[FVB_UNIQUE_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == FVB_UNIQUE_ID_BYTES, "ncclUniqueId size");
  std::string why = nccl().load();
  if (!why.empty()) return set_error(FVB_ERR_NCCL, why);
  ncclUniqueId uid;
  FVB_NCCL(nccl().GetUniqueId(&uid));
  memcpy(id, &uid, FVB_UNIQUE_ID_BYTES);
  return FVB_OK;
}

int fvb_comm_init(fvb_handle h, int nranks, int rank, const uint8_t id[FVB_UNIQUE_ID_BYTES]) {
  FVB_TRY(check_handle(h, false));
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(FVB_ERR_BAD_INPUT, "bad rank/nranks");
  // A problem assembled under the previous rank count owns buffers whose allocation route depends on it
  // (u: arena when single-rank, plain cudaMalloc for IPC export otherwise): drop it while that is still known.
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  free_problem(h);
  if (h->nranks > 1 && h->u) { cudaFree(h->u); h->u = nullptr; h->u_cap = 0; }
  h->nranks = nranks;
  h->rank = rank;
  if (nranks == 1) return FVB_OK;
  std::string why = nccl().load();
  if (!why.empty()) return set_error(FVB_ERR_NCCL, why);
  ncclUniqueId uid;
  memcpy(&uid, id, FVB_UNIQUE_ID_BYTES);
  h->comm = new Comm();
  FVB_NCCL(nccl().CommInitRank(&h->comm->comm, nranks, uid, rank));
  return FVB_OK;
}

int fvb_assemble(fvb_handle h, int64_t n_nodes, int64_t node_lo1, int64_t node_hi1, int64_t n_faces,
                 const int64_t *neighbors, const double *aol, const double *cond, int64_t n_cond,
                 const int64_t *metaindex, int logk, const double *sources, int64_t nd, const int64_t *dnodes,
                 const double *dheads) {
  FVB_TRY(check_handle(h, false));
  if (n_nodes < 0 || n_faces < 0 || nd < 0 || n_cond < 0) return set_error(FVB_ERR_BAD_INPUT, "negative size");
  if (node_lo1 < 1 || node_hi1 > n_nodes || node_hi1 < node_lo1 - 1)
    return set_error(FVB_ERR_BAD_INPUT, "owned node range must satisfy 1 <= lo, hi <= N");
  if ((n_faces && (!neighbors || !aol || !cond)) || (n_nodes && !sources) || (nd && (!dnodes || !dheads)))
    return set_error(FVB_ERR_BAD_INPUT, "null input array");
  const int64_t n_own = node_hi1 - node_lo1 + 1;
  // Local row / node indices are 32-bit.  The face-indexed arrays of the general path (adjacency, CSR) are too, but
  // regulargrid-ordered problems never build them (box.cuh), so that limit is only enforced once the closed-form
  // path has been ruled out, below.
  if (n_own >= INT_MAX - 1 || nd >= INT_MAX - 1)
    return set_error(FVB_ERR_BAD_INPUT, "per-GPU part too large for 32-bit local indices; use more ranks");
  const bool faces_fit_32 = 2 * n_faces < INT_MAX - 1;
  if (!faces_fit_32 && (h->box_request == 1 || h->fmt_request == 1 || n_faces > 4 * n_own))
    return set_error(FVB_ERR_BAD_INPUT, "per-GPU part too large for 32-bit local indices; use more ranks");
  const bool dbg = getenv("FVB_DEBUG") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tdbg0 = now();
  double tdbg = tdbg0;
  auto lap = [&](const char *what) {
    if (dbg) { double t = now(); fprintf(stderr, "[fvb_assemble] %-28s %8.1f ms\n", what, (t - tdbg) * 1e3); tdbg = t; }
  };
  free_problem(h);
  lap("free previous problem");
  cudaStream_t st = h->stream;
  h->n_nodes = n_nodes; h->node_lo = node_lo1 - 1; h->node_hi = node_hi1; h->n_own_nodes = n_own;
  h->n_faces = n_faces; h->n_dirichlet = nd;
  h->logk = logk ? 1 : 0;
  const bool whole = (h->node_lo == 0 && h->node_hi == n_nodes);

  // ---- host -> device ----------------------------------------------------------------------
  FVB_CUDA(cudaEventRecord(h->ev[0], st));
  int64_t *d_nb = nullptr, *d_dnodes = nullptr;
  double *d_cond = nullptr;
  // inputs that already live on this device are read in place for the duration of the call (the neighbor list of
  // a 512^3 slab is 6.4 GB: no second copy); everything retained beyond the call is copied as before
  auto on_this_device = [&](const void *p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice && a.device == h->device;
  };
  const bool nb_borrowed = n_faces > 0 && on_this_device(neighbors);
  const bool cond_borrowed = n_cond > 0 && on_this_device(cond);
  std::vector<int64_t> dn_sorted;  // ascending distinct 0-based Dirichlet nodes of the whole problem (slab ranks)
  int *d_dslot = nullptr, *d_cnt = nullptr, *d_scratch = nullptr, *d_err = nullptr;
  int64_t *d_dsorted = nullptr, *d_refs = nullptr;
  int *d_dsorted_slot = nullptr;
  unsigned long long *d_noff = nullptr;
  auto cleanup = [&]() {
    if (!nb_borrowed) dfree(h, d_nb);
    if (!cond_borrowed) dfree(h, d_cond);
    d_nb = nullptr; d_cond = nullptr;
    dfree(h, d_dnodes); dfree(h, d_dslot); dfree(h, d_cnt); dfree(h, d_scratch); dfree(h, d_err);
    dfree(h, d_dsorted); dfree(h, d_dsorted_slot); dfree(h, d_refs); dfree(h, d_noff);
  };
#define A_TRY(expr) do { int s__ = (expr); if (s__ != FVB_OK) { cleanup(); free_problem(h); return s__; } } while (0)
#define A_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); free_problem(h); \
    return set_error(FVB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)

  if (nb_borrowed) d_nb = const_cast<int64_t *>(neighbors);
  else A_TRY(dalloc(h, &d_nb, 2 * n_faces));
  A_TRY(dalloc(h, &h->aol, n_faces));
  if (cond_borrowed) d_cond = const_cast<double *>(cond);
  else A_TRY(dalloc(h, &d_cond, n_cond));
  A_TRY(dalloc(h, &h->sources, n_own));
  A_TRY(dalloc(h, &d_dnodes, nd));
  A_TRY(dalloc(h, &h->dheads, nd));
  if (!nb_borrowed) A_CUDA(cudaMemcpyAsync(d_nb, neighbors, sizeof(int64_t) * 2 * (size_t)n_faces, cudaMemcpyDefault, st));
  A_CUDA(cudaMemcpyAsync(h->aol, aol, sizeof(double) * (size_t)n_faces, cudaMemcpyDefault, st));
  if (!cond_borrowed) A_CUDA(cudaMemcpyAsync(d_cond, cond, sizeof(double) * (size_t)n_cond, cudaMemcpyDefault, st));
  A_CUDA(cudaMemcpyAsync(h->sources, sources, sizeof(double) * (size_t)n_own, cudaMemcpyDefault, st));
  A_CUDA(cudaMemcpyAsync(d_dnodes, dnodes, sizeof(int64_t) * (size_t)nd, cudaMemcpyDefault, st));
  A_CUDA(cudaMemcpyAsync(h->dheads, dheads, sizeof(double) * (size_t)nd, cudaMemcpyDefault, st));
  if (metaindex) {
    A_TRY(dalloc(h, &h->meta, n_faces));
    A_CUDA(cudaMemcpyAsync(h->meta, metaindex, sizeof(int64_t) * (size_t)n_faces, cudaMemcpyDefault, st));
  }
  A_CUDA(cudaEventRecord(h->ev[1], st));
  lap("alloc + enqueue H2D");

  // ---- Dirichlet table for off-rank endpoints (host: ND is small next to N) ---------------------
  int64_t nd_sorted = 0;
  h->row_start = 0;
  if (!whole) {
    std::vector<int64_t> hd((size_t)nd);
    A_CUDA(cudaMemcpyAsync(hd.data(), d_dnodes, sizeof(int64_t) * (size_t)nd, cudaMemcpyDeviceToHost, st));
    A_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> nodes;
    std::vector<int> slot;
    dirichlet_table(hd, nodes, slot);  // sorted by node, last duplicate wins (host_util.h)
    nd_sorted = (int64_t)nodes.size();
    dn_sorted = nodes;
    A_TRY(dalloc(h, &d_dsorted, nd_sorted));
    A_TRY(dalloc(h, &d_dsorted_slot, nd_sorted));
    A_CUDA(cudaMemcpyAsync(d_dsorted, nodes.data(), sizeof(int64_t) * nodes.size(), cudaMemcpyHostToDevice, st));
    A_CUDA(cudaMemcpyAsync(d_dsorted_slot, slot.data(), sizeof(int) * slot.size(), cudaMemcpyHostToDevice, st));
    A_CUDA(cudaStreamSynchronize(st));
    int64_t below = std::lower_bound(nodes.begin(), nodes.end(), h->node_lo) - nodes.begin();
    int64_t valid = std::lower_bound(nodes.begin(), nodes.end(), n_nodes) - std::lower_bound(nodes.begin(), nodes.end(), (int64_t)0);
    h->row_start = h->node_lo - below;
    h->nf_global = n_nodes - valid;
  }

  // ---- 1. node map --------------------------------------------------------------------------
  A_TRY(dalloc(h, &d_err, ERR_COUNT));
  {
    int init[ERR_COUNT];
    for (int &v : init) v = INT_MAX;
    A_CUDA(cudaMemcpyAsync(d_err, init, sizeof(init), cudaMemcpyHostToDevice, st));
  }
  A_TRY(dalloc(h, &d_dslot, n_own));
  A_TRY(dalloc(h, &d_cnt, std::max(n_own, n_faces) + 2));
  A_TRY(dalloc(h, &d_scratch, scan_scratch_ints(std::max(n_own, 2 * n_faces) + 1)));
  A_CUDA(cudaMemsetAsync(d_dslot, 0xFF, sizeof(int) * (size_t)std::max<int64_t>(n_own, 1), st));
  if (nd) {
    k_mark_dirichlet<<<grid_for(nd), kBlock, 0, st>>>(d_dnodes, nd, n_nodes, h->node_lo, h->node_hi, h->sources,
                                                      d_dslot, d_err);
    h->tm.kernel_launches++;
  }
  int nf_local = 0;
  if (n_own) {
    k_free_flags<<<grid_for(n_own), kBlock, 0, st>>>(d_dslot, n_own, d_cnt);
    h->tm.kernel_launches++;
    exclusive_scan(d_cnt, n_own, d_cnt, d_scratch, st, &h->tm.kernel_launches);
    A_CUDA(cudaMemcpyAsync(&nf_local, d_cnt + n_own, sizeof(int), cudaMemcpyDeviceToHost, st));
    A_CUDA(cudaStreamSynchronize(st));
  }
  h->nf_local = nf_local;
  lap("nodemap scan (sync)");
  if (whole) h->nf_global = nf_local;
  A_TRY(dalloc(h, &h->nodemap, n_own));
  A_TRY(dalloc(h, &h->row2node, nf_local));
  if (n_own) {
    k_finish_nodemap<<<grid_for(n_own), kBlock, 0, st>>>(d_dslot, d_cnt, n_own, h->nodemap, h->row2node);
    h->tm.kernel_launches++;
  }

  // ---- 2. per-face conductance ----------------------------------------------------------------
  A_TRY(dalloc(h, &h->cface, n_faces));
  if (n_faces) {
    k_face_conductance<<<grid_for(n_faces), kBlock, 0, st>>>(n_faces, d_cond, n_cond, h->meta, h->aol, logk,
                                                             h->cface, d_err);
    h->tm.kernel_launches++;
  }

  // ---- closed-form path for regulargrid-ordered face lists (box.cuh): no adjacency, no CSR --------------------
  if (h->box_request != 1 && h->fmt_request != 1 && nf_local >= 2) {
    BoxDesc B;
    bool is_box = false;
    A_TRY(box_detect(h, d_nb, dn_sorted, &B, &is_box));
    lap("box detection (sync)");
    if (is_box) {
      h->d_dsorted = d_dsorted; h->d_dsorted_slot = d_dsorted_slot; h->nd_sorted = nd_sorted;  // retained
      d_dsorted = nullptr; d_dsorted_slot = nullptr;
      A_TRY(box_install(h, B));
      if (h->precond_request == 1) A_TRY(precond_setup(h));
      A_CUDA(cudaEventRecord(h->ev[2], st));
      int herr[ERR_COUNT];
      A_CUDA(cudaMemcpyAsync(herr, d_err, sizeof(herr), cudaMemcpyDeviceToHost, st));
      A_CUDA(cudaStreamSynchronize(st));
      A_CUDA(cudaGetLastError());
      if (herr[ERR_SRC_ON_DIRICHLET] != INT_MAX) {
        int64_t node = 0;
        memcpy_sync(h->stream, &node, d_dnodes + herr[ERR_SRC_ON_DIRICHLET], sizeof(int64_t), cudaMemcpyDeviceToHost);
        cleanup(); free_problem(h);
        return set_error(FVB_ERR_BAD_INPUT, "There cannot be a source at a Dirichlet node, but node " + std::to_string(node) +
                                                " is a Dirichlet node where a source is located.");
      }
      if (herr[ERR_BAD_NODE] != INT_MAX) {
        cleanup(); free_problem(h);
        return set_error(FVB_ERR_BAD_INPUT, "node index out of range 1..N (neighbors or dirichletnodes entry " +
                                                std::to_string(herr[ERR_BAD_NODE] + 1) + ")");
      }
      if (herr[ERR_BAD_META] != INT_MAX) {
        cleanup(); free_problem(h);
        return set_error(FVB_ERR_BAD_INPUT, "metaindex(" + std::to_string(herr[ERR_BAD_META] + 1) + ") is outside conductivities");
      }
      float msb = 0;
      cudaEventElapsedTime(&msb, h->ev[0], h->ev[1]); h->tm.h2d_ms = msb;
      cudaEventElapsedTime(&msb, h->ev[1], h->ev[2]); h->tm.assemble_ms = msb;
      lap("box rows (sync)");
      cleanup();
      h->assembled = true;
      h->halo_ready = (h->nranks == 1);
      return FVB_OK;
    }
  }
  if (!faces_fit_32) {
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "per-GPU part too large for 32-bit local indices; use more ranks");
  }

  // ---- 3. adjacency ------------------------------------------------------------------------------
  Resolver res{h->nodemap, n_nodes, h->node_lo, h->node_hi, d_dsorted, d_dsorted_slot, nd_sorted};
  A_TRY(dalloc(h, &d_noff, 1));
  A_TRY(dalloc(h, &h->adjptr, (int64_t)nf_local + 1));
  unsigned long long n_off = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    A_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * ((size_t)nf_local + 1), st));
    A_CUDA(cudaMemsetAsync(d_noff, 0, sizeof(unsigned long long), st));
    if (n_faces) {
      k_adjacency<0><<<grid_for(n_faces), kBlock, 0, st>>>(n_faces, d_nb, res, d_cnt, nullptr, nullptr, nullptr,
                                                           d_noff, d_refs, nullptr, 0, nf_local, d_err);
      h->tm.kernel_launches++;
    }
    A_CUDA(cudaMemcpyAsync(&n_off, d_noff, sizeof(n_off), cudaMemcpyDeviceToHost, st));
    A_CUDA(cudaStreamSynchronize(st));
    if (n_off == 0 || d_refs) break;
    A_TRY(dalloc(h, &d_refs, (int64_t)n_off));  // second attempt records the references
  }
  h->n_halo = 0;
  if (n_off) {
    h->halo_host.resize((size_t)n_off);
    A_CUDA(memcpy_sync(h->stream, h->halo_host.data(), d_refs, sizeof(int64_t) * (size_t)n_off, cudaMemcpyDeviceToHost));
    sort_unique_i64(h->halo_host);  // ascending distinct columns (host_util.h: bitmap pass for dense planes)
    h->n_halo = (int64_t)h->halo_host.size();
    A_TRY(dalloc(h, &h->halo_glob, h->n_halo));
    A_CUDA(memcpy_sync(h->stream, h->halo_glob, h->halo_host.data(), sizeof(int64_t) * (size_t)h->n_halo, cudaMemcpyHostToDevice));
  }
  exclusive_scan(d_cnt, nf_local, h->adjptr, d_scratch, st, &h->tm.kernel_launches);
  int n_adj = 0;
  A_CUDA(cudaMemcpyAsync(&n_adj, h->adjptr + nf_local, sizeof(int), cudaMemcpyDeviceToHost, st));
  A_CUDA(cudaStreamSynchronize(st));
  h->n_adj = n_adj;
  lap("adjacency count (sync)");
  A_TRY(dalloc(h, &h->adj_face, n_adj));
  A_TRY(dalloc(h, &h->adj_col, n_adj));
  A_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * ((size_t)nf_local + 1), st));
  if (n_faces) {
    k_adjacency<1><<<grid_for(n_faces), kBlock, 0, st>>>(n_faces, d_nb, res, d_cnt, h->adjptr, h->adj_face,
                                                         h->adj_col, nullptr, nullptr, h->halo_glob, h->n_halo,
                                                         nf_local, d_err);
    h->tm.kernel_launches++;
  }

  // ---- 4. row structure, 5. values ------------------------------------------------------------------
  ColKey key{nf_local, h->row_start, h->halo_glob};
  A_TRY(dalloc(h, &h->rowptr, (int64_t)nf_local + 1 + kRowptrPad));
  if (nf_local) {
    k_row_structure<<<grid_for(nf_local), kBlock, 0, st>>>(nf_local, h->adjptr, h->adj_face, h->adj_col, key, d_cnt);
    h->tm.kernel_launches++;
  }
  exclusive_scan(d_cnt, nf_local, h->rowptr, d_scratch, st, &h->tm.kernel_launches);
  k_fill_tail<<<1, kBlock, 0, st>>>(h->rowptr, nf_local, kRowptrPad);  // spare entries = nnz (empty rows)
  h->tm.kernel_launches++;
  int nnz = 0;
  int herr[ERR_COUNT];
  A_CUDA(cudaMemcpyAsync(&nnz, h->rowptr + nf_local, sizeof(int), cudaMemcpyDeviceToHost, st));
  A_CUDA(cudaMemcpyAsync(herr, d_err, sizeof(herr), cudaMemcpyDeviceToHost, st));
  A_CUDA(cudaStreamSynchronize(st));
  if (herr[ERR_SRC_ON_DIRICHLET] != INT_MAX) {
    int64_t node = 0;
    memcpy_sync(h->stream, &node, d_dnodes + herr[ERR_SRC_ON_DIRICHLET], sizeof(int64_t), cudaMemcpyDeviceToHost);
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "There cannot be a source at a Dirichlet node, but node " + std::to_string(node) +
                                            " is a Dirichlet node where a source is located.");
  }
  if (herr[ERR_BAD_NODE] != INT_MAX) {
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "node index out of range 1..N (neighbors or dirichletnodes entry " +
                                            std::to_string(herr[ERR_BAD_NODE] + 1) + ")");
  }
  if (herr[ERR_BAD_META] != INT_MAX) {
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "metaindex(" + std::to_string(herr[ERR_BAD_META] + 1) + ") is outside conductivities");
  }
  h->nnz = nnz;
  lap("row structure (sync)");
  A_TRY(dalloc(h, &h->colidx, (int64_t)nnz + kCsrPad));
  A_TRY(dalloc(h, &h->vals, (int64_t)nnz + kCsrPad));
  A_CUDA(cudaMemsetAsync(h->colidx + nnz, 0, sizeof(int) * kCsrPad, st));
  A_CUDA(cudaMemsetAsync(h->vals + nnz, 0, sizeof(double) * kCsrPad, st));
  A_TRY(dalloc(h, &h->b, nf_local));
  A_TRY(dalloc(h, &h->diag, nf_local));
  if (nf_local) {
    k_row_values<<<grid_for(nf_local), kBlock, 0, st>>>(nf_local, h->adjptr, h->adj_face, h->adj_col, key, h->cface,
                                                        h->sources, h->dheads, h->row2node, h->rowptr, h->colidx,
                                                        h->vals, h->diag, h->b, 1);
    h->tm.kernel_launches++;
  }
  if (h->fmt_request != 1) A_TRY(build_dia(h, true));
  if (h->precond_request == 1) A_TRY(precond_setup(h));
  A_CUDA(cudaEventRecord(h->ev[2], st));
  A_CUDA(cudaStreamSynchronize(st));
  A_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->tm.h2d_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->tm.assemble_ms = ms;
  lap("values + dia (sync)");
  cleanup();
  lap("free temporaries");
#undef A_TRY
#undef A_CUDA
  h->assembled = true;
  h->halo_ready = (h->nranks == 1);
  return FVB_OK;
}

int fvb_update_values(fvb_handle h, const double *cond, int64_t n_cond, int logk, const double *sources,
                      const double *dheads) {
  FVB_TRY(check_handle(h, true));
  if (!cond) return set_error(FVB_ERR_BAD_INPUT, "null conductivities");
  cudaStream_t st = h->stream;
  if (h->box_implicit) {
    // grid-implicit problem (fvb_assemble_regulargrid): `cond` are the node values of the same planes as before
    if (n_cond != h->nodek_n)
      return set_error(FVB_ERR_BAD_INPUT, "grid-implicit problem: pass the node conductivities of the same " +
                                              std::to_string(h->nodek_n) + " nodes as at assembly");
    if (sources && !h->sources) FVB_TRY(dalloc(h, &h->sources, h->n_own_nodes));
    FVB_CUDA(cudaMemcpyAsync(h->nodek, cond, sizeof(double) * (size_t)n_cond, cudaMemcpyDefault, st));
    if (sources) FVB_CUDA(cudaMemcpyAsync(h->sources, sources, sizeof(double) * (size_t)h->n_own_nodes, cudaMemcpyDefault, st));
    if (dheads) FVB_CUDA(cudaMemcpyAsync(h->dheads, dheads, sizeof(double) * (size_t)h->n_dirichlet, cudaMemcpyDefault, st));
    FVB_CUDA(cudaEventRecord(h->ev[1], st));
    h->logk = logk ? 1 : 0;
    FVB_TRY(box_fill(h));
    dfree(h, h->rowptr); dfree(h, h->colidx); dfree(h, h->vals);
    if (h->mg && h->mg->ready) FVB_TRY(mg_setup(h, false));
    FVB_CUDA(cudaEventRecord(h->ev[2], st));
    FVB_CUDA(cudaStreamSynchronize(st));
    float msi = 0;
    cudaEventElapsedTime(&msi, h->ev[1], h->ev[2]); h->tm.assemble_ms = msi;
    return FVB_OK;
  }
  double *d_cond = nullptr;
  int *d_err = nullptr;
  FVB_TRY(dalloc(h, &d_cond, n_cond));
  FVB_TRY(dalloc(h, &d_err, ERR_COUNT));
  int init[ERR_COUNT];
  for (int &v : init) v = INT_MAX;
  {
    cudaError_t ce = cudaMemcpyAsync(d_err, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_cond, cond, sizeof(double) * (size_t)n_cond, cudaMemcpyDefault, st);
    if (ce == cudaSuccess && sources)
      ce = cudaMemcpyAsync(h->sources, sources, sizeof(double) * (size_t)h->n_own_nodes, cudaMemcpyDefault, st);
    if (ce == cudaSuccess && dheads)
      ce = cudaMemcpyAsync(h->dheads, dheads, sizeof(double) * (size_t)h->n_dirichlet, cudaMemcpyDefault, st);
    if (ce != cudaSuccess) {
      dfree(h, d_cond); dfree(h, d_err);
      return set_error(FVB_ERR_CUDA, std::string("fvb_update_values upload: ") + cudaGetErrorString(ce));
    }
  }
  cudaEventRecord(h->ev[1], st);
  if (h->n_faces) {
    k_face_conductance<<<grid_for(h->n_faces), kBlock, 0, st>>>(h->n_faces, d_cond, n_cond, h->meta, h->aol, logk,
                                                                h->cface, d_err);
    h->tm.kernel_launches++;
  }
  ColKey key{(int)h->nf_local, h->row_start, h->halo_glob};
  if (h->nf_local && !h->box) {
    k_row_values<<<grid_for(h->nf_local), kBlock, 0, st>>>((int)h->nf_local, h->adjptr, h->adj_face, h->adj_col, key,
                                                           h->cface, h->sources, h->dheads, h->row2node, h->rowptr,
                                                           h->colidx, h->vals, h->diag, h->b, 0);
    h->tm.kernel_launches++;
  }
  h->logk = logk ? 1 : 0;  // the gradient gather picks dc = c (log K) or aol (plain K) from this
  int st_dia = FVB_OK, st_mg = FVB_OK;
  if (h->box) {
    st_dia = box_fill(h);
    dfree(h, h->rowptr); dfree(h, h->colidx); dfree(h, h->vals);  // a CSR image built on demand is stale now
  } else if (h->dia_on) st_dia = build_dia(h, false);
  if (st_dia == FVB_OK && h->mg && h->mg->ready) st_mg = mg_setup(h, false);
  if (st_dia == FVB_OK && st_mg == FVB_OK && h->amg && h->amg->ready) st_mg = amg_setup(h);  // aggregates follow the values
  cudaEventRecord(h->ev[2], st);
  int herr[ERR_COUNT];
  cudaError_t e = cudaMemcpyAsync(herr, d_err, sizeof(herr), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  dfree(h, d_cond); dfree(h, d_err);
  if (st_dia != FVB_OK) return st_dia;
  if (st_mg != FVB_OK) return st_mg;
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->tm.assemble_ms = ms;
  if (herr[ERR_BAD_META] != INT_MAX) return set_error(FVB_ERR_BAD_INPUT, "metaindex outside conductivities");
  return FVB_OK;
}

int fvb_assemble_regulargrid(fvb_handle h, const double mins[3], const double maxs[3], const int64_t ns[3],
                             int64_t plane_lo, int64_t plane_hi, const double *nodehycos, int logmean, int logk,
                             const double *sources, int64_t nd, const int64_t *dnodes, const double *dheads) {
  FVB_TRY(check_handle(h, false));
  if (!mins || !maxs || !ns || !nodehycos) return set_error(FVB_ERR_BAD_INPUT, "null grid description");
  if (ns[0] < 2 || ns[1] < 2 || ns[2] < 2) return set_error(FVB_ERR_BAD_INPUT, "regulargrid needs at least 2 points per axis");
  if (plane_lo < 1 || plane_hi > ns[0] || plane_hi < plane_lo) return set_error(FVB_ERR_BAD_INPUT, "bad plane range");
  if (nd < 0 || (nd && (!dnodes || !dheads))) return set_error(FVB_ERR_BAD_INPUT, "bad Dirichlet arguments");
  const int64_t plane = ns[1] * ns[2], n_nodes = ns[0] * plane;
  const int64_t n_own = (plane_hi - plane_lo + 1) * plane;
  if (n_own >= INT_MAX - 1 || nd >= INT_MAX - 1)
    return set_error(FVB_ERR_BAD_INPUT, "per-GPU part too large for 32-bit local indices; use more ranks");
  free_problem(h);
  cudaStream_t st = h->stream;
  h->n_nodes = n_nodes; h->node_lo = (plane_lo - 1) * plane; h->node_hi = plane_hi * plane; h->n_own_nodes = n_own;
  h->n_faces = 0; h->n_dirichlet = nd; h->logk = logk ? 1 : 0; h->box_logmean = logmean ? 1 : 0;
  const bool whole = plane_lo == 1 && plane_hi == ns[0];
  BoxDesc B = {};
  GridDesc &G = B.G;
  G.n1 = ns[0]; G.n2 = ns[1]; G.n3 = ns[2];
  auto axis = [](double lo, double hi, int64_t n) { double step = (hi - lo) / (double)(n - 1); return (lo + 1 * step) - lo; };
  G.dx = axis(mins[0], maxs[0], ns[0]); G.dy = axis(mins[1], maxs[1], ns[1]); G.dz = axis(mins[2], maxs[2], ns[2]);
  G.p_lo = plane_lo; G.p_hi = plane_hi; G.e_lo = plane_lo > 1 ? plane_lo - 1 : plane_lo;
  const int64_t k_lo_plane = std::max<int64_t>(1, plane_lo - 1), k_hi_plane = std::min<int64_t>(ns[0], plane_hi + 1);
  h->nodek_n = (k_hi_plane - k_lo_plane + 1) * plane;
  h->nodek_ofs = (plane_lo - k_lo_plane) * plane;

  int64_t *d_dnodes = nullptr;
  int *d_dslot = nullptr, *d_cnt = nullptr, *d_scratch = nullptr, *d_err = nullptr, *d_flag = nullptr;
  auto cleanup = [&]() { dfree(h, d_dnodes); dfree(h, d_dslot); dfree(h, d_cnt); dfree(h, d_scratch); dfree(h, d_err); dfree(h, d_flag); };
#define G_TRY(expr) do { int s__ = (expr); if (s__ != FVB_OK) { cleanup(); free_problem(h); return s__; } } while (0)
#define G_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); free_problem(h); \
    return set_error(FVB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  G_CUDA(cudaEventRecord(h->ev[0], st));
  G_TRY(dalloc(h, &h->nodek, h->nodek_n));
  G_TRY(dalloc(h, &d_dnodes, nd));
  G_TRY(dalloc(h, &h->dheads, nd));
  G_CUDA(cudaMemcpyAsync(h->nodek, nodehycos, sizeof(double) * (size_t)h->nodek_n, cudaMemcpyDefault, st));
  if (sources) {
    G_TRY(dalloc(h, &h->sources, n_own));
    G_CUDA(cudaMemcpyAsync(h->sources, sources, sizeof(double) * (size_t)n_own, cudaMemcpyDefault, st));
  }
  if (nd) {
    G_CUDA(cudaMemcpyAsync(d_dnodes, dnodes, sizeof(int64_t) * (size_t)nd, cudaMemcpyDefault, st));
    G_CUDA(cudaMemcpyAsync(h->dheads, dheads, sizeof(double) * (size_t)nd, cudaMemcpyDefault, st));
  }
  G_CUDA(cudaEventRecord(h->ev[1], st));
  // Dirichlet table of the whole problem: row offsets of a slab and the kind of the planes outside it
  std::vector<int64_t> nodes;
  h->row_start = 0;
  if (!whole) {
    std::vector<int64_t> hd((size_t)nd);
    std::vector<int> slot;
    if (nd) G_CUDA(memcpy_sync(st, hd.data(), d_dnodes, sizeof(int64_t) * (size_t)nd, cudaMemcpyDeviceToHost));
    dirichlet_table(hd, nodes, slot);
    h->nd_sorted = (int64_t)nodes.size();
    G_TRY(dalloc(h, &h->d_dsorted, h->nd_sorted));
    G_TRY(dalloc(h, &h->d_dsorted_slot, h->nd_sorted));
    G_CUDA(memcpy_sync(st, h->d_dsorted, nodes.data(), sizeof(int64_t) * nodes.size(), cudaMemcpyHostToDevice));
    G_CUDA(memcpy_sync(st, h->d_dsorted_slot, slot.data(), sizeof(int) * slot.size(), cudaMemcpyHostToDevice));
    const int64_t below = std::lower_bound(nodes.begin(), nodes.end(), h->node_lo) - nodes.begin();
    const int64_t valid = std::lower_bound(nodes.begin(), nodes.end(), n_nodes) - std::lower_bound(nodes.begin(), nodes.end(), (int64_t)0);
    h->row_start = h->node_lo - below;
    h->nf_global = n_nodes - valid;
  }
  // node map (getfreenodes / getnodei2dirichleti, src/FiniteVolume.jl:20-44), as in fvb_assemble
  G_TRY(dalloc(h, &d_err, ERR_COUNT));
  {
    int init[ERR_COUNT];
    for (int &v : init) v = INT_MAX;
    G_CUDA(cudaMemcpyAsync(d_err, init, sizeof(init), cudaMemcpyHostToDevice, st));
  }
  G_TRY(dalloc(h, &d_dslot, n_own));
  G_TRY(dalloc(h, &d_cnt, n_own + 2));
  G_TRY(dalloc(h, &d_scratch, scan_scratch_ints(n_own + 1)));
  G_CUDA(cudaMemsetAsync(d_dslot, 0xFF, sizeof(int) * (size_t)n_own, st));
  if (nd) {
    k_mark_dirichlet<<<grid_for(nd), kBlock, 0, st>>>(d_dnodes, nd, n_nodes, h->node_lo, h->node_hi, h->sources, d_dslot, d_err);
    h->tm.kernel_launches++;
  }
  int nf_local = 0;
  k_free_flags<<<grid_for(n_own), kBlock, 0, st>>>(d_dslot, n_own, d_cnt);
  h->tm.kernel_launches++;
  exclusive_scan(d_cnt, n_own, d_cnt, d_scratch, st, &h->tm.kernel_launches);
  G_CUDA(memcpy_sync(st, &nf_local, d_cnt + n_own, sizeof(int), cudaMemcpyDeviceToHost));
  h->nf_local = nf_local;
  if (whole) h->nf_global = nf_local;
  G_TRY(dalloc(h, &h->nodemap, n_own));
  G_TRY(dalloc(h, &h->row2node, nf_local));
  k_finish_nodemap<<<grid_for(n_own), kBlock, 0, st>>>(d_dslot, d_cnt, n_own, h->nodemap, h->row2node);
  h->tm.kernel_launches++;
  dfree(h, d_dslot); dfree(h, d_cnt); dfree(h, d_scratch);
  int herr[ERR_COUNT];
  G_CUDA(memcpy_sync(st, herr, d_err, sizeof(herr), cudaMemcpyDeviceToHost));
  if (herr[ERR_SRC_ON_DIRICHLET] != INT_MAX) {
    int64_t node = 0;
    memcpy_sync(st, &node, d_dnodes + herr[ERR_SRC_ON_DIRICHLET], sizeof(int64_t), cudaMemcpyDeviceToHost);
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "There cannot be a source at a Dirichlet node, but node " + std::to_string(node) +
                                            " is a Dirichlet node where a source is located.");
  }
  if (herr[ERR_BAD_NODE] != INT_MAX) {
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, "node index out of range 1..N (dirichletnodes entry " + std::to_string(herr[ERR_BAD_NODE] + 1) + ")");
  }
  // the diagonal format needs what box_detect checks for explicit lists
  B.n_own = n_own;
  B.nd_owned = n_own - nf_local;
  const char *why = nullptr;
  if (nf_local < 2) why = "fewer than two free nodes on this rank";
  if (!why && plane_lo > 1) {
    B.lo_kind = box_plane_kind(nodes, h->node_lo - plane, plane);
    if (B.lo_kind < 0 || (B.lo_kind == 1 && box_plane_kind(nodes, h->node_lo, plane) != 1))
      why = "the x-planes at the lower slab boundary mix free and Dirichlet nodes";
  }
  if (!why && plane_hi < ns[0]) {
    B.hi_kind = box_plane_kind(nodes, h->node_hi, plane);
    if (B.hi_kind < 0 || (B.hi_kind == 1 && box_plane_kind(nodes, h->node_hi - plane, plane) != 1))
      why = "the x-planes at the upper slab boundary mix free and Dirichlet nodes";
  }
  if (!why) {
    G_TRY(dalloc(h, &d_flag, 2));
    G_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), st));
    k_box_check_shift<<<std::max(1, std::min(cdiv(n_own, kBlock), h->num_sms * 16)), kBlock, 0, st>>>(B, h->nodemap, d_flag);
    h->tm.kernel_launches++;
    int flag[2] = {0, 0};
    G_CUDA(memcpy_sync(st, flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag[1]) why = "the Dirichlet set does not leave the free nodes on the grid's diagonals";
  }
  if (why) {
    cleanup(); free_problem(h);
    return set_error(FVB_ERR_BAD_INPUT, std::string("fvb_assemble_regulargrid: ") + why +
                                            " (build the face list with fvb_regulargrid and call fvb_assemble instead)");
  }
  h->box_implicit = true;
  G_TRY(box_install(h, B));
  if (h->precond_request == 1) G_TRY(precond_setup(h));
  G_CUDA(cudaEventRecord(h->ev[2], st));
  G_CUDA(cudaStreamSynchronize(st));
  G_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->tm.h2d_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->tm.assemble_ms = ms;
  cleanup();
#undef G_TRY
#undef G_CUDA
  h->assembled = true;
  h->halo_ready = (h->nranks == 1);
  return FVB_OK;
}

int fvb_sizes(fvb_handle h, int64_t *nf_local, int64_t *nnz_local, int64_t *row_start, int64_t *nf_global,
              int64_t *n_halo) {
  FVB_TRY(check_handle(h, true));
  if (nf_local) *nf_local = h->nf_local;
  if (nnz_local) *nnz_local = h->nnz;
  if (row_start) *row_start = h->row_start + 1;
  if (nf_global) *nf_global = h->nf_global;
  if (n_halo) *n_halo = h->n_halo;
  return FVB_OK;
}

int fvb_get_csr(fvb_handle h, int64_t *ptr, int64_t *idx, double *val) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_csr(h));  // box problems keep no CSR image until somebody asks for A
  cudaStream_t st = h->stream;
  ColKey key{(int)h->nf_local, h->row_start, h->halo_glob};
  // export through a bounded device staging buffer (the int64 image of colidx can be 7 GiB)
  const int64_t chunk = 1 << 24;
  int64_t *d_tmp = nullptr;
  FVB_TRY(dalloc(h, &d_tmp, chunk));
  if (ptr) {
    for (int64_t o = 0; o < h->nf_local + 1; o += chunk) {
      int64_t m = std::min(chunk, h->nf_local + 1 - o);
      k_export_ptr<<<grid_for(m), kBlock, 0, st>>>(h->rowptr + o, m, d_tmp);
      h->tm.kernel_launches++;
      FVB_CUDA(cudaMemcpyAsync(ptr + o, d_tmp, sizeof(int64_t) * (size_t)m, cudaMemcpyDefault, st));
      FVB_CUDA(cudaStreamSynchronize(st));
    }
  }
  if (idx) {
    for (int64_t o = 0; o < h->nnz; o += chunk) {
      int64_t m = std::min(chunk, h->nnz - o);
      k_export_cols<<<grid_for(m), kBlock, 0, st>>>(h->colidx + o, m, key, d_tmp);
      h->tm.kernel_launches++;
      FVB_CUDA(cudaMemcpyAsync(idx + o, d_tmp, sizeof(int64_t) * (size_t)m, cudaMemcpyDefault, st));
      FVB_CUDA(cudaStreamSynchronize(st));
    }
  }
  dfree(h, d_tmp);
  if (val) FVB_CUDA(memcpy_sync(h->stream, val, h->vals, sizeof(double) * (size_t)h->nnz, cudaMemcpyDefault));
  return FVB_OK;
}

int fvb_get_b(fvb_handle h, double *b) {
  FVB_TRY(check_handle(h, true));
  FVB_CUDA(memcpy_sync(h->stream, b, h->b, sizeof(double) * (size_t)h->nf_local, cudaMemcpyDefault));
  return FVB_OK;
}
int fvb_get_diag(fvb_handle h, double *d) {
  FVB_TRY(check_handle(h, true));
  FVB_CUDA(memcpy_sync(h->stream, d, h->diag, sizeof(double) * (size_t)h->nf_local, cudaMemcpyDefault));
  return FVB_OK;
}

static int export_nodemap(fvb_handle h, uint8_t *freenode, int64_t *n2f) {
  FVB_TRY(check_handle(h, true));
  const int64_t n = h->n_own_nodes;
  uint8_t *d_f = nullptr;
  int64_t *d_m = nullptr;
  if (freenode) FVB_TRY(dalloc(h, &d_f, n));
  if (n2f) FVB_TRY(dalloc(h, &d_m, n));
  if (n) {
    k_export_nodemap<<<grid_for(n), kBlock, 0, h->stream>>>(h->nodemap, n, h->row_start, d_f, d_m);
    h->tm.kernel_launches++;
  }
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && freenode) e = memcpy_sync(h->stream, freenode, d_f, (size_t)n, cudaMemcpyDefault);
  if (e == cudaSuccess && n2f) e = memcpy_sync(h->stream, n2f, d_m, sizeof(int64_t) * (size_t)n, cudaMemcpyDefault);
  dfree(h, d_f); dfree(h, d_m);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  return FVB_OK;
}
int fvb_get_freenode(fvb_handle h, uint8_t *freenode) { return export_nodemap(h, freenode, nullptr); }
int fvb_get_nodei2freenodei(fvb_handle h, int64_t *map) { return export_nodemap(h, nullptr, map); }

int fvb_get_halo_cols(fvb_handle h, int64_t *cols) {
  FVB_TRY(check_handle(h, true));
  for (int64_t k = 0; k < h->n_halo; ++k) cols[k] = h->halo_host[(size_t)k] + 1;
  return FVB_OK;
}

int fvb_set_halo_plan(fvb_handle h, int n_peers, const int32_t *peer_ranks, const int64_t *send_counts,
                      const int32_t *send_rows, const int64_t *recv_counts) {
  FVB_TRY(check_handle(h, true));
  if (n_peers < 0) return set_error(FVB_ERR_BAD_INPUT, "negative peer count");
  int64_t ns = 0, nr = 0;
  for (int p = 0; p < n_peers; ++p) { ns += send_counts[p]; nr += recv_counts[p]; }
  if (nr != h->n_halo) return set_error(FVB_ERR_BAD_INPUT, "recv counts do not add up to the halo size");
  std::vector<int32_t> rows(send_rows, send_rows + ns);
  for (int32_t r : rows)
    if (r < 0 || r >= h->nf_local) return set_error(FVB_ERR_BAD_INPUT, "send row outside the owned range");
  h->peers.assign(peer_ranks, peer_ranks + n_peers);
  h->send_counts.assign(send_counts, send_counts + n_peers);
  h->recv_counts.assign(recv_counts, recv_counts + n_peers);
  h->send_first.assign((size_t)n_peers, 0);
  h->send_contig = true;
  {
    int64_t at = 0;
    for (int p = 0; p < n_peers; ++p) {
      if (send_counts[p] > 0) h->send_first[(size_t)p] = rows[(size_t)at];
      for (int64_t k = 1; k < send_counts[p] && h->send_contig; ++k)
        h->send_contig = rows[(size_t)(at + k)] == rows[(size_t)(at + k - 1)] + 1;
      at += send_counts[p];
    }
  }
  dfree(h, h->send_rows); dfree(h, h->sendbuf);
  FVB_TRY(dalloc(h, &h->send_rows, ns));
  FVB_TRY(dalloc(h, &h->sendbuf, ns));
  FVB_CUDA(memcpy_sync(h->stream, h->send_rows, rows.data(), sizeof(int32_t) * (size_t)ns, cudaMemcpyHostToDevice));
  h->n_send = ns;
  h->halo_ready = true;
  return FVB_OK;
}

int fvb_gradient_begin(fvb_handle h, const int64_t *neighbors) {
  FVB_TRY(check_handle(h, true));
  if (h->box_implicit)
    return set_error(FVB_ERR_STATE, "the gradient gather needs per-face arrays: assemble with fvb_assemble, not fvb_assemble_regulargrid");
  if (h->n_faces && !neighbors) return set_error(FVB_ERR_BAD_INPUT, "null neighbors");
  if (h->nranks != 1 || h->node_lo != 0 || h->node_hi != h->n_nodes)
    return set_error(FVB_ERR_STATE, "the gradient gather is implemented for unpartitioned problems");
  cudaStream_t st = h->stream;
  const int64_t F = h->n_faces, n = h->nf_local;
  if (!h->g_e1) {
    FVB_TRY(dalloc(h, &h->g_e1, F)); FVB_TRY(dalloc(h, &h->g_e2, F));
    FVB_TRY(dalloc(h, &h->g_face, F)); FVB_TRY(dalloc(h, &h->g_dh, F)); FVB_TRY(dalloc(h, &h->g_src, n));
    int64_t *d_nb = nullptr;
    FVB_TRY(dalloc(h, &d_nb, 2 * F));
    FVB_CUDA(cudaMemcpyAsync(d_nb, neighbors, sizeof(int64_t) * 2 * (size_t)F, cudaMemcpyDefault, st));
    Resolver res{h->nodemap, h->n_nodes, h->node_lo, h->node_hi, nullptr, nullptr, 0};
    if (F) {
      k_face_endpoints<<<grid_for(F), kBlock, 0, st>>>(F, d_nb, res, h->g_e1, h->g_e2);
      h->tm.kernel_launches++;
    }
    dfree(h, d_nb);
  }
  FVB_CUDA(cudaMemsetAsync(h->g_face, 0, sizeof(double) * (size_t)std::max<int64_t>(F, 1), st));
  FVB_CUDA(cudaMemsetAsync(h->g_dh, 0, sizeof(double) * (size_t)std::max<int64_t>(F, 1), st));
  FVB_CUDA(cudaMemsetAsync(h->g_src, 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1), st));
  FVB_CUDA(cudaStreamSynchronize(st));
  return FVB_OK;
}

int fvb_gradient_accumulate(fvb_handle h, int u_slot, int lambda_slot, double weight) {
  FVB_TRY(check_handle(h, true));
  if (!h->g_e1) return set_error(FVB_ERR_STATE, "call fvb_gradient_begin first");
  FVB_TRY(ensure_slot(h, u_slot));
  FVB_TRY(ensure_slot(h, lambda_slot));
  cudaStream_t st = h->stream;
  if (h->n_faces) {
    k_gradient_faces<<<grid_for(h->n_faces), kBlock, 0, st>>>(h->n_faces, h->g_e1, h->g_e2, h->cface, h->aol, h->logk,
                                                              h->slots[u_slot], h->slots[lambda_slot], h->Dvec,
                                                              h->dheads, weight, h->g_face, h->g_dh);
    h->tm.kernel_launches++;
  }
  if (h->nf_local) {
    k_gradient_rows<<<grid_for(h->nf_local), kBlock, 0, st>>>(h->nf_local, h->slots[lambda_slot], h->Dvec, weight,
                                                              h->g_src);
    h->tm.kernel_launches++;
  }
  return FVB_OK;
}

int fvb_gradient_end(fvb_handle h, double *grad_cond_face, double *grad_dhead_face, int64_t *dhead_slot_face,
                     double *grad_source_rows) {
  FVB_TRY(check_handle(h, true));
  if (!h->g_e1) return set_error(FVB_ERR_STATE, "call fvb_gradient_begin first");
  cudaStream_t st = h->stream;
  const size_t F = (size_t)h->n_faces;
  if (grad_cond_face) FVB_CUDA(cudaMemcpyAsync(grad_cond_face, h->g_face, sizeof(double) * F, cudaMemcpyDefault, st));
  if (grad_dhead_face) FVB_CUDA(cudaMemcpyAsync(grad_dhead_face, h->g_dh, sizeof(double) * F, cudaMemcpyDefault, st));
  if (grad_source_rows)
    FVB_CUDA(cudaMemcpyAsync(grad_source_rows, h->g_src, sizeof(double) * (size_t)h->nf_local, cudaMemcpyDefault, st));
  if (dhead_slot_face) {
    int64_t *d_s = nullptr;
    FVB_TRY(dalloc(h, &d_s, h->n_faces));
    if (F) {
      k_gradient_dslots<<<grid_for(h->n_faces), kBlock, 0, st>>>(h->n_faces, h->g_e1, h->g_e2, d_s);
      h->tm.kernel_launches++;
    }
    cudaError_t e = cudaMemcpyAsync(dhead_slot_face, d_s, sizeof(int64_t) * F, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dfree(h, d_s);
    if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  }
  FVB_CUDA(cudaStreamSynchronize(st));
  return FVB_OK;
}

int fvb_peer_export(fvb_handle h, uint8_t blob[FVB_PEER_BLOB_BYTES]) {
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == FVB_PEER_BLOB_BYTES, "blob = two IPC handles");
  FVB_TRY(check_handle(h, true));
  if (h->nranks < 2 || h->nranks > kMaxRanks) return set_error(FVB_ERR_STATE, "peer exchange needs 2..8 ranks (fvb_comm_init)");
  FVB_TRY(ensure_workspace(h));
  if (!h->peer) h->peer = new PeerState();
  PeerState &P = *h->peer;
  P.active = false;
  if (!P.mail) {
    FVB_CUDA(cudaMalloc((void **)&P.mail, sizeof(PeerMail)));
    FVB_CUDA(cudaMemsetAsync(P.mail, 0, sizeof(PeerMail), h->stream));
  }
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  cudaIpcMemHandle_t hu, hm;
  FVB_CUDA(cudaIpcGetMemHandle(&hu, h->u));
  FVB_CUDA(cudaIpcGetMemHandle(&hm, P.mail));
  memcpy(blob, &hu, sizeof(hu));
  memcpy(blob + sizeof(hu), &hm, sizeof(hm));
  return FVB_OK;
}

int fvb_peer_import(fvb_handle h, const uint8_t *blobs, const int64_t *send_dst_index) {
  FVB_TRY(check_handle(h, true));
  if (!h->peer || !h->peer->mail) return set_error(FVB_ERR_STATE, "call fvb_peer_export first");
  if (!h->halo_ready) return set_error(FVB_ERR_STATE, "call fvb_set_halo_plan first");
  PeerState &P = *h->peer;
  // Re-assembly of a same-sized problem exports the very same allocations: keep the mappings
  // (cudaIpcOpenMemHandle costs ~0.1 s per peer) and only rebuild the plan tables.
  std::vector<uint8_t> nb(blobs, blobs + (size_t)h->nranks * FVB_PEER_BLOB_BYTES);
  bool reuse = nb == P.last_blobs;
  for (size_t p = 0; reuse && p < h->peers.size(); ++p) reuse = P.tab.u[h->peers[p]] != nullptr;
  if (!reuse) {
    for (int r = 0; r < kMaxRanks; ++r) {
      if (P.opened_u[r]) { cudaIpcCloseMemHandle(P.opened_u[r]); P.opened_u[r] = nullptr; }
      if (P.opened_mail[r]) { cudaIpcCloseMemHandle(P.opened_mail[r]); P.opened_mail[r] = nullptr; }
    }
    P.tab = PeerTable{};
    P.last_blobs.clear();
  }
  P.tab.nranks = h->nranks;
  P.tab.rank = h->rank;
  for (int r = 0; r < h->nranks && !reuse; ++r) {
    if (r == h->rank) { P.tab.u[r] = h->u; P.tab.mail[r] = P.mail; continue; }
    cudaIpcMemHandle_t hu, hm;
    memcpy(&hu, blobs + (size_t)r * FVB_PEER_BLOB_BYTES, sizeof(hu));
    memcpy(&hm, blobs + (size_t)r * FVB_PEER_BLOB_BYTES + sizeof(hu), sizeof(hm));
    // only neighbours' vectors are written; every rank's mailbox is
    bool is_peer = std::find(h->peers.begin(), h->peers.end(), r) != h->peers.end();
    if (is_peer) {
      FVB_CUDA(cudaIpcOpenMemHandle(&P.opened_u[r], hu, cudaIpcMemLazyEnablePeerAccess));
      P.tab.u[r] = (double *)P.opened_u[r];
    }
    FVB_CUDA(cudaIpcOpenMemHandle(&P.opened_mail[r], hm, cudaIpcMemLazyEnablePeerAccess));
    P.tab.mail[r] = (PeerMail *)P.opened_mail[r];
  }
  P.last_blobs = nb;
  return peer_finish_plan(h, send_dst_index);
}

int fvb_solve(fvb_handle h, double rtol, int64_t maxiter, const double *x0_free, double *head_nodes, double *x_free,
              int64_t *iters, int *converged, double *resnorm_hist, int64_t hist_cap) {
  FVB_TRY(check_handle(h, true));
  if (maxiter < 0) return set_error(FVB_ERR_BAD_INPUT, "maxiter must be >= 0");
  FVB_TRY(ensure_workspace(h));
  cudaStream_t st = h->stream;
  const int64_t n = h->nf_local;
  FVB_CUDA(cudaEventRecord(h->ev[3], st));
  if (x0_free) FVB_CUDA(cudaMemcpyAsync(h->x, x0_free, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
  int64_t it = 0;
  int conv = 0;
  if (use_geometric_mg(h) || use_amg(h))
    FVB_TRY(pcg_mg_run(h, h->b, x0_free != nullptr, rtol, maxiter, &it, &conv));
  else
    FVB_TRY(pcg_run(h, h->b, x0_free != nullptr, 0.0, rtol, maxiter, &it, &conv));
  FVB_CUDA(cudaEventRecord(h->ev[4], st));
  if (head_nodes) {
    double *d_head = nullptr;
    FVB_TRY(dalloc(h, &d_head, h->n_own_nodes));
    if (h->n_own_nodes) {
      k_scatter_heads<<<grid_for(h->n_own_nodes), kBlock, 0, st>>>(h->nodemap, h->n_own_nodes, h->x, h->dheads, d_head);
      h->tm.kernel_launches++;
    }
    cudaError_t e = cudaMemcpyAsync(head_nodes, d_head, sizeof(double) * (size_t)h->n_own_nodes, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dfree(h, d_head);
    if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  }
  if (x_free) FVB_CUDA(cudaMemcpyAsync(x_free, h->x, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
  if (resnorm_hist && hist_cap > 0) {
    int64_t m = std::min<int64_t>(std::min<int64_t>(it, hist_cap), h->hist_cap);
    if (m > 0) FVB_CUDA(cudaMemcpyAsync(resnorm_hist, h->hist, sizeof(double) * (size_t)m, cudaMemcpyDefault, st));
  }
  FVB_CUDA(cudaEventRecord(h->ev[5], st));
  FVB_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->tm.solve_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->tm.d2h_ms = ms;
  if (iters) *iters = it;
  if (converged) *converged = conv;
  return FVB_OK;
}

int fvb_spmv(fvb_handle h, double alpha, const double *x, double beta, double *y) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_workspace(h));
  cudaStream_t st = h->stream;
  const int64_t n = h->nf_local;
  if (!h->yio) FVB_TRY(dalloc(h, &h->yio, n));
  FVB_CUDA(cudaMemcpyAsync(h->u, x, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
  if (beta != 0.0) FVB_CUDA(cudaMemcpyAsync(h->yio, y, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
  FVB_TRY(launch_spmv(h, h->u, h->c, 0.0, false));
  k_axpby<<<vgrid(h, n), kBlock, 0, st>>>(n, alpha, h->c, beta, h->yio);
  h->tm.kernel_launches++;
  FVB_CUDA(cudaMemcpyAsync(y, h->yio, sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
  FVB_CUDA(cudaStreamSynchronize(st));
  return FVB_OK;
}

// ---- transient --------------------------------------------------------------------------------------
int fvb_vec_upload(fvb_handle h, int slot, const double *host) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, slot));
  FVB_CUDA(cudaMemcpyAsync(h->slots[slot], host, sizeof(double) * (size_t)h->nf_local, cudaMemcpyDefault, h->stream));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  return FVB_OK;
}
int fvb_vec_download(fvb_handle h, int slot, double *host) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, slot));
  FVB_CUDA(cudaMemcpyAsync(host, h->slots[slot], sizeof(double) * (size_t)h->nf_local, cudaMemcpyDefault, h->stream));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  return FVB_OK;
}
int fvb_vec_copy(fvb_handle h, int dst, int src) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, dst));
  FVB_TRY(ensure_slot(h, src));
  if (dst != src)
    FVB_CUDA(cudaMemcpyAsync(h->slots[dst], h->slots[src], sizeof(double) * (size_t)h->nf_local,
                             cudaMemcpyDeviceToDevice, h->stream));
  return FVB_OK;
}
int fvb_vec_load_b(fvb_handle h, int slot) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, slot));
  FVB_CUDA(cudaMemcpyAsync(h->slots[slot], h->b, sizeof(double) * (size_t)h->nf_local, cudaMemcpyDeviceToDevice, h->stream));
  return FVB_OK;
}
int fvb_vec_diffnorm(fvb_handle h, int a, int b, double *out) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, a));
  FVB_TRY(ensure_slot(h, b));
  FVB_TRY(ensure_workspace(h));
  const int64_t n = h->nf_local;
  k_diffnorm2<<<vgrid(h, n), kBlock, 0, h->stream>>>(n, h->slots[a], h->slots[b], h->partials, h->ticket, h->scal->red);
  h->tm.kernel_launches++;
  FVB_TRY(allreduce_fin(h, 1, FIN_NONE));
  double s = 0;
  FVB_CUDA(cudaMemcpyAsync(&s, h->scal->red, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  *out = std::sqrt(s);
  return FVB_OK;
}
int fvb_set_storage(fvb_handle h, double Ss, const double *volumes) {
  FVB_TRY(check_handle(h, true));
  if (!volumes) { dfree(h, h->Dvec); return FVB_OK; }
  double *d_vol = nullptr;
  FVB_TRY(dalloc(h, &d_vol, h->n_own_nodes));
  if (!h->Dvec) FVB_TRY(dalloc(h, &h->Dvec, h->nf_local));
  cudaMemcpyAsync(d_vol, volumes, sizeof(double) * (size_t)h->n_own_nodes, cudaMemcpyDefault, h->stream);
  k_make_D<<<vgrid(h, h->nf_local), kBlock, 0, h->stream>>>(h->nf_local, h->row2node, d_vol, Ss, h->Dvec);
  h->tm.kernel_launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  dfree(h, d_vol);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  return FVB_OK;
}
int fvb_step(fvb_handle h, int rhs_slot, int u_slot, double dt, int out_slot, int adjoint, double rtol,
             int64_t maxiter, int64_t *iters, int *converged) {
  FVB_TRY(check_handle(h, true));
  if (!(dt > 0)) return set_error(FVB_ERR_BAD_INPUT, "time step must be positive");
  FVB_TRY(ensure_slot(h, rhs_slot));
  FVB_TRY(ensure_slot(h, u_slot));
  FVB_TRY(ensure_slot(h, out_slot));
  FVB_TRY(ensure_workspace(h));
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  const int vg = vgrid(h, n);
  const double sigma = 1.0 / dt;
  if (!h->rhs) FVB_TRY(dalloc(h, &h->rhs, n));
  if (!adjoint) {
    k_axpby_D<<<vg, kBlock, 0, st>>>(n, h->slots[rhs_slot], h->slots[u_slot], h->Dvec, sigma, h->rhs);
    FVB_CUDA(cudaMemcpyAsync(h->x, h->slots[u_slot], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  } else {
    k_axpby_D<<<vg, kBlock, 0, st>>>(n, h->slots[rhs_slot], h->slots[u_slot], nullptr, sigma, h->rhs);
    k_scale_D<<<vg, kBlock, 0, st>>>(n, h->slots[u_slot], h->Dvec, 0, h->x);
    h->tm.kernel_launches++;
  }
  h->tm.kernel_launches++;
  FVB_TRY(pcg_run(h, h->rhs, true, sigma, rtol, maxiter, iters, converged));
  if (!adjoint) {
    FVB_CUDA(cudaMemcpyAsync(h->slots[out_slot], h->x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  } else {
    k_scale_D<<<vg, kBlock, 0, st>>>(n, h->x, h->Dvec, 1, h->slots[out_slot]);
    h->tm.kernel_launches++;
  }
  return FVB_OK;
}
int fvb_solve_shifted(fvb_handle h, int rhs_slot, int x0_slot, double sigma, int out_slot, double rtol, int64_t maxiter,
                      int64_t *iters, int *converged) {
  FVB_TRY(check_handle(h, true));
  if (!(sigma >= 0)) return set_error(FVB_ERR_BAD_INPUT, "shift must be non-negative");
  FVB_TRY(ensure_slot(h, rhs_slot));
  FVB_TRY(ensure_slot(h, x0_slot));
  FVB_TRY(ensure_slot(h, out_slot));
  FVB_TRY(ensure_workspace(h));
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  FVB_CUDA(cudaMemcpyAsync(h->x, h->slots[x0_slot], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  FVB_TRY(pcg_run(h, h->slots[rhs_slot], true, sigma, rtol, maxiter, iters, converged));
  FVB_CUDA(cudaMemcpyAsync(h->slots[out_slot], h->x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  return FVB_OK;
}

int fvb_vec_to_nodes(fvb_handle h, int slot, double *head_nodes) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_slot(h, slot));
  double *d_head = nullptr;
  FVB_TRY(dalloc(h, &d_head, h->n_own_nodes));
  if (h->n_own_nodes) {
    k_scatter_heads<<<grid_for(h->n_own_nodes), kBlock, 0, h->stream>>>(h->nodemap, h->n_own_nodes, h->slots[slot],
                                                                        h->dheads, d_head);
    h->tm.kernel_launches++;
  }
  cudaError_t e = cudaMemcpyAsync(head_nodes, d_head, sizeof(double) * (size_t)h->n_own_nodes, cudaMemcpyDefault, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dfree(h, d_head);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  return FVB_OK;
}

// ---- device-resident integrator (src/transient.jl:78-154) ----------------------------------------------------------
}  // extern "C" (helpers below are internal)

namespace {

// The step controller of the reference -- backwardeulertwostep! (:78-87), adaptivebackwardeulerstep! (:89-121),
// fixedbackwardeulerstep! (:130-134) and the outer loop (:136-154) -- driving device-resident vectors.  Small
// single-GPU systems run every step-doubling attempt (up to three solves + the error norm) as ONE cooperative
// kernel (coop.cuh); everything else goes through fvb_step / fvb_vec_diffnorm, one call per solve.
struct Integrator {
  fvb_handle h;
  const fvb_integrate_options *opt;
  bool coop = false;
  int coop_grid = 0;
  double *work[6] = {};      // x, r, p, c, dinv, rhs of the cooperative kernel
  double *coop_partials = nullptr, *coop_result = nullptr;
  double *result_host = nullptr;
  std::vector<int> free_slots;
  int b_slot = 0;
  int64_t solves = 0, cg_its = 0, attempts = 0;
  bool all_converged = true;

  int take() { int s = free_slots.back(); free_slots.pop_back(); return s; }
  void give(int s) { if (s >= 0) free_slots.push_back(s); }

  int load_b(double t) {
    if (!opt->getb) return FVB_OK;  // constant b, loaded once by the caller
    std::vector<double> hb((size_t)std::max<int64_t>(h->nf_local, 1));
    opt->getb(t, hb.data(), opt->getb_ctx);
    return fvb_vec_upload(h, b_slot, hb.data());
  }

  // one backward-Euler solve through the general path
  int onestep(int u, double t, double dt, int out) {
    if (!(dt > 0)) return set_error(FVB_ERR_BAD_INPUT, "time step must be positive");
    FVB_TRY(load_b(t));
    int64_t it = 0;
    int conv = 0;
    FVB_TRY(fvb_step(h, b_slot, u, dt, out, opt->adjoint, opt->rtol, opt->maxiter, &it, &conv));
    ++solves; cg_its += it; all_converged = all_converged && conv;
    return FVB_OK;
  }

  int launch_coop(int nsolves, const CoopSolve *sv, int na, int nb, double *err) {
    CoopJob J = {};
    J.nsolves = nsolves;
    for (int q = 0; q < nsolves; ++q) J.s[q] = sv[q];
    J.na = na >= 0 ? h->slots[na] : nullptr;
    J.nb = nb >= 0 ? h->slots[nb] : nullptr;
    J.adjoint = opt->adjoint;
    J.rtol = opt->rtol;
    J.maxiter = opt->maxiter;
    J.n = (int)h->nf_local;
    if (h->dia_on && h->fmt_request != 1) {
      J.D.K = h->dia_K;
      for (int k = 0; k < kDiaMaxOff; ++k) { J.D.off[k] = h->dia_off[k]; J.D.U[k] = h->dia_U[k]; }
      J.D.diag = h->diag; J.D.row_start = h->row_start; J.D.nf = h->nf_local;
    } else {
      J.rowptr = h->rowptr; J.colidx = h->colidx; J.vals = h->vals;
    }
    J.diag = h->diag; J.Dvec = h->Dvec;
    J.x = work[0]; J.r = work[1]; J.p = work[2]; J.c = work[3]; J.dinv = work[4]; J.rhs = work[5];
    J.partials = coop_partials;
    J.result = coop_result;
    void *args[] = {&J};
    FVB_CUDA(cudaLaunchCooperativeKernel((void *)k_coop_attempt, dim3((unsigned)coop_grid), dim3(kBlock), args, 0, h->stream));
    h->tm.kernel_launches++;
    FVB_CUDA(cudaMemcpyAsync(result_host, coop_result, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FVB_CUDA(cudaStreamSynchronize(h->stream));
    if (err) *err = result_host[0];
    solves += nsolves;
    cg_its += (int64_t)result_host[1];
    all_converged = all_converged && result_host[2] != 0.0;
    return FVB_OK;
  }

  // backwardeulertwostep!(..., onestep = given or computed): returns the new state's slot (*res), the time actually
  // stepped and whether the step size may grow.  `one` < 0: compute the full step here.  Consumes `one`.
  int twostep(int u, double t, double dt, int one, int *res, double *laststep, bool *increase) {
    if (!(dt > 0)) return set_error(FVB_ERR_BAD_INPUT, "time step must be positive");
    ++attempts;
    const int h1 = take(), two = take();
    double err = 0.0;
    if (coop) {
      CoopSolve sv[3];
      int q = 0;
      if (one < 0) {
        one = take();
        sv[q++] = CoopSolve{h->slots[b_slot], h->slots[u], h->slots[one], 1.0 / dt};
      }
      sv[q++] = CoopSolve{h->slots[b_slot], h->slots[u], h->slots[h1], 1.0 / (0.5 * dt)};
      sv[q++] = CoopSolve{h->slots[b_slot], h->slots[h1], h->slots[two], 1.0 / (0.5 * dt)};
      FVB_TRY(launch_coop(q, sv, one, two, &err));
    } else {
      if (one < 0) {
        one = take();
        FVB_TRY(onestep(u, t, dt, one));
      }
      FVB_TRY(onestep(u, t, 0.5 * dt, h1));
      FVB_TRY(onestep(h1, t + 0.5 * dt, 0.5 * dt, two));
      FVB_TRY(fvb_vec_diffnorm(h, one, two, &err));
    }
    give(one);
    if (err < opt->atol) {
      give(h1);
      *res = two; *laststep = dt; *increase = err < opt->atol / 4;
    } else {
      give(two);
      *res = h1; *laststep = 0.5 * dt; *increase = false;
    }
    return FVB_OK;
  }

  // adaptivebackwardeulerstep! (:89-121).  u stays owned by the caller; *res is a fresh slot.
  int adaptive(int u, double t, double dt, int *res, double *laststep, bool *increase) {
    if (opt->callback) opt->callback(t, dt, opt->callback_ctx);
    int u_new = -1;
    FVB_TRY(twostep(u, t, dt, -1, &u_new, laststep, increase));
    if (*laststep < dt) {  // it could not take the step we asked: cover dt with smaller ones
      bool failed = true;
      double elapsed = 0.0, target = *laststep;
      int u_el = u;  // not owned while it equals u
      while (elapsed < dt) {
        if (opt->callback) opt->callback(t, dt, opt->callback_ctx);
        int nxt = -1;
        // after a failed attempt the half step it returned serves as the full step of the next one (:100-101)
        FVB_TRY(twostep(u_el, t + elapsed, target, failed ? u_new : -1, &nxt, laststep, increase));
        if (!failed) give(u_new);
        u_new = nxt;
        if (*laststep == target) {
          elapsed += *laststep;
          if (u_el != u) give(u_el);
          // the accepted state is both the new starting point and (if the loop ends here) the result: keep one
          // slot for each role
          u_el = take();
          FVB_TRY(fvb_vec_copy(h, u_el, u_new));
          if (*increase) target = 2 * *laststep;
          failed = false;
        } else if (*laststep < target) {
          target = *laststep;
          failed = true;
        } else {
          return set_error(FVB_ERR_STATE, "Code is broken -- laststeptime should never be greater than targetdt");
        }
        target = std::min(target, dt - elapsed);
      }
      if (u_el != u) give(u_el);
    }
    *res = u_new;
    return FVB_OK;
  }
};

}  // namespace

extern "C" {

int fvb_integrate(fvb_handle h, const double *u0_free, double t0, double tfinal, const fvb_integrate_options *opt,
                  int64_t max_states, double *ts, double *us_free, double *heads_nodes, int64_t *n_states,
                  int64_t *n_solves, int64_t *n_cg_iterations, int64_t *n_attempts) {
  FVB_TRY(check_handle(h, true));
  if (!opt || !u0_free || !ts || max_states < 1) return set_error(FVB_ERR_BAD_INPUT, "bad arguments");
  if (!(opt->dt0 > 0)) return set_error(FVB_ERR_BAD_INPUT, "time step must be positive");
  FVB_TRY(ensure_workspace(h));
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  Integrator I;
  I.h = h;
  I.opt = opt;
  for (int s = FVB_NSLOT - 1; s >= 1; --s) I.free_slots.push_back(s);
  for (int s = 0; s < FVB_NSLOT; ++s) FVB_TRY(ensure_slot(h, s));
  I.b_slot = 0;
  if (!opt->getb) {
    if (opt->adjoint) FVB_CUDA(cudaMemsetAsync(h->slots[0], 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1), st));
    else FVB_TRY(fvb_vec_load_b(h, 0));
  }
  // the one-launch-per-attempt path: single GPU, small system, constant right-hand side
  const bool dia = h->dia_on && h->fmt_request != 1;
  I.coop = h->nranks == 1 && n >= 1 && n <= kCoopMaxRows && !opt->getb && !getenv("FVB_COOP_OFF") && (dia || h->rowptr || h->box);
  if (I.coop) {
    if (!dia) FVB_TRY(ensure_csr(h));
    int coop_ok = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, h->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_coop_attempt, kBlock, 0);
    I.coop = coop_ok && per_sm >= 1;
    I.coop_grid = std::max(1, std::min(cdiv(n, kBlock), h->num_sms));
  }
  auto release = [&]() {
    for (auto &w : I.work) dfree(h, w);
    dfree(h, I.coop_partials); dfree(h, I.coop_result);
    if (I.result_host) cudaFreeHost(I.result_host);
    I.result_host = nullptr;
  };
  if (I.coop) {
    int s = FVB_OK;
    for (auto &w : I.work) if (s == FVB_OK) s = dalloc(h, &w, n);
    if (s == FVB_OK) s = dalloc(h, &I.coop_partials, 3 * (int64_t)I.coop_grid);
    if (s == FVB_OK) s = dalloc(h, &I.coop_result, 4);
    if (s == FVB_OK && cudaMallocHost((void **)&I.result_host, 4 * sizeof(double)) != cudaSuccess) s = set_error(FVB_ERR_OOM, "pinned result buffer");
    if (s != FVB_OK) { release(); return s; }
  }
  auto fail = [&](int s) { std::string keep = g_last_error; release(); g_last_error = keep; return s; };
  // us = [u0], ts = [t0]   (:137-138)
  int u = I.take();
  int s = fvb_vec_upload(h, u, u0_free);
  if (s != FVB_OK) return fail(s);
  int64_t count = 0;
  auto store = [&](int slot, double t) -> int {
    if (count >= max_states) return set_error(FVB_ERR_STATE, "more accepted states than max_states (" + std::to_string(max_states) + "): enlarge the output buffers");
    ts[count] = t;
    if (us_free) FVB_CUDA(cudaMemcpyAsync(us_free + count * n, h->slots[slot], sizeof(double) * (size_t)n, cudaMemcpyDefault, st));
    if (heads_nodes) FVB_TRY(fvb_vec_to_nodes(h, slot, heads_nodes + count * h->n_own_nodes));
    ++count;
    return FVB_OK;
  };
  if ((s = store(u, t0)) != FVB_OK) return fail(s);
  double t = t0, dt = std::min(opt->dt0, tfinal - t0);
  while (t < tfinal) {
    int nxt = -1;
    double last = 0.0;
    bool inc = false;
    if (opt->fixed_step) {
      if (opt->callback) opt->callback(t, dt, opt->callback_ctx);
      nxt = I.take();
      if (I.coop) {
        CoopSolve sv{h->slots[I.b_slot], h->slots[u], h->slots[nxt], 1.0 / dt};
        s = dt > 0 ? I.launch_coop(1, &sv, -1, -1, nullptr) : set_error(FVB_ERR_BAD_INPUT, "time step must be positive");
      } else {
        s = I.onestep(u, t, dt, nxt);
      }
      last = dt;
    } else {
      s = I.adaptive(u, t, dt, &nxt, &last, &inc);
    }
    if (s != FVB_OK) return fail(s);
    I.give(u);
    u = nxt;
    t += dt;  // push!(ts, ts[end] + dt)   (:145)
    if ((s = store(u, t)) != FVB_OK) return fail(s);
    dt = inc ? std::min(tfinal - t, 2 * last) : std::min(tfinal - t, last);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  release();
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  if (n_states) *n_states = count;
  if (n_solves) *n_solves = I.solves;
  if (n_cg_iterations) *n_cg_iterations = I.cg_its;
  if (n_attempts) *n_attempts = I.attempts;
  return FVB_OK;
}

// ---- measurement ---------------------------------------------------------------------------------------
int fvb_time_spmv(fvb_handle h, int warmup, int reps, double *ms_avg) {
  FVB_TRY(check_handle(h, true));
  FVB_TRY(ensure_workspace(h));
  if (reps < 1) return set_error(FVB_ERR_BAD_INPUT, "reps must be >= 1");
  const int64_t n = h->nf_local;
  cudaStream_t st = h->stream;
  // x = diag(A) (any resident non-trivial vector will do; contents do not affect traffic)
  FVB_CUDA(cudaMemcpyAsync(h->u, h->diag, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < warmup; ++i) FVB_TRY(launch_spmv(h, h->u, h->c, 0.0, false));
  FVB_CUDA(cudaEventRecord(h->ev[3], st));
  for (int i = 0; i < reps; ++i) FVB_TRY(launch_spmv(h, h->u, h->c, 0.0, false));
  FVB_CUDA(cudaEventRecord(h->ev[4], st));
  FVB_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  FVB_CUDA(cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]));
  *ms_avg = (double)ms / reps;
  return FVB_OK;
}

int fvb_set_spmv_format(fvb_handle h, int fmt) {
  FVB_TRY(check_handle(h, false));
  if (fmt < 0 || fmt > 3)
    return set_error(FVB_ERR_BAD_INPUT, "format must be 0 (auto), 1 (CSR), 2 (diagonal, per-thread loads) or 3 (diagonal, TMA)");
  h->fmt_request = fmt;
  if (fmt != 1 && h->assembled && !h->dia_on) FVB_TRY(build_dia(h, true));
  return FVB_OK;
}

int fvb_get_spmv_format(fvb_handle h, int *active, int *n_offsets) {
  FVB_TRY(check_handle(h, true));
  const bool dia = h->dia_on && h->fmt_request != 1;
  int kind = 1;
  if (dia) {
    const DiaTmaLayout L = dia_tma_layout(h->dia_K, h->dia_off, false);
    kind = use_dia_tma(h, h->u, L, false) ? 3 : 2;  // (h->u may not exist yet: alignment is then taken for granted)
  }
  if (active) *active = kind;
  if (n_offsets) *n_offsets = dia ? h->dia_K : 0;
  return FVB_OK;
}

int fvb_set_assembly(fvb_handle h, int mode) {
  FVB_TRY(check_handle(h, false));
  if (mode != 0 && mode != 1) return set_error(FVB_ERR_BAD_INPUT, "assembly mode must be 0 (auto) or 1 (general path only)");
  h->box_request = mode;
  return FVB_OK;
}

int fvb_get_assembly(fvb_handle h, int *active) {
  FVB_TRY(check_handle(h, true));
  if (active) *active = h->box ? (h->box_implicit ? 2 : 1) : 0;
  return FVB_OK;
}

int fvb_set_pcg_scaling(fvb_handle h, int mode) {
  FVB_TRY(check_handle(h, false));
  if (mode != 0 && mode != 1) return set_error(FVB_ERR_BAD_INPUT, "scaling mode must be 0 (auto) or 1 (off)");
  h->scale_request = mode;
  return FVB_OK;
}

int fvb_get_pcg_scaling(fvb_handle h, int *last_solve_scaled) {
  FVB_TRY(check_handle(h, true));
  if (last_solve_scaled) *last_solve_scaled = h->last_solve_scaled ? 1 : 0;
  return FVB_OK;
}

int fvb_device_alloc(fvb_handle h, int64_t bytes, void **dev_ptr) {
  FVB_TRY(check_handle(h, false));
  if (bytes < 0 || !dev_ptr) return set_error(FVB_ERR_BAD_INPUT, "bad allocation request");
  unsigned char *p = nullptr;
  FVB_TRY(dalloc(h, &p, bytes));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  *dev_ptr = p;
  return FVB_OK;
}

int fvb_device_free(fvb_handle h, void *dev_ptr) {
  FVB_TRY(check_handle(h, false));
  unsigned char *p = static_cast<unsigned char *>(dev_ptr);
  dfree(h, p);
  return FVB_OK;
}

int fvb_device_copy(fvb_handle h, void *dst, const void *src, int64_t bytes) {
  FVB_TRY(check_handle(h, false));
  FVB_CUDA(memcpy_sync(h->stream, dst, src, (size_t)bytes, cudaMemcpyDefault));
  return FVB_OK;
}

int fvb_regulargrid(fvb_handle h, const double mins[3], const double maxs[3], const int64_t ns[3], int64_t plane_lo,
                    int64_t plane_hi, int64_t *n_faces, int64_t *neighbors, double *aol, double *volumes) {
  FVB_TRY(check_handle(h, false));
  if (!mins || !maxs || !ns) return set_error(FVB_ERR_BAD_INPUT, "null grid description");
  if (ns[0] < 2 || ns[1] < 2 || ns[2] < 2) return set_error(FVB_ERR_BAD_INPUT, "regulargrid needs at least 2 points per axis");
  if (plane_lo < 1 || plane_hi > ns[0] || plane_hi < plane_lo) return set_error(FVB_ERR_BAD_INPUT, "bad plane range");
  GridDesc G;
  G.n1 = ns[0]; G.n2 = ns[1]; G.n3 = ns[2];
  // range(lo; stop=hi, length=n): dx = xs[2] - xs[1]   (src/grid.jl:62-67)
  auto axis = [](double lo, double hi, int64_t n) { double step = (hi - lo) / (double)(n - 1); return (lo + 1 * step) - lo; };
  G.dx = axis(mins[0], maxs[0], ns[0]); G.dy = axis(mins[1], maxs[1], ns[1]); G.dz = axis(mins[2], maxs[2], ns[2]);
  G.p_lo = plane_lo; G.p_hi = plane_hi; G.e_lo = plane_lo > 1 ? plane_lo - 1 : plane_lo;
  const int64_t plane = G.n2 * G.n3, pfull = plane + (G.n2 - 1) * G.n3 + G.n2 * (G.n3 - 1);
  int64_t F = (G.e_lo < G.p_lo ? plane : 0);
  for (int64_t i1 = plane_lo; i1 <= plane_hi; ++i1) F += pfull - (i1 < G.n1 ? 0 : plane);
  if (n_faces) *n_faces = F;
  if (!neighbors) return FVB_OK;
  if (!aol) return set_error(FVB_ERR_BAD_INPUT, "null areasoverlengths buffer");
  const int64_t nodes = (G.p_hi - G.e_lo + 1) * plane;
  k_regulargrid<<<std::max(1, std::min(cdiv(nodes, kBlock), h->num_sms * 16)), kBlock, 0, h->stream>>>(
      G, reinterpret_cast<longlong2 *>(neighbors), aol, volumes);
  h->tm.kernel_launches++;
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  FVB_CUDA(cudaGetLastError());
  return FVB_OK;
}

int fvb_nodehycos2neighborhycos(fvb_handle h, int64_t n_faces, const int64_t *neighbors_dev, const double *nodehycos,
                                int64_t node_lo, int64_t n_have, int logmean, double *out_dev) {
  FVB_TRY(check_handle(h, false));
  if (n_faces < 0 || n_have < 0 || (n_faces && (!neighbors_dev || !nodehycos || !out_dev)))
    return set_error(FVB_ERR_BAD_INPUT, "bad arguments");
  cudaStream_t st = h->stream;
  double *d_k = nullptr;
  int *d_err = nullptr;
  FVB_TRY(dalloc(h, &d_k, n_have));
  FVB_TRY(dalloc(h, &d_err, 1));
  int init = INT_MAX;
  cudaMemcpyAsync(d_err, &init, sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_k, nodehycos, sizeof(double) * (size_t)n_have, cudaMemcpyDefault, st);
  if (n_faces) {
    k_node2face<<<std::max(1, std::min(cdiv(n_faces, kBlock), h->num_sms * 16)), kBlock, 0, st>>>(
        n_faces, reinterpret_cast<const longlong2 *>(neighbors_dev), d_k, node_lo, n_have, logmean, out_dev, d_err);
    h->tm.kernel_launches++;
  }
  int bad = INT_MAX;
  cudaError_t e = memcpy_sync(st, &bad, d_err, sizeof(int), cudaMemcpyDeviceToHost);
  dfree(h, d_k); dfree(h, d_err);
  if (e != cudaSuccess) return set_error(FVB_ERR_CUDA, cudaGetErrorString(e));
  if (bad != INT_MAX) return set_error(FVB_ERR_BAD_INPUT, "face " + std::to_string(bad + 1) + " references a node outside the supplied nodehycos range");
  return FVB_OK;
}

int fvb_set_preconditioner(fvb_handle h, int kind, int nu, double omega, double oc) {
  FVB_TRY(check_handle(h, false));
  if (kind != 0 && kind != 1) return set_error(FVB_ERR_BAD_INPUT, "preconditioner kind must be 0 (Jacobi) or 1 (multigrid)");
  if (kind == 1 && h->assembled && h->box_implicit && h->mg && h->mg->ready) { /* hierarchy already there */ }
  if (nu < 0 || nu > 8 || omega < 0 || omega >= 2 || oc < 0) return set_error(FVB_ERR_BAD_INPUT, "bad multigrid parameters");
  h->precond_request = kind;
  if (nu > 0) h->mg_nu = nu;
  if (omega > 0) h->mg_omega = omega;
  if (oc > 0) h->mg_oc = oc;
  if (kind == 1 && h->assembled) {
    if (!(h->mg && h->mg->ready) && !(h->amg && h->amg->ready)) FVB_TRY(precond_setup(h));
    FVB_CUDA(cudaStreamSynchronize(h->stream));
    if (!use_geometric_mg(h) && !use_amg(h)) {
      h->precond_request = 0;
      return set_error(FVB_ERR_BAD_INPUT, "no multigrid hierarchy for this matrix: the geometric one needs a box-structured 7-point "
                                          "matrix (diagonal format with offsets 1, nz, ny*nz), the algebraic one a single-rank CSR");
    }
  }
  return FVB_OK;
}

int fvb_get_preconditioner(fvb_handle h, int *active_kind, int *n_levels) {
  FVB_TRY(check_handle(h, true));
  const bool mg = use_geometric_mg(h), amg = use_amg(h);
  if (active_kind) *active_kind = mg ? 1 : (amg ? 2 : 0);
  if (n_levels) *n_levels = mg ? h->mg->nlev : (amg ? h->amg->nlev : 0);
  return FVB_OK;
}

int fvb_set_profiling(fvb_handle h, int stride) {
  FVB_TRY(check_handle(h, false));
  if (stride < 0) return set_error(FVB_ERR_BAD_INPUT, "stride must be >= 0");
  if (stride > 0 && !h->prof_ev[0])
    for (auto &ev : h->prof_ev) FVB_CUDA(cudaEventCreate(&ev));
  h->prof_stride = stride;
  return FVB_OK;
}

int fvb_get_timings(fvb_handle h, fvb_timings *out) {
  if (!h || !out) return set_error(FVB_ERR_BAD_INPUT, "null argument");
  *out = h->tm;
  return FVB_OK;
}

int fvb_sync(fvb_handle h) {
  FVB_TRY(check_handle(h, false));
  FVB_CUDA(cudaStreamSynchronize(h->stream));
  return FVB_OK;
}

}  // extern "C"

#include "multi_impl.h"
