// multi_impl.h -- single-process multi-GPU front end (included at the end of fvb200.cu).
//
// The reference is one single-threaded Julia process (SURVEY 8b "Threading"): a drop-in must let ONE host thread
// make ONE call and have the problem solved on all the GPUs of the box, without torchrun or MPI.  An fvb_multi
// owns one per-device handle (the very same code path as the one-process-per-GPU mode: slab of the node range,
// NVLink peer-memory halo + in-kernel all-reduce) and drives the devices from short-lived worker threads, one
// per device, so the caller's thread blocks in a single call exactly as a `ccall` would.  Differences from the
// multi-process mode: peers are mapped by cudaDeviceEnablePeerAccess (no CUDA IPC: one address space), and the
// halo plan is computed inside the library (the port of distributed.halo_plan_from_ranges) instead of being
// exchanged over torch.distributed.
//
// Partition: contiguous node ranges.  When the face list looks like regulargrid's (first faces and
// the face count match the closed form) the ranges are whole x-planes balanced by free planes, every device gets
// exactly the slab list regulargrid(planes=...) would emit -- the +x faces of the plane below gathered on the
// host (n2*n3 faces), the rest one contiguous block copied straight from the caller's arrays -- and each
// device verifies its list on the GPU (box.cuh).  Anything else: equal node counts, faces filtered on the host.
//
// Invariant the worker threads rely on: between the first and the last cross-device kernel of a collective call no
// thread may call into the driver in a way that synchronises devices (cudaMalloc / cudaFree map into every peer's
// address space once peer access is on, and a device that is spinning on a neighbour's flag never becomes idle).
// Every allocation of a solve therefore happens in multi_connect (vectors, mailboxes) or comes out of the handle's
// arena, whose chunks are in place after assembly (the temporaries freed there leave gigabytes of free blocks).
#pragma once
#include <thread>

struct fvb_multi_s {
  int ndev = 0;
  std::vector<int> devs;
  std::vector<fvb_handle> h;
  int64_t n_nodes = 0;
  std::vector<int64_t> lo, hi;  // 1-based inclusive node ranges
  bool assembled = false;
  bool regular = false;
};

namespace {

template <class Fn>
int multi_run_all(fvb_multi m, Fn fn) {
  std::vector<int> st((size_t)m->ndev, FVB_OK);
  std::vector<std::string> err((size_t)m->ndev);
  if (m->ndev == 1) {
    st[0] = fn(0);
    if (st[0] != FVB_OK) err[0] = g_last_error;
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < m->ndev; ++r)
      th.emplace_back([&, r]() {
        st[(size_t)r] = fn(r);
        if (st[(size_t)r] != FVB_OK) err[(size_t)r] = g_last_error;  // thread-local in the worker
      });
    for (auto &t : th) t.join();
  }
  for (int r = 0; r < m->ndev; ++r)
    if (st[(size_t)r] != FVB_OK) return set_error(st[(size_t)r], "device " + std::to_string(m->devs[(size_t)r]) + ": " + err[(size_t)r]);
  return FVB_OK;
}

// Per-rank halo plans from the assembled handles (host_util.h: plan_halo_exchange, the port of
// distributed.halo_plan_from_ranges + send_destinations).
using MultiPlan = RankHaloPlan;
int multi_halo_plans(fvb_multi m, std::vector<MultiPlan> &plans) {
  const int P = m->ndev;
  std::vector<int64_t> start((size_t)P), nf((size_t)P);
  std::vector<std::vector<int64_t>> halo((size_t)P);
  for (int r = 0; r < P; ++r) {
    start[(size_t)r] = m->h[(size_t)r]->row_start;
    nf[(size_t)r] = m->h[(size_t)r]->nf_local;
    halo[(size_t)r] = m->h[(size_t)r]->halo_host;
  }
  const char *why = plan_halo_exchange(start, nf, halo, plans);
  if (why[0]) return set_error(FVB_ERR_STATE, why);
  return FVB_OK;
}

// After every device's problem is assembled: halo plans, work vectors, peer mappings (same address space).
int multi_connect(fvb_multi m) {
  if (m->ndev == 1) return FVB_OK;
  std::vector<MultiPlan> plans;
  FVB_TRY(multi_halo_plans(m, plans));
  FVB_TRY(multi_run_all(m, [&](int r) {
    fvb_handle h = m->h[(size_t)r];
    const MultiPlan &pl = plans[(size_t)r];
    FVB_TRY(fvb_set_halo_plan(h, (int)pl.peers.size(), pl.peers.data(), pl.send_counts.data(), pl.send_rows.data(),
                              pl.recv_counts.data()));
    FVB_TRY(ensure_workspace(h));
    if (!h->peer) h->peer = new PeerState();
    h->peer->active = false;
    if (!h->peer->mail) {
      FVB_CUDA(cudaMalloc((void **)&h->peer->mail, sizeof(PeerMail)));
      FVB_CUDA(cudaMemsetAsync(h->peer->mail, 0, sizeof(PeerMail), h->stream));
    }
    FVB_CUDA(cudaStreamSynchronize(h->stream));
    return (int)FVB_OK;
  }));
  // every device's u and mailbox exist now: hand the raw pointers around (UVA + peer access)
  return multi_run_all(m, [&](int r) {
    fvb_handle h = m->h[(size_t)r];
    FVB_TRY(check_handle(h, true));
    PeerState &P = *h->peer;
    for (int q = 0; q < kMaxRanks; ++q) {  // mappings of an earlier multi-process life of this handle
      if (P.opened_u[q]) { cudaIpcCloseMemHandle(P.opened_u[q]); P.opened_u[q] = nullptr; }
      if (P.opened_mail[q]) { cudaIpcCloseMemHandle(P.opened_mail[q]); P.opened_mail[q] = nullptr; }
    }
    P.last_blobs.clear();
    P.tab = PeerTable{};
    P.tab.nranks = m->ndev;
    P.tab.rank = r;
    for (int q = 0; q < m->ndev; ++q) {
      P.tab.u[q] = m->h[(size_t)q]->u;
      P.tab.mail[q] = m->h[(size_t)q]->peer->mail;
    }
    return peer_finish_plan(h, plans[(size_t)r].send_dst.data());
  });
}

inline int64_t multi_faces_before(int64_t n1, int64_t n2, int64_t n3, int64_t i1, int64_t i2, int64_t i3) {
  return regulargrid_faces_before(n1, n2, n3, i1, i2, i3);
}
inline void multi_slab_planes(int64_t n1, int P, bool dirichlet_ends, std::vector<int64_t> &plo, std::vector<int64_t> &phi) {
  slab_planes(n1, P, dirichlet_ends, plo, phi);
}

struct MultiInputs {
  int64_t n_nodes, n_faces;
  const int64_t *nb; const double *aol; const double *cond; int64_t n_cond; const int64_t *meta; int logk;
  const double *sources; int64_t nd; const int64_t *dnodes; const double *dheads;
};

// Regular attempt: returns FVB_OK with *done=false when the list does not look like regulargrid's or a device
// did not take the closed-form path (the caller then filters faces on the host).
int multi_assemble_regular(fvb_multi m, const MultiInputs &in, const std::vector<int64_t> &dn_sorted, bool *done) {
  *done = false;
  const int P = m->ndev;
  const int64_t N = in.n_nodes, F = in.n_faces;
  if (F < 3 || N < 8) return FVB_OK;
  const int64_t plane = in.nb[1] - in.nb[0];
  if (in.nb[0] != 1 || plane < 4 || N % plane) return FVB_OK;
  const int64_t n1 = N / plane;
  if (n1 < 2 || n1 < P) return FVB_OK;
  if (in.nb[2] != 1 || in.nb[4] != 1 || in.nb[5] != 2) return FVB_OK;
  const int64_t n3 = in.nb[3] - 1;
  if (n3 < 2 || n3 >= plane || plane % n3 || plane / n3 < 2) return FVB_OK;
  const int64_t n2 = plane / n3;
  if (F != 3 * N - n1 * n2 - n1 * n3 - n2 * n3) return FVB_OK;
  auto dcount = [&](int64_t first) {
    return std::lower_bound(dn_sorted.begin(), dn_sorted.end(), first + plane) - std::lower_bound(dn_sorted.begin(), dn_sorted.end(), first);
  };
  const bool ends = dcount(0) == plane && dcount(N - plane) == plane;
  std::vector<int64_t> plo, phi;
  multi_slab_planes(n1, P, ends, plo, phi);
  for (int r = 0; r < P; ++r) { m->lo[(size_t)r] = (plo[(size_t)r] - 1) * plane + 1; m->hi[(size_t)r] = phi[(size_t)r] * plane; }
  FVB_TRY(multi_run_all(m, [&](int r) {
    fvb_handle h = m->h[(size_t)r];
    FVB_TRY(check_handle(h, false));
    const int64_t p_lo = plo[(size_t)r], p_hi = phi[(size_t)r];
    const bool halo = p_lo > 1;
    const int64_t f0 = multi_faces_before(n1, n2, n3, p_lo, 1, 1);
    const int64_t f1 = p_hi < n1 ? multi_faces_before(n1, n2, n3, p_hi + 1, 1, 1) : F;
    const int64_t nh = halo ? plane : 0, Fr = nh + (f1 - f0);
    // +x faces of the plane below, in node order
    std::vector<int64_t> hnb((size_t)(2 * nh)), hmeta;
    std::vector<double> haol((size_t)nh), hcond;
    if (in.meta) hmeta.resize((size_t)nh); else hcond.resize((size_t)nh);
    for (int64_t t = 0; t < nh; ++t) {
      const int64_t j = multi_faces_before(n1, n2, n3, p_lo - 1, t / n3 + 1, t % n3 + 1);
      hnb[(size_t)(2 * t)] = in.nb[2 * j]; hnb[(size_t)(2 * t + 1)] = in.nb[2 * j + 1];
      haol[(size_t)t] = in.aol[j];
      if (in.meta) hmeta[(size_t)t] = in.meta[j]; else hcond[(size_t)t] = in.cond[j];
    }
    int64_t *d_nb = nullptr, *d_meta = nullptr;
    double *d_aol = nullptr, *d_cond = nullptr;
    cudaStream_t st = h->stream;
    auto drop = [&]() { dfree(h, d_nb); dfree(h, d_meta); dfree(h, d_aol); dfree(h, d_cond); };
    int s = dalloc(h, &d_nb, 2 * Fr);
    if (s == FVB_OK) s = dalloc(h, &d_aol, Fr);
    if (s == FVB_OK) s = in.meta ? dalloc(h, &d_meta, Fr) : dalloc(h, &d_cond, Fr);
    cudaError_t e = cudaSuccess;
    auto up = [&](void *dst, const void *head, const void *block, size_t elem) {
      if (e == cudaSuccess && nh) e = cudaMemcpyAsync(dst, head, elem * (size_t)nh, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess && f1 > f0)
        e = cudaMemcpyAsync((char *)dst + elem * (size_t)nh, (const char *)block + elem * (size_t)f0, elem * (size_t)(f1 - f0),
                            cudaMemcpyHostToDevice, st);
    };
    if (s == FVB_OK) {
      up(d_nb, hnb.data(), in.nb, 2 * sizeof(int64_t));
      up(d_aol, haol.data(), in.aol, sizeof(double));
      if (in.meta) up(d_meta, hmeta.data(), in.meta, sizeof(int64_t)); else up(d_cond, hcond.data(), in.cond, sizeof(double));
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the small host vectors die with this scope
      if (e != cudaSuccess) s = set_error(FVB_ERR_CUDA, std::string("slab upload: ") + cudaGetErrorString(e));
    }
    if (s == FVB_OK)
      s = fvb_assemble(h, N, m->lo[(size_t)r], m->hi[(size_t)r], Fr, d_nb, d_aol, in.meta ? in.cond : d_cond,
                       in.meta ? in.n_cond : Fr, d_meta, in.logk, in.sources + (m->lo[(size_t)r] - 1), in.nd, in.dnodes, in.dheads);
    drop();
    return s;
  }));
  for (int r = 0; r < P; ++r)
    if (!m->h[(size_t)r]->box) return FVB_OK;  // some device fell back to the general path: redo with filtered lists
  *done = true;
  return FVB_OK;
}

int multi_assemble_filtered(fvb_multi m, const MultiInputs &in) {
  const int P = m->ndev;
  const int64_t N = in.n_nodes, F = in.n_faces;
  for (int r = 0; r < P; ++r) { m->lo[(size_t)r] = N * r / P + 1; m->hi[(size_t)r] = N * (r + 1) / P; }
  return multi_run_all(m, [&](int r) {
    fvb_handle h = m->h[(size_t)r];
    const int64_t lo = m->lo[(size_t)r], hi = m->hi[(size_t)r];
    std::vector<int64_t> nb, meta;
    std::vector<double> aol, cond;
    for (int64_t i = 0; i < F; ++i) {
      const int64_t a = in.nb[2 * i], b = in.nb[2 * i + 1];
      if ((a >= lo && a <= hi) || (b >= lo && b <= hi)) {
        nb.push_back(a); nb.push_back(b);
        aol.push_back(in.aol[i]);
        if (in.meta) meta.push_back(in.meta[i]); else cond.push_back(in.cond[i]);
      }
    }
    const int64_t Fr = (int64_t)aol.size();
    static const int64_t zero_nb[2] = {0, 0};
    static const double zero_d[1] = {0.0};
    return fvb_assemble(h, N, lo, hi, Fr, Fr ? nb.data() : zero_nb, Fr ? aol.data() : zero_d,
                        in.meta ? in.cond : (Fr ? cond.data() : zero_d), in.meta ? in.n_cond : Fr,
                        in.meta ? (Fr ? meta.data() : zero_nb) : nullptr, in.logk, in.sources + (lo - 1), in.nd, in.dnodes,
                        in.dheads);
  });
}

int multi_check(fvb_multi m, bool need_assembled) {
  if (!m) return set_error(FVB_ERR_BAD_INPUT, "null multi-device handle");
  if (need_assembled && !m->assembled) return set_error(FVB_ERR_STATE, "fvb_multi_assemble has not succeeded on this handle");
  return FVB_OK;
}

}  // namespace

extern "C" {

int fvb_multi_create(int ndev, const int *device_ids, fvb_multi *out) {
  if (!out) return set_error(FVB_ERR_BAD_INPUT, "null out pointer");
  *out = nullptr;
  int have = 0;
  cudaError_t e = cudaGetDeviceCount(&have);
  if (e != cudaSuccess || have == 0) {
    cudaGetLastError();
    return set_error(FVB_ERR_CUDA, "no CUDA device available (libfvb200 has no CPU fallback)");
  }
  if (ndev < 1 || ndev > kMaxRanks) return set_error(FVB_ERR_BAD_INPUT, "ndev must be 1..8");
  fvb_multi m = new fvb_multi_s();
  m->ndev = ndev;
  for (int r = 0; r < ndev; ++r) m->devs.push_back(device_ids ? device_ids[r] : r);
  for (int r = 0; r < ndev; ++r)
    for (int q = 0; q < r; ++q)
      if (m->devs[(size_t)r] == m->devs[(size_t)q]) {
        delete m;
        return set_error(FVB_ERR_BAD_INPUT, "a device is listed twice (kernels of different ranks wait on one another and must not share a GPU)");
      }
  m->h.assign((size_t)ndev, nullptr);
  m->lo.assign((size_t)ndev, 1);
  m->hi.assign((size_t)ndev, 0);
  auto fail = [&](int s) { std::string keep = g_last_error; fvb_multi_destroy(m); g_last_error = keep; return s; };
  for (int r = 0; r < ndev; ++r) {
    int s = fvb_create(m->devs[(size_t)r], &m->h[(size_t)r]);
    if (s != FVB_OK) return fail(s);
  }
  if (ndev > 1) {
    for (int r = 0; r < ndev; ++r) {
      if (cudaSetDevice(m->devs[(size_t)r]) != cudaSuccess) return fail(set_error(FVB_ERR_CUDA, "cudaSetDevice failed"));
      for (int q = 0; q < ndev; ++q) {
        if (q == r) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->devs[(size_t)r], m->devs[(size_t)q]);
        if (!can) return fail(set_error(FVB_ERR_CUDA, "devices " + std::to_string(m->devs[(size_t)r]) + " and " +
                                                       std::to_string(m->devs[(size_t)q]) + " cannot access each other's memory"));
        cudaError_t pe = cudaDeviceEnablePeerAccess(m->devs[(size_t)q], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
          return fail(set_error(FVB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)));
        cudaGetLastError();
      }
    }
    uint8_t id[FVB_UNIQUE_ID_BYTES];
    int s = fvb_comm_unique_id(id);
    if (s != FVB_OK) return fail(s);
    s = multi_run_all(m, [&](int r) { return fvb_comm_init(m->h[(size_t)r], ndev, r, id); });
    if (s != FVB_OK) return fail(s);
  }
  *out = m;
  return FVB_OK;
}

int fvb_multi_destroy(fvb_multi m) {
  if (!m) return FVB_OK;
  for (auto &h : m->h)
    if (h) { fvb_destroy(h); h = nullptr; }
  delete m;
  return FVB_OK;
}

int fvb_multi_device_handle(fvb_multi m, int i, fvb_handle *out) {
  FVB_TRY(multi_check(m, false));
  if (i < 0 || i >= m->ndev || !out) return set_error(FVB_ERR_BAD_INPUT, "device index out of range");
  *out = m->h[(size_t)i];
  return FVB_OK;
}

int fvb_multi_assemble(fvb_multi m, int64_t n_nodes, int64_t n_faces, const int64_t *neighbors, const double *aol,
                       const double *cond, int64_t n_cond, const int64_t *metaindex, int logk, const double *sources,
                       int64_t nd, const int64_t *dnodes, const double *dheads) {
  FVB_TRY(multi_check(m, false));
  m->assembled = false;
  if (m->ndev == 1) {
    m->lo[0] = 1; m->hi[0] = n_nodes; m->n_nodes = n_nodes;
    FVB_TRY(fvb_assemble(m->h[0], n_nodes, 1, n_nodes, n_faces, neighbors, aol, cond, n_cond, metaindex, logk, sources, nd,
                         dnodes, dheads));
    m->assembled = true;
    return FVB_OK;
  }
  if (n_nodes < 0 || n_faces < 0 || nd < 0 || n_cond < 0) return set_error(FVB_ERR_BAD_INPUT, "negative size");
  if ((n_faces && (!neighbors || !aol || !cond)) || (n_nodes && !sources) || (nd && (!dnodes || !dheads)))
    return set_error(FVB_ERR_BAD_INPUT, "null input array");
  if (!metaindex && n_cond < n_faces) return set_error(FVB_ERR_BAD_INPUT, "conductivities is shorter than neighbors");
  if (n_nodes < m->ndev) return set_error(FVB_ERR_BAD_INPUT, "fewer nodes than devices");
  {  // the multi-device front end slices HOST arrays
    cudaPointerAttributes a;
    if (n_faces && cudaPointerGetAttributes(&a, neighbors) == cudaSuccess && a.type == cudaMemoryTypeDevice)
      return set_error(FVB_ERR_BAD_INPUT, "fvb_multi_assemble takes host arrays (it slices them across the devices)");
    cudaGetLastError();
  }
  m->n_nodes = n_nodes;
  MultiInputs in{n_nodes, n_faces, neighbors, aol, cond, n_cond, metaindex, logk, sources, nd, dnodes, dheads};
  std::vector<int64_t> dn(dnodes, dnodes + nd), dn_sorted;
  std::vector<int> slot;
  for (int64_t v : dn)
    if (v < 1 || v > n_nodes) return set_error(FVB_ERR_BAD_INPUT, "node index out of range 1..N (dirichletnodes)");
  dirichlet_table(dn, dn_sorted, slot);
  bool done = false;
  FVB_TRY(multi_assemble_regular(m, in, dn_sorted, &done));
  m->regular = done;
  if (!done) FVB_TRY(multi_assemble_filtered(m, in));
  FVB_TRY(multi_connect(m));
  m->assembled = true;
  return FVB_OK;
}

int fvb_multi_assemble_regulargrid(fvb_multi m, const double mins[3], const double maxs[3], const int64_t ns[3],
                                   const double *nodehycos, int logmean, int logk, const double *sources, int64_t nd,
                                   const int64_t *dnodes, const double *dheads) {
  FVB_TRY(multi_check(m, false));
  m->assembled = false;
  if (!mins || !maxs || !ns || !nodehycos) return set_error(FVB_ERR_BAD_INPUT, "null grid description");
  if (ns[0] < 2 || ns[1] < 2 || ns[2] < 2) return set_error(FVB_ERR_BAD_INPUT, "regulargrid needs at least 2 points per axis");
  if (ns[0] < m->ndev) return set_error(FVB_ERR_BAD_INPUT, "fewer x-planes than devices");
  if (nd < 0 || (nd && (!dnodes || !dheads))) return set_error(FVB_ERR_BAD_INPUT, "bad Dirichlet arguments");
  const int64_t plane = ns[1] * ns[2], N = ns[0] * plane;
  m->n_nodes = N;
  std::vector<int64_t> dn(dnodes, dnodes + nd), dn_sorted;
  std::vector<int> slot;
  dirichlet_table(dn, dn_sorted, slot);
  auto dcount = [&](int64_t first) {
    return std::lower_bound(dn_sorted.begin(), dn_sorted.end(), first + plane) - std::lower_bound(dn_sorted.begin(), dn_sorted.end(), first);
  };
  std::vector<int64_t> plo, phi;
  multi_slab_planes(ns[0], m->ndev, dcount(0) == plane && dcount(N - plane) == plane, plo, phi);
  for (int r = 0; r < m->ndev; ++r) { m->lo[(size_t)r] = (plo[(size_t)r] - 1) * plane + 1; m->hi[(size_t)r] = phi[(size_t)r] * plane; }
  FVB_TRY(multi_run_all(m, [&](int r) {
    const int64_t k_lo = std::max<int64_t>(1, plo[(size_t)r] - 1);
    return fvb_assemble_regulargrid(m->h[(size_t)r], mins, maxs, ns, plo[(size_t)r], phi[(size_t)r], nodehycos + (k_lo - 1) * plane,
                                    logmean, logk, sources ? sources + (m->lo[(size_t)r] - 1) : nullptr, nd, dnodes, dheads);
  }));
  m->regular = true;
  FVB_TRY(multi_connect(m));
  m->assembled = true;
  return FVB_OK;
}

int fvb_multi_sizes(fvb_multi m, int64_t *nf_global, int64_t *nnz_global, int *ndev, int64_t *node_lo, int64_t *node_hi) {
  FVB_TRY(multi_check(m, true));
  int64_t nf = 0, nnz = 0;
  for (int r = 0; r < m->ndev; ++r) { nf += m->h[(size_t)r]->nf_local; nnz += m->h[(size_t)r]->nnz; }
  if (nf_global) *nf_global = nf;
  if (nnz_global) *nnz_global = nnz;
  if (ndev) *ndev = m->ndev;
  for (int r = 0; r < m->ndev; ++r) {
    if (node_lo) node_lo[r] = m->lo[(size_t)r];
    if (node_hi) node_hi[r] = m->hi[(size_t)r];
  }
  return FVB_OK;
}

int fvb_multi_solve(fvb_multi m, double rtol, int64_t maxiter, double *head_nodes, double *x_free, int64_t *iters,
                    int *converged, double *resnorm_hist, int64_t hist_cap) {
  FVB_TRY(multi_check(m, true));
  std::vector<int64_t> it((size_t)m->ndev, 0);
  std::vector<int> cv((size_t)m->ndev, 0);
  FVB_TRY(multi_run_all(m, [&](int r) {
    fvb_handle h = m->h[(size_t)r];
    return fvb_solve(h, rtol, maxiter, nullptr, head_nodes ? head_nodes + (m->lo[(size_t)r] - 1) : nullptr,
                     x_free ? x_free + h->row_start : nullptr, &it[(size_t)r], &cv[(size_t)r], r == 0 ? resnorm_hist : nullptr,
                     r == 0 ? hist_cap : 0);
  }));
  for (int r = 1; r < m->ndev; ++r)
    if (it[(size_t)r] != it[0] || cv[(size_t)r] != cv[0])
      return set_error(FVB_ERR_STATE, "the devices disagree on the iteration count (reduced sums are bit-identical by construction)");
  if (iters) *iters = it[0];
  if (converged) *converged = cv[0];
  return FVB_OK;
}

int fvb_multi_set_preconditioner(fvb_multi m, int kind, int nu, double omega, double oc) {
  FVB_TRY(multi_check(m, false));
  if (m->assembled && m->ndev > 1)
    return set_error(FVB_ERR_STATE, "choose the preconditioner before fvb_multi_assemble (the hierarchy is agreed at assembly)");
  for (auto h : m->h) FVB_TRY(fvb_set_preconditioner(h, kind, nu, omega, oc));
  return FVB_OK;
}

int fvb_multi_get_b(fvb_multi m, double *b) {
  FVB_TRY(multi_check(m, true));
  for (auto h : m->h) FVB_TRY(fvb_get_b(h, b + h->row_start));
  return FVB_OK;
}

int fvb_multi_get_freenode(fvb_multi m, uint8_t *freenode) {
  FVB_TRY(multi_check(m, true));
  for (int r = 0; r < m->ndev; ++r) FVB_TRY(fvb_get_freenode(m->h[(size_t)r], freenode + (m->lo[(size_t)r] - 1)));
  return FVB_OK;
}

int fvb_multi_get_csr(fvb_multi m, int64_t *ptr, int64_t *idx, double *val) {
  FVB_TRY(multi_check(m, true));
  int64_t off = 0;
  for (int r = 0; r < m->ndev; ++r) {
    fvb_handle h = m->h[(size_t)r];
    std::vector<int64_t> p((size_t)h->nf_local + 1);
    FVB_TRY(fvb_get_csr(h, ptr ? p.data() : nullptr, idx ? idx + off : nullptr, val ? val + off : nullptr));
    if (ptr)
      for (int64_t i = 0; i <= h->nf_local; ++i) ptr[h->row_start + i] = p[(size_t)i] + off;
    off += h->nnz;
  }
  return FVB_OK;
}

}  // extern "C"
