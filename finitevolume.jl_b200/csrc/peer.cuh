// peer.cuh -- the per-iteration exchanges of the slab-partitioned PCG done by the GPUs themselves
// over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC), instead of
// three NCCL calls per iteration whose latency (~25 us each at these message sizes) capped
// 8-GPU strong scaling at 74 %:
//   * halo planes: k_halo_push stores the boundary rows of u straight into the neighbours' halo
//     slots (remote st.global over NVLink), fences, then raises a per-source sequence flag in the
//     neighbour's mailbox; the neighbour's k_halo_wait (one warp) spins on its LOCAL flag before
//     the SpMV is released.
//   * CG scalars (u.Au; (D^-1 r).r and r.r): k_allreduce_fin -- one warp -- writes this rank's
//     partial sums into every rank's mailbox, spins until all ranks' slots carry the current
//     sequence number, adds them in rank order (every rank gets bit-identical sums, so all stop on
//     the same iteration) and advances the recurrence in the same launch.
// Sequence numbers are kept on the host and are identical on all ranks because every rank issues
// the same collective calls in the same order.  Slots alternate by parity: a slot is rewritten two
// reductions later, which cannot start before every rank consumed the previous use (the next
// reduction needs every rank's contribution, made after that rank finished reading).
// Every spin is bounded (kPeerTimeoutNs); on timeout the solve is flagged and stops.
// The reducing kernels of the Jacobi-PCG perform the same all-reduce themselves in their last CTA
// (peer_base.cuh: peer_allreduce_thread, finalize_mode 2); k_allreduce_fin remains for the multigrid
// PCG, the global norms of the transient stepper and FVB_FUSED_ALLREDUCE=0.
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include "pcg.cuh"
#include "peer_base.cuh"

namespace fvb {

// ---- halo ------------------------------------------------------------------------------------------------
struct HaloPlanDev {
  int npeers;
  int peer[kMaxRanks];
  long long send_begin[kMaxRanks + 1];  // prefix offsets into send_rows
  long long dst_off[kMaxRanks];         // index in the peer's u where my first value goes
};

__global__ void __launch_bounds__(kBlock)
k_halo_push(PeerTable T, HaloPlanDev P, const int *__restrict__ send_rows, const double *__restrict__ u,
            unsigned long long seq, unsigned int *ticket) {
  const long long total = P.send_begin[P.npeers];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = 0;
    while (p + 1 < P.npeers && i >= P.send_begin[p + 1]) ++p;
    T.u[P.peer[p]][P.dst_off[p] + (i - P.send_begin[p])] = u[send_rows[i]];
  }
  __threadfence_system();  // my remote stores are visible system-wide before the ticket
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < P.npeers) {
    __threadfence_system();
    st_release_sys(&T.mail[P.peer[threadIdx.x]]->hseq[T.rank], seq);
  }
}

// in_cg: launched inside the CG loop, where iterations enqueued past convergence are no-ops (the producer raises the
// flag anyway, but there is nothing to wait for).
__global__ void k_halo_wait(PeerMail *mail, HaloPlanDev P, unsigned long long seq, PcgScal *scal, int in_cg) {
  if (in_cg && scal->done) return;
  bool ok = true;
  if ((int)threadIdx.x < P.npeers) ok = wait_flag(&mail->hseq[P.peer[threadIdx.x]], seq);
  if (!ok) { mail->error = 1; scal->done = 1; scal->converged = 0; }
}

// ---- all-reduce of <= 4 doubles + recurrence update -------------------------------------------------
enum { FIN_NONE = 0, FIN_INIT = 1, FIN_UC = 2, FIN_ITER = 3, FIN_RZ = 4, FIN_R = 5 };

__global__ void k_allreduce_fin(PeerTable T, unsigned long long seq, double *red, int count, int mode,
                                PcgScal *scal, double *hist) {
  const int lane = threadIdx.x;
  const int par = (int)(seq & 1ull);
  PeerMail *mine = T.mail[T.rank];
  if (lane < T.nranks) {
    PeerMail *dst = T.mail[lane];
    for (int i = 0; i < count; ++i) dst->vals[par][T.rank][i] = red[i];
    __threadfence_system();
    st_release_sys(&dst->vseq[par][T.rank], seq);
  }
  bool ok = true;
  if (lane < T.nranks) ok = wait_flag(&mine->vseq[par][lane], seq);
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    if (!ok) {
      mine->error = 1; scal->done = 1; scal->converged = 0;
      return;
    }
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < T.nranks; ++r)
      for (int i = 0; i < count; ++i) s[i] += *(volatile double *)&mine->vals[par][r][i];
    for (int i = 0; i < count; ++i) red[i] = s[i];
    if (mode == FIN_INIT) pcg_finish_init(scal, s[0], s[1]);
    else if (mode == FIN_UC) { if (!scal->done) scal->uc = s[0]; }
    else if (mode == FIN_ITER) { if (!scal->done) pcg_finish_iter(scal, s[0], s[1], hist); }
    else if (mode == FIN_RZ) mgpcg_finish_rz(scal, s[0]);
    else if (mode == FIN_R) mgpcg_finish_r(scal, s[0], hist);
  }
}

}  // namespace fvb
