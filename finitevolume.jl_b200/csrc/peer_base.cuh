// peer_base.cuh -- mailbox types and the in-kernel all-reduce of the CG scalars over NVLink peer
// memory (see peer.cuh for the protocol).  Kept apart from peer.cuh so that the reducing kernels
// themselves (spmv.cuh, dia.cuh, dia_tma.cuh, pcg.cuh) can finish their reduction ACROSS the
// GPUs in the same launch: the CTA that folds the per-CTA partials writes this rank's sum into
// every rank's mailbox, waits for the other ranks' sums, adds them in rank order and advances the
// recurrence -- compute and collective in one kernel, no separate all-reduce launch on the
// critical path of the iteration.
#pragma once
#include "common.cuh"

namespace fvb {

constexpr int kMaxRanks = 8;
constexpr unsigned long long kPeerTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;

struct PeerMail {
  double vals[2][kMaxRanks][4];            // [parity][source rank][value]
  unsigned long long vseq[2][kMaxRanks];   // sequence number of the values above
  unsigned long long hseq[kMaxRanks];      // last halo push received from each source rank
  int error;                               // set locally when a wait timed out
};

struct PeerTable {
  PeerMail *mail[kMaxRanks];  // mail[r] = rank r's mailbox as mapped into THIS process (mail[rank] local)
  double *u[kMaxRanks];       // u[r]    = rank r's search-direction vector (owned rows + halo slots)
  int nranks, rank;
};

// What a reducing kernel needs to finish its sums across the ranks (finalize_mode == 2).
struct PeerRed {
  const PeerTable *tab;       // device copy of the table (null: no fused all-reduce)
  unsigned long long seq;     // sequence number of this reduction (host-side counter, same on all ranks)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *flag >= want; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long want) {
  if (ld_acquire_sys(flag) >= want) return true;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) < want) {
    if (globaltimer_ns() - t0 > kPeerTimeoutNs) return false;
    __nanosleep(64);
  }
  return true;
}

// All-reduce (sum) of v[0..count) over the ranks, executed by ONE thread (the thread that just
// folded this rank's per-CTA partials).  Every rank adds the contributions in rank order, so all
// ranks obtain bit-identical sums.  Returns false on timeout (the solve is flagged and stopped).
__device__ __forceinline__ bool peer_allreduce_thread(const PeerRed &pr, double *v, int count, PcgScal *scal) {
  const PeerTable &T = *pr.tab;
  const int par = (int)(pr.seq & 1ull);
  PeerMail *mine = T.mail[T.rank];
  for (int r = 0; r < T.nranks; ++r)
    for (int i = 0; i < count; ++i) T.mail[r]->vals[par][T.rank][i] = v[i];
  __threadfence_system();
  for (int r = 0; r < T.nranks; ++r) st_release_sys(&T.mail[r]->vseq[par][T.rank], pr.seq);
  bool ok = true;
  for (int r = 0; r < T.nranks; ++r) ok = wait_flag(&mine->vseq[par][r], pr.seq) && ok;
  if (!ok) {
    mine->error = 1; scal->done = 1; scal->converged = 0;
    return false;
  }
  for (int i = 0; i < count; ++i) {
    double s = 0.0;
    for (int r = 0; r < T.nranks; ++r) s += *(volatile double *)&mine->vals[par][r][i];
    v[i] = s;
  }
  return true;
}

// The same all-reduce executed by the first WARP of the CTA that folded this rank's partials (all 32 lanes must
// call it, converged; v is valid in lane 0 on entry and in every lane on return): lane q writes this rank's sums
// into rank q's mailbox, fences and raises rank q's flag, then waits for rank q's contribution -- the eight
// remote store/fence/flag sequences and the eight waits overlap instead of running back to back in one thread.
__device__ __forceinline__ bool peer_allreduce_warp(const PeerRed &pr, double *v, int count, PcgScal *scal) {
  const PeerTable &T = *pr.tab;
  const int lane = threadIdx.x & 31;
  const int par = (int)(pr.seq & 1ull);
  for (int i = 0; i < count; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], 0);
  PeerMail *mine = T.mail[T.rank];
  bool ok = true;
  if (lane < T.nranks) {
    PeerMail *dst = T.mail[lane];
    for (int i = 0; i < count; ++i) dst->vals[par][T.rank][i] = v[i];
    __threadfence_system();
    st_release_sys(&dst->vseq[par][T.rank], pr.seq);
    ok = wait_flag(&mine->vseq[par][lane], pr.seq);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (!ok) {
    if (lane == 0) { mine->error = 1; scal->done = 1; scal->converged = 0; }
    return false;
  }
  for (int i = 0; i < count; ++i) {
    double s = 0.0;
    for (int r = 0; r < T.nranks; ++r) s += *(volatile double *)&mine->vals[par][r][i];
    v[i] = s;
  }
  return true;
}

// Halo push fused into the kernel that produces the search direction (pcg.cuh: k_update_u): slab partitions send
// contiguous row ranges (the first / last plane), so the thread that writes u[i] also stores it into the
// neighbour's halo slot; the CTA drawing the last ticket raises the neighbours' flags.  npeers = 0: nothing fused.
struct FusedPush {
  int npeers;
  double *dst[2];                // neighbour's vector, already offset to this rank's first halo slot there
  long long begin[2], count[2];  // my rows [begin, begin + count)
  unsigned long long *flag[2];   // &neighbour_mailbox->hseq[my rank]
  unsigned long long seq;
};
// Halo wait fused into the SpMV (dia_tma.cuh): only the tiles at the slab ends read halo values, they are dealt
// last, and a CTA spins on the (local) flags right before its first such tile.  n = 0: nothing to wait for.
struct FusedWait {
  int n;
  const unsigned long long *flag[2];  // &my_mailbox->hseq[neighbour]
  unsigned long long seq;
  int *error;                         // &my_mailbox->error
};

}  // namespace fvb
