// pcg.cuh -- fused vector kernels of the Jacobi-preconditioned CG.
// Recurrence, naming and stopping rule follow IterativeSolvers.cg 0.8.1, the solver the
// reference calls at src/FiniteVolume.jl:161 and src/transient.jl:52,55 (Pl = Jacobi here,
// per north_star, instead of Ruge-Stueben AMG):
//     c = Pl \ r ; rho_prev = rho ; rho = c.r ; beta = rho/rho_prev ; u = c + beta u
//     c = A u ; alpha = rho / u.c ; x += alpha u ; r -= alpha c ; residual = ||r||
//     stop when residual <= tol * ||r0|| or iteration == maxiter
// finalize_mode of the reducing kernels: 0 = leave the local sums in scal->red (an NCCL all-reduce and a
// k_fin_* launch follow), 1 = single rank: advance the recurrence in the last CTA, 2 = sum over the ranks
// through NVLink peer memory in the last CTA (peer_base.cuh), then advance the recurrence.
// One iteration = three launches: k_update_u, k_spmv<DOT>, k_update_xr.  The update x += alpha u
// of iteration k is deferred into k_update_u of iteration k+1, where u is read anyway (saves one
// pass over u per iteration); k_finish_x applies the last one after the loop.  All scalars stay
// in device memory (PcgScal); every kernel returns at once when scal->done is set, so the
// host can enqueue iterations in batches without synchronising.
//
// SC = true: the same recurrence on the symmetrically scaled system A^ x^ = b^ with
// A^ = D^-1/2 A D^-1/2 (unit diagonal), x^ = D^1/2 x, b^ = D^-1/2 b.  In exact arithmetic Jacobi-PCG
// on A and plain CG on A^ produce the same iterates (u^ = D^1/2 u, r^ = D^-1/2 r, rho = r^.r^ =
// r.D^-1 r, u^.A^u^ = u.Au), but the vector kernels no longer read D^-1 and the SpMV no longer reads
// diag(A): 40 + 40 + 32 = 112 bytes per row and iteration instead of 48 + 48 + 32 = 128.  The
// stopping rule stays the reference's: ||r||_2 = sqrt(sum d_i r^_i^2) is accumulated from diag(A)
// in k_update_xr and written to the history.  x is unscaled by k_finish_x.  Used for cold-started
// steady solves on the diagonal format (fvb200.cu: pcg_run).
#pragma once
#include "common.cuh"
#include "peer_base.cuh"
#include "reduce.cuh"

namespace fvb {

__device__ __forceinline__ void pcg_finish_init(PcgScal *s, double rho0, double rr0) {
  s->rho = rho0;
  s->rho_prev = 1.0;
  s->resid0 = sqrt(rr0);
  s->resid = s->resid0;
  s->reltol = s->resid0 * s->tol;
  s->iter = 0;
  s->converged = s->resid0 <= s->reltol;
  s->done = s->converged || s->maxiter <= 0;
}

__device__ __forceinline__ void pcg_finish_iter(PcgScal *s, double rho_new, double rr, double *hist) {
  s->alpha_prev = s->rho / s->uc;
  s->rho_prev = s->rho;
  s->rho = rho_new;
  const double resid = sqrt(rr);
  s->resid = resid;
  if (s->iter < s->hist_cap) hist[s->iter] = resid;
  s->iter += 1;
  if (resid <= s->reltol) { s->done = 1; s->converged = 1; }
  else if (s->iter >= s->maxiter) s->done = 1;
}

// single-thread kernels used after the NCCL all-reduce in multi-rank runs
__global__ void k_fin_init(PcgScal *s) { pcg_finish_init(s, s->red[0], s->red[1]); }
__global__ void k_fin_uc(PcgScal *s) { if (!s->done) s->uc = s->red[0]; }
__global__ void k_fin_iter(PcgScal *s, double *hist) { if (!s->done) pcg_finish_iter(s, s->red[0], s->red[1], hist); }

__global__ void k_set_scal(PcgScal *s, double tol, long long maxiter, long long hist_cap) {
  s->tol = tol; s->maxiter = maxiter; s->hist_cap = hist_cap;
  s->done = 0; s->converged = 0; s->iter = 0; s->alpha_prev = 0.0;
  s->rho = 0; s->rho_prev = 1; s->uc = 0; s->resid = 0; s->resid0 = 0; s->reltol = 0;
}

// dinv = 1 / (diag + sigma * D)   (Jacobi preconditioner of A + sigma*D)
__global__ void k_make_dinv(int64_t n, const double *__restrict__ diag, const double *__restrict__ Dvec,
                            double sigma, double *__restrict__ dinv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dinv[i] = 1.0 / (diag[i] + sigma * (Dvec ? Dvec[i] : 1.0));
}

// r = rhs - c (c = A x0, when have_x0) or r = rhs, x = 0; sums (dinv r).r and r.r
// SC (cold start only): dinv carries s = diag^-1/2; r^ = s .* rhs, x^ = 0; sums r^.r^ and rhs.rhs
template <bool SC>
__global__ void __launch_bounds__(kBlock)
k_pcg_init(int64_t n, const double *__restrict__ rhs, const double *__restrict__ c, int have_x0,
           const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
           double *partials, unsigned int *ticket, PcgScal *scal, int finalize_mode, PeerRed pr) {
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double ri;
    if (!SC && have_x0) ri = rhs[i] - c[i];
    else { ri = rhs[i]; x[i] = 0.0; }
    if (SC) {
      const double rs = dinv[i] * ri;
      r[i] = rs;
      s0 += rs * rs;
    } else {
      r[i] = ri;
      s0 += dinv[i] * ri * ri;
    }
    s1 += ri * ri;
  }
  s0 = block_sum(s0);
  s1 = block_sum(s1);
  double t0, t1;
  if (last_block_sum2_all(s0, s1, partials, ticket, &t0, &t1) && threadIdx.x < 32) {
    double v[2] = {t0, t1};  // valid in lane 0
    const bool ok = finalize_mode != 2 || peer_allreduce_warp(pr, v, 2, scal);  // 2: sum over the ranks here
    if (threadIdx.x == 0) {
      scal->red[0] = v[0]; scal->red[1] = v[1];
      if (ok && finalize_mode >= 1) pcg_finish_init(scal, v[0], v[1]);
    }
  }
}

// x += alpha_prev * u (the deferred update of the previous iteration); u = dinv .* r + beta * u
// SC: u^ = r^ + beta * u^ (dinv is not read)
// fp.npeers > 0 (slab partitions over NVLink peer memory): the thread that produces u[i] of a boundary row also
// stores it into the neighbour's halo slot, and the CTA drawing the last ticket raises the neighbours' flags --
// the halo exchange of the next product rides on this kernel instead of a launch of its own (peer.cuh).
template <bool SC, bool PUSH>
__global__ void __launch_bounds__(kBlock)
k_update_u(int64_t n, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ u,
           double *__restrict__ x, const PcgScal *__restrict__ scal, FusedPush fp, unsigned int *ticket) {
  if (scal->done) {
    // iterations enqueued past convergence do no work, but EVERY halo sequence number must still be raised on the
    // neighbours: a consumer that is not the fused SpMV (k_halo_wait before the per-thread-load or CSR kernels)
    // waits for it unconditionally
    if (PUSH && blockIdx.x == 0 && (int)threadIdx.x < fp.npeers) st_release_sys(fp.flag[threadIdx.x], fp.seq);
    return;
  }
  const bool first = scal->iter == 0;
  const double beta = first ? 0.0 : scal->rho / scal->rho_prev;
  const double ap = scal->alpha_prev;
  // (PUSH = false is the single-GPU instantiation: exactly the two plain streaming loops)
  auto push = [&](int64_t i, double un) {
    if (PUSH) {
      const long long k0 = i - fp.begin[0];
      if (k0 >= 0 && k0 < fp.count[0]) fp.dst[0][k0] = un;
      if (fp.npeers > 1) {
        const long long k1 = i - fp.begin[1];
        if (k1 >= 0 && k1 < fp.count[1]) fp.dst[1][k1] = un;
      }
    }
  };
  if (first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double un = SC ? r[i] : dinv[i] * r[i];
      u[i] = un;
      push(i, un);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double ui = u[i];
      x[i] += ap * ui;
      const double un = (SC ? r[i] : dinv[i] * r[i]) + beta * ui;
      u[i] = un;
      push(i, un);
    }
  }
  if (PUSH) {
    __threadfence_system();  // my remote stores are visible system-wide before the ticket
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && (int)threadIdx.x < fp.npeers) {
      __threadfence_system();
      st_release_sys(fp.flag[threadIdx.x], fp.seq);
    }
  }
}

// r -= alpha c ; sums (dinv r).r and r.r ; last block closes the iteration
// SC: dinv carries diag(A): sums r^.r^ (= r.D^-1 r) and diag .* r^ . r^ (= r.r, the reference's residual)
template <bool SC>
__global__ void __launch_bounds__(kBlock)
k_update_xr(int64_t n, const double *__restrict__ c, const double *__restrict__ dinv, double *__restrict__ r,
            double *partials, unsigned int *ticket, PcgScal *scal, double *hist, int finalize_mode, PeerRed pr) {
  if (scal->done) return;
  const double alpha = scal->rho / scal->uc;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i] - alpha * c[i];
    r[i] = ri;
    if (SC) {
      s0 += ri * ri;
      s1 += dinv[i] * ri * ri;
    } else {
      s0 += dinv[i] * ri * ri;
      s1 += ri * ri;
    }
  }
  s0 = block_sum(s0);
  s1 = block_sum(s1);
  double t0, t1;
  if (last_block_sum2_all(s0, s1, partials, ticket, &t0, &t1) && threadIdx.x < 32) {
    double v[2] = {t0, t1};  // valid in lane 0
    const bool ok = finalize_mode != 2 || peer_allreduce_warp(pr, v, 2, scal);
    if (threadIdx.x == 0) {
      scal->red[0] = v[0]; scal->red[1] = v[1];
      if (ok && finalize_mode >= 1) pcg_finish_iter(scal, v[0], v[1], hist);
    }
  }
}

// after the loop: the x update of the last closed iteration (alpha_prev = 0 if none)
__global__ void __launch_bounds__(kBlock)
k_finish_x(int64_t n, const double *__restrict__ u, double *__restrict__ x, PcgScal *scal) {
  const double ap = scal->alpha_prev;
  if (ap == 0.0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] += ap * u[i];
}
// scaled variant: x = s .* (x^ + alpha_prev u^)
__global__ void __launch_bounds__(kBlock)
k_finish_x_scaled(int64_t n, const double *__restrict__ u, const double *__restrict__ sinv, double *__restrict__ x,
                  PcgScal *scal) {
  const double ap = scal->alpha_prev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = sinv[i] * (ap == 0.0 ? x[i] : x[i] + ap * u[i]);
}

// ---- small vector helpers for the transient path (src/transient.jl:71, :81) ---------------
// out = b + scale * (D ? D .* u : u)
__global__ void k_axpby_D(int64_t n, const double *__restrict__ b, const double *__restrict__ u,
                          const double *__restrict__ Dvec, double scale, double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = b[i] + scale * (Dvec ? Dvec[i] * u[i] : u[i]);
}
// out = D ? (mul ? D .* v : v ./ D) : v
__global__ void k_scale_D(int64_t n, const double *__restrict__ v, const double *__restrict__ Dvec, int mul,
                          double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = Dvec ? (mul ? Dvec[i] * v[i] : v[i] / Dvec[i]) : v[i];
}
// y = alpha * c + beta * y
__global__ void k_axpby(int64_t n, double alpha, const double *__restrict__ c, double beta, double *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = beta == 0.0 ? alpha * c[i] : alpha * c[i] + beta * y[i];
}
// sum (a-b)^2 -> red[0]
__global__ void __launch_bounds__(kBlock)
k_diffnorm2(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *partials,
            unsigned int *ticket, double *out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    s += d * d;
  }
  s = block_sum(s);
  double t;
  if (last_block_sum1(s, partials, ticket, &t)) *out = t;
}
// D_r = Ss * volumes[node(r)]
__global__ void k_make_D(int64_t nf, const int *__restrict__ row2node, const double *__restrict__ vol, double Ss,
                         double *__restrict__ D) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x)
    D[i] = Ss * vol[row2node[i]];
}

}  // namespace fvb
