// reduce.cuh -- deterministic block / grid reductions for the CG scalars.
// Warp shuffles inside a warp, a fixed tree across warps, and a "last block folds the
// per-block partials in block order" finish, so dot products are reproducible run to run
// without floating-point atomics.
#pragma once
#include "common.cuh"

namespace fvb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA (blockDim.x == kBlock); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double ws[kBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect ws against a previous call
  if (lane == 0) ws[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = lane < kBlock / 32 ? ws[lane] : 0.0;
    s = warp_sum(s);
  }
  return s;
}

// Thread 0 of every block deposits its partial; the block that draws the last ticket
// re-reads all partials in block order and returns true (in thread 0) with the total.
// atomicInc wraps the ticket to 0, so it is ready for the next launch.
__device__ __forceinline__ bool last_block_sum1(double mine, double *partials, unsigned int *ticket,
                                                double *total) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = mine;
    __threadfence();
    unsigned int old = atomicInc(ticket, gridDim.x - 1);
    is_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double s = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(&partials[i]);
  s = block_sum(s);
  *total = s;
  return threadIdx.x == 0;
}

__device__ __forceinline__ bool last_block_sum2(double a, double b, double *partials, unsigned int *ticket,
                                                double *ta, double *tb) {
  __shared__ bool is_last2;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = a;
    partials[gridDim.x + blockIdx.x] = b;
    __threadfence();
    unsigned int old = atomicInc(ticket, gridDim.x - 1);
    is_last2 = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last2) return false;
  __threadfence();
  double s = 0.0, q = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    s += __ldcg(&partials[i]);
    q += __ldcg(&partials[gridDim.x + i]);
  }
  s = block_sum(s);
  q = block_sum(q);
  *ta = s;
  *tb = q;
  return threadIdx.x == 0;
}

// Variants for kernels that finish their sums across the GPUs with the first warp (peer_base.cuh:
// peer_allreduce_warp): every thread of the block learns whether this block drew the last ticket; the totals are
// valid in thread 0.
__device__ __forceinline__ bool last_block_sum1_all(double mine, double *partials, unsigned int *ticket, double *total) {
  __shared__ bool is_last1a;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = mine;
    __threadfence();
    is_last1a = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last1a) return false;
  __threadfence();
  double s = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(&partials[i]);
  *total = block_sum(s);
  return true;
}
__device__ __forceinline__ bool last_block_sum2_all(double a, double b, double *partials, unsigned int *ticket,
                                                    double *ta, double *tb) {
  __shared__ bool is_last2a;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = a;
    partials[gridDim.x + blockIdx.x] = b;
    __threadfence();
    is_last2a = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last2a) return false;
  __threadfence();
  double s = 0.0, q = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    s += __ldcg(&partials[i]);
    q += __ldcg(&partials[gridDim.x + i]);
  }
  *ta = block_sum(s);
  *tb = block_sum(q);
  return true;
}

}  // namespace fvb
