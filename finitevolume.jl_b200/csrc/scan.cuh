// scan.cuh -- exclusive prefix sum of int32 arrays (three passes: per-chunk sums, scan of
// the chunk sums by one block, per-chunk rescan).  Replaces the serial running counters
// of getfreenodes (src/FiniteVolume.jl:36-42) and the counting passes of sparse! (:107).
// Totals are bounded by the caller (< 2^31), so int32 arithmetic is exact.
#pragma once
#include "common.cuh"

namespace fvb {

constexpr int kScanItems = 16;                        // per thread
constexpr int kScanChunk = kBlock * kScanItems;       // 4096 per block

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread across a kBlock-thread block; returns the
// exclusive prefix and writes the block total to *total (same for every thread).
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int wsum[kBlock / 32];
  __shared__ int tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = lane < kBlock / 32 ? wsum[lane] : 0;
    int si = warp_incl_scan(s);
    if (lane < kBlock / 32) wsum[lane] = si - s;
    if (lane == kBlock / 32 - 1) tot = si;
  }
  __syncthreads();
  int res = incl - v + wsum[w];
  *total = tot;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kBlock) k_scan_chunk_sums(const int *__restrict__ in, int64_t n,
                                                            int *__restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + (int64_t)j * kBlock + threadIdx.x;
    if (i < n) s += in[i];
  }
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// One block: exclusive scan of `sums` in place; sums[nchunks] = grand total.
__global__ void __launch_bounds__(kBlock) k_scan_sums(int *__restrict__ sums, int nchunks) {
  int carry = 0;
  for (int base = 0; base < nchunks; base += kBlock) {
    int i = base + threadIdx.x;
    int v = i < nchunks ? sums[i] : 0;
    int tot;
    int ex = block_excl_scan(v, &tot);
    if (i < nchunks) sums[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0) sums[nchunks] = carry;
}

// out[i] = exclusive prefix of in[0..i); out[n] = total.  in == out allowed.
// Each thread owns kScanItems CONSECUTIVE items so the running order is the array order.
__device__ __forceinline__ int scan_pad(int k) { return k + (k >> 5); }  // conflict-free stride-16

__global__ void __launch_bounds__(kBlock) k_scan_final(const int *in, int64_t n,
                                                       const int *__restrict__ sums, int *out) {
  __shared__ int tile[kScanChunk + kScanChunk / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int k = j * kBlock + threadIdx.x;
    int64_t i = base + k;
    tile[scan_pad(k)] = i < n ? in[i] : 0;
  }
  __syncthreads();
  int loc[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    loc[j] = tile[scan_pad(threadIdx.x * kScanItems + j)];
    s += loc[j];
  }
  int tot;
  int ex = block_excl_scan(s, &tot) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int v = loc[j];
    tile[scan_pad(threadIdx.x * kScanItems + j)] = ex;
    ex += v;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int k = j * kBlock + threadIdx.x;
    int64_t i = base + k;
    if (i < n) out[i] = tile[scan_pad(k)];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = sums[gridDim.x];
}

// Host driver.  `scratch` must hold cdiv(n, kScanChunk) + 1 ints.  out needs n+1 slots.
inline int64_t scan_scratch_ints(int64_t n) { return (n + kScanChunk - 1) / kScanChunk + 2; }

inline void exclusive_scan(const int *in, int64_t n, int *out, int *scratch, cudaStream_t st,
                           int64_t *launches) {
  if (n == 0) {
    cudaMemsetAsync(out, 0, sizeof(int), st);
    return;
  }
  int nchunks = (int)((n + kScanChunk - 1) / kScanChunk);
  k_scan_chunk_sums<<<nchunks, kBlock, 0, st>>>(in, n, scratch);
  k_scan_sums<<<1, kBlock, 0, st>>>(scratch, nchunks);
  k_scan_final<<<nchunks, kBlock, 0, st>>>(in, n, scratch, out);
  *launches += 3;
}

}  // namespace fvb
