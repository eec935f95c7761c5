// grid.cuh -- device-side regulargrid / nodehycos2neighborhycos, bit-identical in
// ordering and values to the serial loops of src/grid.jl:56-110 and :14-33, so a grid never has to exist
// on the host (the neighbor list of 1024^3 is 48 GiB of Pair{Int64,Int64}).
//
// Face order of the reference: nodes in linear order i3 + n3*(i2-1) + n3*n2*(i1-1) (:60), each emitting
// +x, +y, +z faces when the neighbour exists (:91-105).  The number of faces emitted before node
// (i1,i2,i3) therefore has the closed form
//   (i1-1)*(n2*n3 + (n2-1)*n3 + n2*(n3-1)) + (i2-1)*(hx*n3 + n3 + (n3-1)) + (i3-1)*(hx + hy + 1),
// hx = [i1<n1], hy = [i2<n2], because every earlier plane/row/cell has its +x/+y/+z neighbour.
// One thread per emitting node writes its (up to three) faces at that offset.
// A slab (planes p_lo..p_hi) lists every face with an endpoint in those planes: the faces emitted by its
// own nodes plus the +x faces of plane p_lo-1 -- the same subsequence grid.py builds on the host.
#pragma once
#include "common.cuh"

namespace fvb {

struct GridDesc {
  long long n1, n2, n3;
  double dx, dy, dz;
  long long p_lo, p_hi;   // owned planes, 1-based inclusive
  long long e_lo;         // first emitting plane (p_lo-1 if it exists, else p_lo)
};

__device__ __forceinline__ long long faces_before(const GridDesc &G, long long i1, long long i2, long long i3) {
  const long long hx = i1 < G.n1, hy = i2 < G.n2;
  const long long pfull = G.n2 * G.n3 + (G.n2 - 1) * G.n3 + G.n2 * (G.n3 - 1);
  return (i1 - 1) * pfull + (i2 - 1) * (hx * G.n3 + G.n3 + (G.n3 - 1)) + (i3 - 1) * (hx + hy + 1);
}

__global__ void __launch_bounds__(kBlock)
k_regulargrid(GridDesc G, longlong2 *__restrict__ nb, double *__restrict__ aol, double *__restrict__ vol) {
  const long long plane = G.n2 * G.n3;
  const long long nodes = (G.p_hi - G.e_lo + 1) * plane;
  const bool halo_plane = G.e_lo < G.p_lo;
  // local face offset of the first owned node = x-faces of the halo plane
  const long long base_owned = faces_before(G, G.p_lo, 1, 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nodes; t += (long long)gridDim.x * blockDim.x) {
    const long long i1 = G.e_lo + t / plane;
    const long long rem = t % plane;
    const long long i2 = rem / G.n3 + 1, i3 = rem % G.n3 + 1;
    const long long lin = i3 + G.n3 * (i2 - 1) + plane * (i1 - 1);
    const double wx = (i1 == 1 || i1 == G.n1) ? G.dx * 0.5 : G.dx;
    const double wy = (i2 == 1 || i2 == G.n2) ? G.dy * 0.5 : G.dy;
    const double wz = (i3 == 1 || i3 == G.n3) ? G.dz * 0.5 : G.dz;
    if (i1 < G.p_lo) {  // halo plane: only its +x faces touch the slab
      const long long j = rem;
      nb[j] = make_longlong2(lin, lin + plane);
      aol[j] = __ddiv_rn(__dmul_rn(wy, wz), G.dx);
      continue;
    }
    if (vol) vol[(i1 - G.p_lo) * plane + rem] = __dmul_rn(__dmul_rn(wx, wy), wz);
    long long j = (halo_plane ? plane : 0) + faces_before(G, i1, i2, i3) - base_owned;
    if (i1 < G.n1) { nb[j] = make_longlong2(lin, lin + plane); aol[j] = __ddiv_rn(__dmul_rn(wy, wz), G.dx); ++j; }
    if (i2 < G.n2) { nb[j] = make_longlong2(lin, lin + G.n3); aol[j] = __ddiv_rn(__dmul_rn(wx, wz), G.dy); ++j; }
    if (i3 < G.n3) { nb[j] = make_longlong2(lin, lin + 1); aol[j] = __ddiv_rn(__dmul_rn(wx, wy), G.dz); ++j; }
  }
}

// src/grid.jl:14-33: geometric mean, or arithmetic mean of logs, of the two node values of each face.
// nodek holds nodes node_lo.. (1-based) in node order (the (n3,n2,n1) column-major array of the reference).
__global__ void __launch_bounds__(kBlock)
k_node2face(long long nf, const longlong2 *__restrict__ nb, const double *__restrict__ nodek, long long node_lo,
            long long n_have, int logmean, double *__restrict__ out, int *__restrict__ err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (long long)gridDim.x * blockDim.x) {
    const longlong2 p = nb[i];
    const long long a = p.x - node_lo, b = p.y - node_lo;
    if (a < 0 || b < 0 || a >= n_have || b >= n_have) { atomicMin(err, (int)min((long long)INT_MAX - 1, i)); continue; }
    const double ka = nodek[a], kb = nodek[b];
    out[i] = logmean ? __dmul_rn(0.5, __dadd_rn(ka, kb)) : sqrt(__dmul_rn(ka, kb));
  }
}

}  // namespace fvb
