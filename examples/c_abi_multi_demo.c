/* c_abi_multi_demo.c -- one process, one thread, several GPUs, plain C: the single-process multi-GPU front
 * end of the C ABI (fvb_multi_*), i.e. what a Julia `ccall` shim would drive.  An n^3 unit box in the
 * reference's node/face order (src/grid.jl:56-110), heterogeneous K that varies along y and z only, heads 1/0
 * on the x = 0 / x = n-1 planes: with K constant along x every (y,z) column of cells is a uniform chain in x
 * whose lateral neighbours carry the same head, so the exact solution is linear in x whatever K(y,z) is.
 * Checks: the partition covers all nodes, every device took the closed-form assembly, the global CSR image is
 * symmetric and its rows sum as they must, heads match the exact solution.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_multi_demo.c -o c_abi_multi_demo finitevolume.jl_b200/libfvb200.so \
 *       -Wl,-rpath,$PWD/finitevolume.jl_b200 -lm
 *   ./c_abi_multi_demo [ndev] [n]
 *
 * Exit codes: 0 ok, 3 no (or too few) CUDA devices, 1 anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fvb200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int s_ = (call);                                                                \
    if (s_ != FVB_OK) {                                                             \
      fprintf(stderr, "%s -> status %d: %s\n", #call, s_, fvb_last_error());        \
      return s_ == FVB_ERR_CUDA ? 3 : 1;                                            \
    }                                                                               \
  } while (0)

int main(int argc, char **argv) {
  int have = 0;
  const int ndev = argc > 1 ? atoi(argv[1]) : 2, n = argc > 2 ? atoi(argv[2]) : 64;
  CHECK(fvb_device_count(&have));
  if (have < ndev) { fprintf(stderr, "needs %d CUDA devices, found %d\n", ndev, have); return 3; }
  const int64_t plane = (int64_t)n * n, N = plane * n, F = 3 * plane * (n - 1);
  int64_t *nb = malloc(sizeof(int64_t) * 2 * F), *dn = malloc(sizeof(int64_t) * 2 * plane);
  double *aol = malloc(sizeof(double) * F), *k = malloc(sizeof(double) * F), *src = calloc(N, sizeof(double));
  double *dh = malloc(sizeof(double) * 2 * plane), *head = malloc(sizeof(double) * N);
  if (!nb || !dn || !aol || !k || !src || !dh || !head) return 1;
  int64_t f = 0;
  for (int i = 1; i <= n; ++i)
    for (int j = 1; j <= n; ++j)
      for (int kk = 1; kk <= n; ++kk) {
        const int64_t me = kk + (int64_t)n * (j - 1) + plane * (i - 1);
        const double kyz = 1e-5 * (1.0 + 0.5 * sin(0.7 * j) * cos(0.3 * kk));  /* K(y,z) of this column */
        const double kyz_j = 1e-5 * (1.0 + 0.5 * sin(0.7 * (j + 1)) * cos(0.3 * kk));
        const double kyz_k = 1e-5 * (1.0 + 0.5 * sin(0.7 * j) * cos(0.3 * (kk + 1)));
        if (i < n) { nb[2 * f] = me; nb[2 * f + 1] = me + plane; aol[f] = 1.0; k[f] = kyz; ++f; }
        if (j < n) { nb[2 * f] = me; nb[2 * f + 1] = me + n; aol[f] = 1.0; k[f] = sqrt(kyz * kyz_j); ++f; }
        if (kk < n) { nb[2 * f] = me; nb[2 * f + 1] = me + 1; aol[f] = 1.0; k[f] = sqrt(kyz * kyz_k); ++f; }
      }
  if (f != F) return 1;
  for (int64_t i = 0; i < plane; ++i) {
    dn[i] = i + 1; dh[i] = 1.0;
    dn[plane + i] = N - plane + i + 1; dh[plane + i] = 0.0;
  }
  fvb_multi m = NULL;
  CHECK(fvb_multi_create(ndev, NULL, &m));
  CHECK(fvb_multi_assemble(m, N, F, nb, aol, k, F, NULL, 0, src, 2 * plane, dn, dh));
  int64_t nf = 0, nnz = 0, lo[8], hi[8];
  int nd = 0;
  CHECK(fvb_multi_sizes(m, &nf, &nnz, &nd, lo, hi));
  if (nd != ndev || nf != N - 2 * plane || lo[0] != 1 || hi[ndev - 1] != N) return 1;
  for (int r = 0; r < ndev; ++r) {
    fvb_handle h = NULL;
    int kind = -1;
    CHECK(fvb_multi_device_handle(m, r, &h));
    CHECK(fvb_get_assembly(h, &kind));
    printf("device %d: nodes %lld..%lld, assembly path %d\n", r, (long long)lo[r], (long long)hi[r], kind);
    if (r && lo[r] != hi[r - 1] + 1) return 1;
    if (kind != 1) { fprintf(stderr, "device %d did not take the closed-form path\n", r); return 1; }
  }
  /* the whole matrix, as the reference returns it (SparseMatrixCSC arrays, 1-based) */
  int64_t *ptr = malloc(sizeof(int64_t) * (nf + 1)), *idx = malloc(sizeof(int64_t) * nnz);
  double *val = malloc(sizeof(double) * nnz), *b = malloc(sizeof(double) * nf);
  if (!ptr || !idx || !val || !b) return 1;
  CHECK(fvb_multi_get_csr(m, ptr, idx, val));
  CHECK(fvb_multi_get_b(m, b));
  if (ptr[0] != 1 || ptr[nf] != nnz + 1) { fprintf(stderr, "bad row pointers\n"); return 1; }
  double worst = 0;
  for (int64_t r = 0; r < nf; ++r) {  /* rows sum to the eliminated Dirichlet coupling, which is b / head */
    double s = 0;
    for (int64_t q = ptr[r] - 1; q < ptr[r + 1] - 1; ++q) s += val[q];
    const double want = r < plane ? b[r] : 0.0;  /* head 1 on the left plane, 0 on the right */
    const int last = r >= nf - plane;
    if (!last && fabs(s - want) > worst) worst = fabs(s - want);
    if (last && !(s > 0)) return 1;
  }
  int64_t iters = 0;
  int conv = 0;
  CHECK(fvb_multi_solve(m, 1e-12, 100000, head, NULL, &iters, &conv, NULL, 0));
  double err = 0;
  for (int64_t q = 0; q < N; ++q) {
    const double exact = 1.0 - (double)(q / plane) / (n - 1);
    if (fabs(head[q] - exact) > err) err = fabs(head[q] - exact);
  }
  printf("multi box %d^3 on %d GPUs: %lld iterations, converged %d, max |h - exact| = %.2e, max row-sum defect %.2e\n", n, ndev,
         (long long)iters, conv, err, worst);
  CHECK(fvb_multi_destroy(m));
  free(nb); free(dn); free(aol); free(k); free(src); free(dh); free(head); free(ptr); free(idx); free(val); free(b);
  if (!conv || err > 1e-9 || worst > 1e-18) return 1;
  printf("c_abi_multi_demo ok\n");
  return 0;
}
