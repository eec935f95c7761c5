/* c_abi_demo.c -- the drop-in boundary used from plain C (no Python, no C++): the reference's own
 * smoke test (test/runtests.jl:4-16: four nodes in a line, heads 1 and 0 at the ends, every face
 * given in both directions so that sparse(...,+) has duplicates to fold) followed by a small
 * regular box, through fvb_create / fvb_assemble / fvb_sizes / fvb_get_csr / fvb_get_b / fvb_solve.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -o c_abi_demo finitevolume.jl_b200/libfvb200.so \
 *       -Wl,-rpath,$PWD/finitevolume.jl_b200 -lm
 *
 * Exit codes: 0 ok, 3 no CUDA device (the library has no CPU fallback), 1 anything else. */
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

static int chain(fvb_handle h) {
  /* neighbors = [1=>2, 2=>1, 2=>3, 3=>2, 3=>4, 4=>3] as 2F interleaved 1-based int64 */
  const int64_t nb[12] = {1, 2, 2, 1, 2, 3, 3, 2, 3, 4, 4, 3};
  const double aol[6] = {1, 1, 1, 1, 1, 1}, k[6] = {1, 1, 1, 1, 1, 1}, src[4] = {0, 0, 0, 0};
  const int64_t dn[2] = {1, 4};
  const double dh[2] = {1.0, 0.0};
  CHECK(fvb_assemble(h, 4, 1, 4, 6, nb, aol, k, 6, NULL, 0, src, 2, dn, dh));
  int64_t nf = 0, nnz = 0;
  CHECK(fvb_sizes(h, &nf, &nnz, NULL, NULL, NULL));
  if (nf != 2 || nnz != 4) { fprintf(stderr, "sizes %lld %lld\n", (long long)nf, (long long)nnz); return 1; }
  int64_t colptr[3], rowval[4];
  double nzval[4], b[2], head[4], hist[16];
  CHECK(fvb_get_csr(h, colptr, rowval, nzval));
  CHECK(fvb_get_b(h, b));
  /* A = [4 -2; -2 4], b = [2, 0] (duplicates summed) */
  if (nzval[0] != 4 || nzval[1] != -2 || nzval[2] != -2 || nzval[3] != 4 || b[0] != 2 || b[1] != 0) return 1;
  int64_t iters = 0;
  int conv = 0;
  CHECK(fvb_solve(h, 1.4901161193847656e-08, 400, NULL, head, NULL, &iters, &conv, hist, 16));
  printf("chain: heads %.10f %.10f %.10f %.10f, %lld iterations, converged %d\n", head[0], head[1], head[2], head[3],
         (long long)iters, conv);
  if (!conv || fabs(head[1] - 2.0 / 3) > 1.5e-8 || fabs(head[2] - 1.0 / 3) > 1.5e-8 || head[0] != 1.0 || head[3] != 0.0) return 1;
  return 0;
}

/* n x n x n unit box built on the host in the reference's order (src/grid.jl:56-110: node index
 * k + n3*(j-1) + n2*n3*(i-1), faces pushed per node in +x, +y, +z order), homogeneous K. */
static int box(fvb_handle h, int n) {
  const int64_t N = (int64_t)n * n * n, F = 3 * (int64_t)n * n * (n - 1);
  int64_t *nb = malloc(sizeof(int64_t) * 2 * F), *dn = malloc(sizeof(int64_t) * 2 * n * n);
  double *aol = malloc(sizeof(double) * F), *k = malloc(sizeof(double) * F), *src = calloc(N, sizeof(double));
  double *dh = malloc(sizeof(double) * 2 * n * n), *head = malloc(sizeof(double) * N);
  if (!nb || !dn || !aol || !k || !src || !dh || !head) return 1;
  int64_t f = 0;
  for (int i = 1; i <= n; ++i)
    for (int j = 1; j <= n; ++j)
      for (int kk = 1; kk <= n; ++kk) {
        const int64_t me = kk + (int64_t)n * (j - 1) + (int64_t)n * n * (i - 1);
        if (i < n) { nb[2 * f] = me; nb[2 * f + 1] = me + (int64_t)n * n; ++f; }
        if (j < n) { nb[2 * f] = me; nb[2 * f + 1] = me + n; ++f; }
        if (kk < n) { nb[2 * f] = me; nb[2 * f + 1] = me + 1; ++f; }
      }
  if (f != F) return 1;
  for (int64_t i = 0; i < F; ++i) { aol[i] = 1.0; k[i] = 1e-5; }
  for (int64_t i = 0; i < (int64_t)n * n; ++i) {
    dn[i] = i + 1; dh[i] = 1.0;                                   /* x = 0 plane */
    dn[(int64_t)n * n + i] = N - (int64_t)n * n + i + 1; dh[(int64_t)n * n + i] = 0.0; /* x = n-1 plane */
  }
  CHECK(fvb_assemble(h, N, 1, N, F, nb, aol, k, F, NULL, 0, src, 2 * (int64_t)n * n, dn, dh));
  int64_t iters = 0;
  int conv = 0;
  CHECK(fvb_solve(h, 1e-12, 100000, NULL, head, NULL, &iters, &conv, NULL, 0));
  /* homogeneous K, planar Dirichlet data: the exact solution is linear in x */
  double err = 0;
  for (int i = 1; i <= n; ++i) {
    const double exact = 1.0 - (double)(i - 1) / (n - 1);
    const double got = head[(int64_t)n * n * (i - 1) + (int64_t)n * (n / 2) + n / 2];
    if (fabs(got - exact) > err) err = fabs(got - exact);
  }
  int fmt = 0, noff = 0, scaled = 0;
  CHECK(fvb_get_spmv_format(h, &fmt, &noff));
  CHECK(fvb_get_pcg_scaling(h, &scaled));
  printf("box %d^3: %lld iterations, converged %d, max |h - exact| = %.2e, spmv format %d (%d offsets), scaled %d\n", n,
         (long long)iters, conv, err, fmt, noff, scaled);
  free(nb); free(dn); free(aol); free(k); free(src); free(dh); free(head);
  return (conv && err < 1e-9) ? 0 : 1;
}

int main(int argc, char **argv) {
  fvb_handle h = NULL;
  printf("libfvb200 version %d\n", fvb_version());
  CHECK(fvb_create(0, &h));
  int rc = chain(h);
  if (!rc) rc = box(h, argc > 1 ? atoi(argv[1]) : 24);
  CHECK(fvb_destroy(h));
  if (!rc) printf("c_abi_demo ok\n");
  return rc;
}
