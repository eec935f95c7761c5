/*
 * fv_oracle.c -- CPU restatement of FiniteVolume.jl's assemble -> solve path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under finitevolume.jl_b200/ may import,
 * link or call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, and nowhere else.
 *
 * Parity status: the solution-level behaviour is pinned by the reference's own
 * known-answer tests (test/runtests.jl:16, test/theis.jl:62,65,
 * test/onenodeadjoint.jl:33,43 -- see tests/test_oracle_pins.py).  The assembled
 * CSR structure/values are asserted by NO reference test and Julia is not
 * installed here, so at the CSR level this oracle is "parity unpinned": it is a
 * faithful restatement of src/FiniteVolume.jl:75-139 plus the published
 * SparseArrays.sparse!/IterativeSolvers.cg algorithms (un-vendored upstream
 * dependencies: Julia stdlib SparseArrays (Julia 1.1), IterativeSolvers 0.8.1,
 * both pinned only in /root/reference/Manifest.toml).
 *
 * All index arrays are int64 and 1-based on the wire, exactly like the Julia
 * arrays (`neighbors::Array{Pair{Int,Int},1}` is 2F interleaved int64).
 * Compile with -ffp-contract=off: Julia never fuses a*b+c on its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

/* ---- src/FiniteVolume.jl:32-44  getfreenodes ------------------------------ */
/* freenode[i] = i not in dirichletnodes; nodei2freenodei = 1-based running rank
 * of free nodes, -1 for Dirichlet nodes.  Returns sum(freenode). */
i64 fvo_getfreenodes(i64 n, const i64 *dnodes, i64 nd, uint8_t *freenode,
                     i64 *nodei2freenodei) {
  for (i64 i = 0; i < n; ++i) freenode[i] = 1;
  for (i64 k = 0; k < nd; ++k) freenode[dnodes[k] - 1] = 0;
  i64 j = 1;
  for (i64 i = 0; i < n; ++i) {
    if (freenode[i]) nodei2freenodei[i] = j++;
    else nodei2freenodei[i] = -1;
  }
  return j - 1;
}

/* ---- src/FiniteVolume.jl:20-30  getnodei2dirichleti ----------------------- */
/* Last duplicate wins.  Returns 0, or the (1-based) offending node when a source
 * sits on a Dirichlet node (the reference throws, :25-27). */
i64 fvo_getnodei2dirichleti(i64 n, const double *sources, const i64 *dnodes,
                            i64 nd, i64 *nodei2dirichleti) {
  for (i64 i = 0; i < n; ++i) nodei2dirichleti[i] = -1;
  for (i64 k = 0; k < nd; ++k) {
    i64 node = dnodes[k];
    nodei2dirichleti[node - 1] = k + 1;
    if (sources[node - 1] != 0) return node;
  }
  return 0;
}

static inline double face_conductance(const double *cond, const i64 *meta,
                                      const double *aol, i64 i, int logk) {
  /* src/FiniteVolume.jl:83 / :96 -- conductivities[metaindex(i)] * aol[i],
   * with exp() applied first when logtransformconductivity. */
  double k = cond[meta ? meta[i] - 1 : i];
  if (logk) k = exp(k);
  return k * aol[i];
}

/* ---- src/FiniteVolume.jl:75-106  assembleA, COO part ---------------------- */
/* Emits the (I,J,V) triples in exactly the push order of the reference loop.
 * I,J,V must hold 4*F entries.  Returns the number of triples. */
i64 fvo_assembleA_coo(i64 F, const i64 *neighbors, const double *aol,
                      const double *cond, const i64 *meta, int logk,
                      const uint8_t *freenode, const i64 *nodei2freenodei,
                      i64 *I, i64 *J, double *V) {
  i64 m = 0;
  for (i64 i = 0; i < F; ++i) {
    i64 n1 = neighbors[2 * i] - 1, n2 = neighbors[2 * i + 1] - 1;
    if (freenode[n1] && freenode[n2]) {
      double c = face_conductance(cond, meta, aol, i, logk);
      i64 r1 = nodei2freenodei[n1], r2 = nodei2freenodei[n2];
      I[m] = r1; J[m] = r1; V[m++] = c;
      I[m] = r1; J[m] = r2; V[m++] = -c;
      I[m] = r2; J[m] = r2; V[m++] = c;
      I[m] = r2; J[m] = r1; V[m++] = -c;
    } else if (freenode[n1]) {
      i64 r1 = nodei2freenodei[n1];
      I[m] = r1; J[m] = r1; V[m++] = face_conductance(cond, meta, aol, i, logk);
    } else if (freenode[n2]) {
      i64 r2 = nodei2freenodei[n2];
      I[m] = r2; J[m] = r2; V[m++] = face_conductance(cond, meta, aol, i, logk);
    }
  }
  return m;
}

/* ---- SparseArrays.sparse(I,J,V,m,n,+)  (called at src/FiniteVolume.jl:107) - */
/* Restatement of the stdlib's sparse! (Julia 1.1):
 *  1. count entries per row, 2. stable counting sort into a CSR scratch,
 *  3. sweep each row in input order, folding an entry into the first-seen slot
 *     of its column (left fold, `combine(old,new)`), explicit zeros kept,
 *  4. counting sort CSR -> CSC, which leaves row indices ascending per column.
 * colptr has n+1 entries, rowval/nzval must hold ncoo entries; all 1-based.
 * Returns nnz. */
i64 fvo_sparse(i64 m, i64 n, i64 ncoo, const i64 *I, const i64 *J,
               const double *V, i64 *colptr, i64 *rowval, double *nzval) {
  i64 *csrrowptr = (i64 *)calloc((size_t)m + 2, sizeof(i64));
  i64 *csrcolval = (i64 *)malloc((size_t)(ncoo ? ncoo : 1) * sizeof(i64));
  double *csrnzval = (double *)malloc((size_t)(ncoo ? ncoo : 1) * sizeof(double));
  i64 *klasttouch = (i64 *)calloc((size_t)n + 1, sizeof(i64));
  /* 1+2: stable counting sort by row */
  for (i64 k = 0; k < ncoo; ++k) csrrowptr[I[k] + 1]++;
  /* csrrowptr[r+1] = count(row r), rows 1-based; prefix into start offsets */
  {
    i64 acc = 0;
    for (i64 r = 1; r <= m; ++r) { i64 c = csrrowptr[r + 1]; csrrowptr[r + 1] = acc; acc += c; }
  }
  /* csrrowptr[r+1] now = start of row r (0-based offset); fill advances it */
  for (i64 k = 0; k < ncoo; ++k) {
    i64 p = csrrowptr[I[k] + 1]++;
    csrcolval[p] = J[k];
    csrnzval[p] = V[k];
  }
  /* after the fill csrrowptr[r+1] = end of row r = start of row r+1; start of
   * row r is csrrowptr[r] (csrrowptr[1] = 0 as calloc'ed) */
  /* 3: combine duplicates in place, count per-column entries */
  for (i64 j = 0; j <= n; ++j) colptr[j] = 0;
  i64 writek = 0;
  i64 *newrowptr = (i64 *)malloc(((size_t)m + 2) * sizeof(i64));
  for (i64 r = 1; r <= m; ++r) {
    i64 lo = csrrowptr[r], hi = csrrowptr[r + 1];
    i64 rowstart = writek;
    newrowptr[r] = rowstart;
    for (i64 k = lo; k < hi; ++k) {
      i64 j = csrcolval[k];
      if (klasttouch[j] > rowstart) { /* stored as pos+1, so > rowstart means seen in this row */
        csrnzval[klasttouch[j] - 1] = csrnzval[klasttouch[j] - 1] + csrnzval[k];
      } else {
        klasttouch[j] = writek + 1;
        csrcolval[writek] = j;
        csrnzval[writek] = csrnzval[k];
        writek++;
        colptr[j]++; /* count, 1-based column j in slot j */
      }
    }
  }
  newrowptr[m + 1] = writek;
  /* 4: CSR -> CSC counting sort */
  {
    i64 acc = 1; /* 1-based pointers */
    /* colptr[j] (j=1..n) holds count of column j; convert to 1-based starts in colptr[0..n] */
    i64 *cnt = (i64 *)malloc(((size_t)n + 1) * sizeof(i64));
    for (i64 j = 1; j <= n; ++j) cnt[j] = colptr[j];
    for (i64 j = 1; j <= n; ++j) { colptr[j - 1] = acc; acc += cnt[j]; }
    colptr[n] = acc;
    /* next-free cursor per column */
    for (i64 j = 1; j <= n; ++j) cnt[j] = colptr[j - 1];
    for (i64 r = 1; r <= m; ++r) {
      for (i64 k = newrowptr[r]; k < newrowptr[r + 1]; ++k) {
        i64 j = csrcolval[k];
        i64 p = cnt[j]++;
        rowval[p - 1] = r;
        nzval[p - 1] = csrnzval[k];
      }
    }
    free(cnt);
  }
  free(newrowptr);
  free(klasttouch);
  free(csrnzval);
  free(csrcolval);
  free(csrrowptr);
  return writek;
}

/* ---- src/FiniteVolume.jl:110-139  assembleb -------------------------------- */
void fvo_assembleb(i64 n, i64 F, const i64 *neighbors, const double *aol,
                   const double *cond, const i64 *meta, int logk,
                   const double *sources, const double *dheads,
                   const uint8_t *freenode, const i64 *nodei2freenodei,
                   const i64 *nodei2dirichleti, double *b) {
  i64 j = 0;
  for (i64 i = 0; i < n; ++i)
    if (freenode[i]) b[j++] = sources[i];
  for (i64 i = 0; i < F; ++i) {
    i64 n1 = neighbors[2 * i] - 1, n2 = neighbors[2 * i + 1] - 1;
    if (freenode[n1] && !freenode[n2]) {
      /* (k*aol)*head, left to right (:133) */
      double t = face_conductance(cond, meta, aol, i, logk) * dheads[nodei2dirichleti[n2] - 1];
      b[nodei2freenodei[n1] - 1] = b[nodei2freenodei[n1] - 1] + t;
    } else if (!freenode[n1] && freenode[n2]) {
      double t = face_conductance(cond, meta, aol, i, logk) * dheads[nodei2dirichleti[n1] - 1];
      b[nodei2freenodei[n2] - 1] = b[nodei2freenodei[n2] - 1] + t;
    }
  }
}

/* ---- src/FiniteVolume.jl:141-155  freenodes2nodes -------------------------- */
void fvo_freenodes2nodes(i64 n, const double *result, const double *dheads,
                         const uint8_t *freenode, const i64 *nodei2dirichleti,
                         double *head) {
  i64 sofar = 0;
  for (i64 i = 0; i < n; ++i) {
    if (freenode[i]) head[i] = result[sofar++];
    else head[i] = dheads[nodei2dirichleti[i] - 1];
  }
}

/* ---- src/grid.jl:56-110  regulargrid --------------------------------------- */
/* xs = range(min; stop=max, length=n): element i is computed here as
 * min + (i-1)*(max-min)/(n-1); dx = xs[2]-xs[1].  (Julia's StepRangeLen uses
 * twice-precision arithmetic; for the grids used in the tests -- integer or
 * exactly representable spacings -- both give the same doubles.)
 * coords: 3 x N column-major; neighbors: 2F interleaved 1-based; returns F. */
i64 fvo_regulargrid(const double *mins, const double *maxs, const i64 *ns,
                    double *coords, i64 *neighbors, double *aol, double *volumes) {
  i64 n1 = ns[0], n2 = ns[1], n3 = ns[2];
  double dx = (mins[0] + 1 * (maxs[0] - mins[0]) / (double)(n1 - 1)) - mins[0];
  double dy = (mins[1] + 1 * (maxs[1] - mins[1]) / (double)(n2 - 1)) - mins[1];
  double dz = (mins[2] + 1 * (maxs[2] - mins[2]) / (double)(n3 - 1)) - mins[2];
  i64 j = 0, v = 0;
  for (i64 i1 = 1; i1 <= n1; ++i1) {
    double areadx = dx;
    if (i1 == 1 || i1 == n1) areadx *= 0.5;
    for (i64 i2 = 1; i2 <= n2; ++i2) {
      double aready = dy;
      if (i2 == 1 || i2 == n2) aready *= 0.5;
      for (i64 i3 = 1; i3 <= n3; ++i3) {
        double areadz = dz;
        if (i3 == 1 || i3 == n3) areadz *= 0.5;
        i64 lin = i3 + n3 * (i2 - 1) + n3 * n2 * (i1 - 1);
        volumes[v++] = areadx * aready * areadz;
        if (coords) {
          coords[3 * (lin - 1) + 0] = (i1 == n1) ? maxs[0] : mins[0] + (double)(i1 - 1) * (maxs[0] - mins[0]) / (double)(n1 - 1);
          coords[3 * (lin - 1) + 1] = (i2 == n2) ? maxs[1] : mins[1] + (double)(i2 - 1) * (maxs[1] - mins[1]) / (double)(n2 - 1);
          coords[3 * (lin - 1) + 2] = (i3 == n3) ? maxs[2] : mins[2] + (double)(i3 - 1) * (maxs[2] - mins[2]) / (double)(n3 - 1);
        }
        if (i1 < n1) {
          neighbors[2 * j] = lin; neighbors[2 * j + 1] = lin + n3 * n2;
          aol[j] = aready * areadz / dx; ++j;
        }
        if (i2 < n2) {
          neighbors[2 * j] = lin; neighbors[2 * j + 1] = lin + n3;
          aol[j] = areadx * areadz / dy; ++j;
        }
        if (i3 < n3) {
          neighbors[2 * j] = lin; neighbors[2 * j + 1] = lin + 1;
          aol[j] = areadx * aready / dz; ++j;
        }
      }
    }
  }
  return j;
}

/* ---- src/grid.jl:14-33  nodehycos2neighborhycos ---------------------------- */
/* nodehycos is (n3,n2,n1) column-major == node-index order, so
 * nodehycos[multiindex(k)...] == flat[k-1]. */
void fvo_nodehycos2neighborhycos(i64 F, const i64 *neighbors, const double *nodehycos,
                                 int logtransform, double *out) {
  for (i64 i = 0; i < F; ++i) {
    double a = nodehycos[neighbors[2 * i] - 1], b = nodehycos[neighbors[2 * i + 1] - 1];
    out[i] = logtransform ? 0.5 * (a + b) : sqrt(a * b);
  }
}

/* ---- y = A x for the CSC arrays of a SparseMatrixCSC (mul!) ---------------- */
/* Column-oriented scatter, as Julia's mul!(y, A::SparseMatrixCSC, x). */
void fvo_spmv_csc(i64 n, const i64 *colptr, const i64 *rowval, const double *nzval,
                  const double *x, double *y) {
  for (i64 i = 0; i < n; ++i) y[i] = 0.0;
  for (i64 j = 0; j < n; ++j) {
    double xj = x[j];
    for (i64 k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) y[rowval[k] - 1] += nzval[k] * xj;
  }
}

/* Row-oriented gather using the same arrays as CSR (valid when A is symmetric in
 * structure AND value, which assembleA guarantees; also used, with the arrays
 * taken as CSR, for row-scaled transient matrices by the Python driver).
 * OpenMP variant for the "all host cores" courtesy baseline. */
void fvo_spmv_csr(i64 n, const i64 *rowptr, const i64 *colval, const double *nzval,
                  const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i) {
    double s = 0.0;
    for (i64 k = rowptr[i] - 1; k < rowptr[i + 1] - 1; ++k) s += nzval[k] * x[colval[k] - 1];
    y[i] = s;
  }
}

static double dot_(i64 n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (i64 i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ---- IterativeSolvers.cg / cg! 0.8.1 (called at src/FiniteVolume.jl:161,
 *      src/transient.jl:52,55) with Pl = Jacobi (diag(A)^-1) or identity -------
 * Recurrence and stopping rule restated from the published package:
 *   r = b - A x0 (x0 = 0 for cg); residual0 = ||r||; reltol = residual0 * tol
 *   while iteration < maxiter && residual > reltol:
 *     c = Pl \ r; rho_prev = rho; rho = c.r; beta = rho/rho_prev (first: u = c)
 *     u = c + beta u;  c = A u;  alpha = rho / (u.c)
 *     x += alpha u;  r -= alpha c;  residual = ||r||;  log residual
 * use_row_gather: 1 = treat arrays as CSR (threaded), 0 = CSC scatter (serial,
 * exactly what Julia's mul! does).  Returns the number of iterations;
 * *converged = residual <= reltol. */
i64 fvo_pcg(i64 n, const i64 *ptr, const i64 *idx, const double *val,
            const double *b, double *x, int have_x0, int jacobi, double tol,
            i64 maxiter, int use_row_gather, double *resnorm_hist, i64 hist_cap,
            int *converged) {
  double *r = (double *)malloc((size_t)n * sizeof(double));
  double *c = (double *)malloc((size_t)n * sizeof(double));
  double *u = (double *)calloc((size_t)n, sizeof(double));
  double *dinv = NULL;
  if (jacobi) {
    dinv = (double *)malloc((size_t)n * sizeof(double));
    for (i64 i = 0; i < n; ++i) {
      double d = 0.0;
      for (i64 k = ptr[i] - 1; k < ptr[i + 1] - 1; ++k)
        if (idx[k] - 1 == i) d = val[k];
      dinv[i] = 1.0 / d;
    }
  }
  if (have_x0) {
    if (use_row_gather) fvo_spmv_csr(n, ptr, idx, val, x, c);
    else fvo_spmv_csc(n, ptr, idx, val, x, c);
    for (i64 i = 0; i < n; ++i) r[i] = b[i] - c[i];
  } else {
    for (i64 i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; }
  }
  double residual = sqrt(dot_(n, r, r));
  double reltol = residual * tol;
  double rho = 1.0;
  i64 it = 0;
  while (it < maxiter && residual > reltol) {
    if (jacobi) {
#pragma omp parallel for schedule(static)
      for (i64 i = 0; i < n; ++i) c[i] = dinv[i] * r[i];
    } else {
      memcpy(c, r, (size_t)n * sizeof(double));
    }
    double rho_prev = rho;
    rho = dot_(n, c, r);
    double beta = rho / rho_prev;
    if (it == 0) beta = 0.0; /* u starts at zero, so u = c regardless */
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) u[i] = c[i] + beta * u[i];
    if (use_row_gather) fvo_spmv_csr(n, ptr, idx, val, u, c);
    else fvo_spmv_csc(n, ptr, idx, val, u, c);
    double alpha = rho / dot_(n, u, c);
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) { x[i] += alpha * u[i]; r[i] -= alpha * c[i]; }
    residual = sqrt(dot_(n, r, r));
    if (it < hist_cap && resnorm_hist) resnorm_hist[it] = residual;
    ++it;
  }
  *converged = residual <= reltol;
  free(dinv); free(u); free(c); free(r);
  return it;
}

int fvo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void fvo_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
