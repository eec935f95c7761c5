/*
 * fvb200.h -- C ABI of libfvb200.so, the B200-native (sm_100a) replacement for the
 * assemble -> solve hot path of madsjulia/FiniteVolume.jl.
 *
 * The reference has no FFI layer of its own: the boundary it exposes is its Julia call
 * surface.  Every entry point below names the reference function (file:line, relative
 * to the reference checkout) whose work it takes over; INTEGRATION.md shows the
 * `ccall` shim (julia/FiniteVolumeB200.jl) that binds them under the reference's own
 * function names.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All index arrays on the wire are int64 and
 *    1-based, exactly as Julia lays them out (`neighbors::Array{Pair{Int,Int},1}` is
 *    2F interleaved int64: n1_1,n2_1,n1_2,n2_2,...).  Conversion to 0-based int32
 *    happens on the device.
 *  - Input pointers are borrowed for the duration of the call only and may be host
 *    (pageable or pinned) or device pointers (cudaMemcpyDefault is used throughout).
 *    Outputs are written into caller-allocated buffers (query sizes first).
 *  - All device memory lives behind the opaque handle.  One handle = one GPU = one owner
 *    thread at a time.  Multi-GPU = one process (or thread) per GPU, one handle each,
 *    joined by fvb_comm_init().
 *  - Every function returns an fvb_status; fvb_last_error() gives the thread-local
 *    message.  Non-convergence of the solver is NOT an error (it mirrors
 *    `ch.isconverged`, src/FiniteVolume.jl:161).
 *  - There is no CPU fallback: without a CUDA device fvb_create() fails.
 */
#ifndef FVB200_H
#define FVB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fvb_handle_s *fvb_handle;

typedef enum {
  FVB_OK = 0,
  FVB_ERR_BAD_INPUT = 1,  /* what the reference reports with error()/BoundsError */
  FVB_ERR_CUDA = 2,
  FVB_ERR_NCCL = 3,
  FVB_ERR_OOM = 4,
  FVB_ERR_STATE = 5       /* call order violated (e.g. solve before assemble) */
} fvb_status;

#define FVB_UNIQUE_ID_BYTES 128

/* ---- lifetime ------------------------------------------------------------------- */
int fvb_version(void);
const char *fvb_last_error(void);
int fvb_device_count(int *count);
/* Binds the handle to CUDA device `device` and creates its stream. */
int fvb_create(int device, fvb_handle *out);
int fvb_destroy(fvb_handle h);

/* ---- multi-GPU (one rank per GPU; SURVEY 8e: slab partition by node index) -------- */
/* Rank 0 calls fvb_comm_unique_id and ships the 128 bytes to the other ranks by any
 * means (torch.distributed broadcast in the Python harness, MPI/Distributed in Julia);
 * then every rank calls fvb_comm_init.  NCCL is dlopen'ed ("libnccl.so.2") on first use,
 * so single-GPU users need no NCCL at all. */
int fvb_comm_unique_id(uint8_t id[FVB_UNIQUE_ID_BYTES]);
int fvb_comm_init(fvb_handle h, int nranks, int rank, const uint8_t id[FVB_UNIQUE_ID_BYTES]);

/* ---- assembly: assembleA + assembleb (src/FiniteVolume.jl:75-108, :110-139) -------
 * with getfreenodes (:32-44) and getnodei2dirichleti (:20-30) folded in.
 *
 *  n_nodes      N = length(sources) of the WHOLE problem
 *  node_lo/hi   1-based inclusive range of nodes this rank owns (1, N for one GPU).
 *               Rows (free nodes in ascending node order) are owned with their node.
 *  n_faces      number of faces passed; they must be, in the global face order, ALL the
 *               faces with at least one endpoint in [node_lo, node_hi]
 *  neighbors    2*n_faces int64, interleaved pairs, global node ids
 *  aol, cond    areasoverlengths[n_faces]; conductivities[n_cond]
 *  metaindex    NULL for i -> i, else int64[n_faces], 1-based into cond (the table of the
 *               reference's `metaindex` callable, src/FiniteVolume.jl:75)
 *  logk         logtransformconductivity
 *  sources      the owned slice sources[node_lo..node_hi]
 *  dnodes/heads the FULL dirichletnodes / dirichletheads lists (every rank passes all)
 *
 * Produces, per owned row: the CSR row (columns ascending, duplicates left-folded in
 * face order, explicit zeros kept -- the semantics of SparseArrays.sparse(I,J,V,m,n,+)
 * at :107), b, and diag(A).  Fails with FVB_ERR_BAD_INPUT when a source sits on a
 * Dirichlet node (the reference's error(), :25-27) or a node id is out of range. */
int fvb_assemble(fvb_handle h, int64_t n_nodes, int64_t node_lo, int64_t node_hi,
                 int64_t n_faces, const int64_t *neighbors, const double *aol,
                 const double *cond, int64_t n_cond, const int64_t *metaindex, int logk,
                 const double *sources, int64_t n_dirichlet, const int64_t *dnodes,
                 const double *dheads);

/* Values-only re-assembly on the retained structure (inverse loops re-solve on a fixed
 * mesh: examples/box_model/ex.jl:53-64).  cond/metaindex/logk as in fvb_assemble;
 * sources/dheads may be NULL to keep the previous ones. */
int fvb_update_values(fvb_handle h, const double *cond, int64_t n_cond, int logk,
                      const double *sources, const double *dheads);

/* ---- closed-form assembly of regulargrid-ordered problems ------------------------------------
 * When the face list passed to fvb_assemble is exactly the one regulargrid emits for the owned
 * x-planes (src/grid.jl:56-110: nodes z-fastest, each pushing its +x, +y, +z faces) and the
 * Dirichlet set keeps the free nodes on the grid's three diagonals (e.g. the left/right planes of
 * examples/box_model/ex.jl:27-37), the library verifies that on the device and then writes the
 * rows straight into the symmetric-diagonal format: the faces of a node are known in closed form
 * and already in ascending face order (-x,-y,-z,+x,+y,+z), which is the order assembleA/sparse!
 * fold in (src/FiniteVolume.jl:94-107).  No adjacency and no CSR are built; fvb_get_csr produces
 * the CSR image on demand, bit-identical to the general path's.  Face-indexed arrays of this path
 * are addressed with 64-bit offsets, so only the NODE count of a rank must stay below 2^31.
 *   mode: 0 = automatic (default), 1 = always the general path.  FVB_BOX=0 presets 1.
 *   active: 0 = general path, 1 = closed form from the caller's arrays, 2 = grid-implicit. */
int fvb_set_assembly(fvb_handle h, int mode);
int fvb_get_assembly(fvb_handle h, int *active);

/* Grid-implicit assembly: the problem that
 *     coords, neighbors, aol, volumes = regulargrid(mins, maxs, ns)            (src/grid.jl:56)
 *     k = nodehycos2neighborhycos(neighbors, nodehycos, logmean)               (src/grid.jl:14)
 *     assembleA / assembleb(neighbors, aol, k, sources, dnodes, dheads, i->i, logk)
 * describes, without any per-face array ever existing (1024^3 has 3.2e9 faces: 51 GB of neighbor
 * pairs alone).  areasoverlengths come from the grid spacing exactly as regulargrid computes them,
 * face conductivities from the two node values exactly as nodehycos2neighborhycos does; rows are
 * bit-identical to fvb_regulargrid + fvb_nodehycos2neighborhycos + fvb_assemble.
 *   plane_lo/hi  owned x-planes, 1-based inclusive (1, ns[0] on one GPU); owned nodes = those planes
 *   nodehycos    node values, node order, of planes max(1,plane_lo-1) .. min(ns[0],plane_hi+1)
 *   sources      owned nodes, or NULL for all zero
 *   dnodes/heads the FULL Dirichlet lists
 * Fails with FVB_ERR_BAD_INPUT when the Dirichlet set does not fit the diagonal format (then build
 * the lists with fvb_regulargrid and use fvb_assemble).  fvb_update_values on such a problem takes
 * the node values (same planes) as `cond`.  The adjoint gradient gather needs face arrays and is
 * not available on it. */
int fvb_assemble_regulargrid(fvb_handle h, const double mins[3], const double maxs[3], const int64_t ns[3],
                             int64_t plane_lo, int64_t plane_hi, const double *nodehycos, int logmean, int logk,
                             const double *sources, int64_t n_dirichlet, const int64_t *dnodes,
                             const double *dheads);

/* Sizes of this rank's part: free rows owned, stored entries, first owned row
 * (1-based global free index), global free count, halo columns referenced. */
int fvb_sizes(fvb_handle h, int64_t *nf_local, int64_t *nnz_local, int64_t *row_start,
              int64_t *nf_global, int64_t *n_halo);

/* The owned rows as the CSC/CSR arrays of a SparseMatrixCSC{Float64,Int64} (A is
 * symmetric, so the same arrays serve as either): ptr[nf_local+1] 1-based into this
 * rank's idx/val, idx = GLOBAL 1-based column (= Julia rowval), val = nzval. */
int fvb_get_csr(fvb_handle h, int64_t *ptr, int64_t *idx, double *val);
int fvb_get_b(fvb_handle h, double *b);                 /* assembleb, owned rows     */
int fvb_get_diag(fvb_handle h, double *d);
int fvb_get_freenode(fvb_handle h, uint8_t *freenode);  /* owned nodes; :32-44       */
int fvb_get_nodei2freenodei(fvb_handle h, int64_t *map);/* owned nodes; -1 = Dirichlet */

/* ---- halo plan (multi-GPU only) ---------------------------------------------------- */
/* Global 1-based free indices of the off-rank columns this rank's rows reference,
 * ascending; x[nf_local + k] on the device holds the value of halo column k. */
int fvb_get_halo_cols(fvb_handle h, int64_t *cols);
/* For every peer p (ascending rank): this rank receives recv_counts[p] consecutive halo
 * entries from it (in halo order) and sends it the owned rows send_rows (0-based local
 * row ids, concatenated per peer, send_counts[p] each). */
int fvb_set_halo_plan(fvb_handle h, int n_peers, const int32_t *peer_ranks,
                      const int64_t *send_counts, const int32_t *send_rows,
                      const int64_t *recv_counts);

/* ---- adjoint gradient (src/transientadjointutils.jl:22-32 `dfdp(t)*lambda`, route A of
 * gradientintegrate, src/transient.jl:207-216): per-face gather of (df/dp)^T lambda for the ODE
 * right-hand side f = D^-1 (b - A u), D = Ss*volumes, accumulated over quadrature points.
 *   face i joining free rows r1,r2, c_i its conductance, dc = c_i (log K) or aol_i (plain K):
 *       d/dk[metaindex(i)]  +=  w * ( -dc (u1-u2) (l1/D1 - l2/D2) )
 *   face with r1 free and Dirichlet node d:
 *       d/dk[metaindex(i)]  +=  w * dc (h_d - u1) l1/D1 ;   d/dh_d += w * c_i l1/D1
 *   free row r:            d/dsources[node(r)] += w * l_r/D_r
 * Accumulators are per FACE and per ROW (one writer each: deterministic, no float atomics); the
 * host folds faces onto conductivity / Dirichlet-head indices in face order.
 * begin: pass the same neighbors list as fvb_assemble (it is not kept after assembly). */
int fvb_gradient_begin(fvb_handle h, const int64_t *neighbors);
int fvb_gradient_accumulate(fvb_handle h, int u_slot, int lambda_slot, double weight);
int fvb_gradient_end(fvb_handle h, double *grad_cond_face, double *grad_dhead_face, int64_t *dhead_slot_face,
                     double *grad_source_rows);

/* ---- NVLink peer-memory exchange (optional, same node, after fvb_set_halo_plan) --------------
 * Replaces the per-iteration NCCL calls (halo send/recv, two all-reduces) by kernels that store
 * straight into the neighbours' memory mapped through CUDA IPC.  Every rank calls
 * fvb_peer_export, the 128-byte blobs are all-gathered by the host harness, then every rank
 * calls fvb_peer_import with the blobs in rank order and, per peer of its halo plan, the index
 * in that peer's vector where this rank's first halo value belongs
 * (= peer's nf_local + number of the peer's halo columns owned by lower ranks).
 * Without these calls the NCCL path is used. */
#define FVB_PEER_BLOB_BYTES 128
int fvb_peer_export(fvb_handle h, uint8_t blob[FVB_PEER_BLOB_BYTES]);
int fvb_peer_import(fvb_handle h, const uint8_t *blobs_by_rank, const int64_t *send_dst_index);

/* ---- steady solve: the cg call of solvediffusion (src/FiniteVolume.jl:160-161) with
 * Pl = Jacobi instead of Ruge-Stueben AMG (north_star), followed by freenodes2nodes
 * (:141-155).  Stopping rule of IterativeSolvers.cg 0.8.1: ||r|| <= rtol * ||r0||.
 *  x0_free      NULL (cg, x0 = 0) or the owned slice of an initial guess (cg!)
 *  head_nodes   out, owned nodes [node_lo..node_hi]: solution on free nodes, prescribed
 *               head on Dirichlet nodes; may be NULL
 *  x_free       out, owned free rows; may be NULL
 *  resnorm_hist out, first min(iters, hist_cap) residual norms (ch.data[:resnorm]) */
int fvb_solve(fvb_handle h, double rtol, int64_t maxiter, const double *x0_free,
              double *head_nodes, double *x_free, int64_t *iters, int *converged,
              double *resnorm_hist, int64_t hist_cap);

/* y = alpha * A x + beta * y on owned rows (host or device vectors of nf_local).
 * Covers mul! inside cg, `b - A u` (src/transient.jl:197) and, A being symmetric, the
 * adjoint's transpose product (:193).  Multi-GPU: collective (halo exchange inside). */
int fvb_spmv(fvb_handle h, double alpha, const double *x, double beta, double *y);

/* ---- transient: device-resident backward Euler (src/transient.jl:65-76, :188-205) --
 * The handle keeps NSLOT work vectors of nf_local doubles so that the host-side step
 * controller (src/transient.jl:78-154) moves no vectors over PCIe per step. */
#define FVB_NSLOT 8
int fvb_vec_upload(fvb_handle h, int slot, const double *host);
int fvb_vec_download(fvb_handle h, int slot, double *host);
int fvb_vec_copy(fvb_handle h, int dst, int src);
int fvb_vec_load_b(fvb_handle h, int slot);              /* slot <- assembled b        */
/* ||a - b||_2 over ALL ranks (LinearAlgebra.norm(onestep - twostep), :81). */
int fvb_vec_diffnorm(fvb_handle h, int a, int b, double *out);
/* D_r = Ss * volumes[node(r)] (scalebyvolume!, :7-22): pass the owned slice of volumes
 * (node order, n_owned_nodes entries); NULL resets D to the identity. */
int fvb_set_storage(fvb_handle h, double Ss, const double *volumes_owned_nodes);
/* One backward-Euler solve, slot -> slot, in the SPD form of the reference's step:
 *  adjoint = 0:  (A + D/dt) u+ = rhs_b + D u/dt            (== (D^-1 A + I/dt) u+ = D^-1 b + u/dt, :71-73)
 *  adjoint = 1:  (A + D/dt) w  = rhs_b + u/dt, u+ = D w    (== (A D^-1 + I/dt) u+ = g + u/dt, :193,:203)
 * rhs_slot holds the UNSCALED b (or the adjoint forcing g); warm start from u as the
 * reference does (:73).  Returns FVB_ERR_BAD_INPUT for dt <= 0 (:68-70). */
int fvb_step(fvb_handle h, int rhs_slot, int u_slot, double dt, int out_slot, int adjoint,
             double rtol, int64_t maxiter, int64_t *iters, int *converged);
/* The `linearsolver(A, rhs, x0)` hook of backwardeulerintegrate (src/transient.jl:136, default :50-58) on the
 * resident matrix: solves (A + sigma * D) x = rhs_slot starting from x0_slot (cg!-style warm start, tolerance
 * relative to the initial residual), D as set by fvb_set_storage (identity if unset).  The reference hands the
 * hook A~ = D^-1 A + I/dt and rhs~ = D^-1 b + u/dt; the equivalent SPD system is (A + D/dt) x = D .* rhs~, so a
 * shim passes sigma = 1/dt and rhs = D .* rhs~ (julia/FiniteVolumeB200.jl: linearsolver). */
int fvb_solve_shifted(fvb_handle h, int rhs_slot, int x0_slot, double sigma, int out_slot, double rtol,
                      int64_t maxiter, int64_t *iters, int *converged);
/* ---- the whole integrator on the device side of the boundary (src/transient.jl:78-154) ----------
 * backwardeulerintegrate(u0, A, getb, dt0, t0, tfinal; stepper!, atol, callback) with
 * stepper! = adaptivebackwardeulerstep! (step doubling: one full step against two half steps,
 * err = ||one - two||_2 < atol accepts, < atol/4 doubles the step, else the half step is kept and the
 * step halved, :78-121) or fixedbackwardeulerstep! (:130-134).  State vectors never leave the device
 * except the accepted ones the caller asks for; on a single GPU and up to 2^20 unknowns with a
 * constant right-hand side every step-doubling attempt (three warm-started solves + the norm) is
 * one cooperative kernel launch (csrc/coop.cuh), otherwise one fvb_step per solve.
 * Call fvb_set_storage first (D = Ss*volumes).  All FVB_NSLOT vector slots are clobbered.
 *   getb       NULL: the assembled b, constant in time (the model-level entry, :156-163; adjoint:
 *              zero forcing); else called once per linear solve with the solve's time and must fill
 *              the UNSCALED right-hand side on the owned free rows (the reference's getb returns
 *              D^-1 b: multiply by Ss*volumes) -- for the adjoint the forcing dg/du(T - t) (:203)
 *   callback   NULL or called as callback(t, dt) once per attempted step, as the reference does
 *   ts         out, [max_states]: t0 and the time of every accepted step (ts[k+1] = ts[k] + dt, :145)
 *   us_free    out or NULL, [max_states * nf_local]: the accepted states on free rows (us, :144)
 *   heads_nodes out or NULL, [max_states * n_owned_nodes]: the same through freenodes2nodes (:172)
 * Fails with FVB_ERR_STATE when more than max_states states are accepted (n_states unchanged). */
typedef void (*fvb_getb_fn)(double t, double *b_free, void *ctx);
typedef void (*fvb_step_callback_fn)(double t, double dt, void *ctx);
typedef struct {
  double atol;       /* step-doubling tolerance (reference default 1e-4)         */
  double dt0;        /* first step (reference default 1.0)                        */
  int fixed_step;    /* 0 adaptivebackwardeulerstep!, 1 fixedbackwardeulerstep!   */
  int adjoint;       /* 0 forward (:65-76), 1 adjoint operator (:188-205)         */
  double rtol;       /* linear solves: relative to the warm start's residual     */
  int64_t maxiter;
  fvb_getb_fn getb;
  void *getb_ctx;
  fvb_step_callback_fn callback;
  void *callback_ctx;
} fvb_integrate_options;
int fvb_integrate(fvb_handle h, const double *u0_free, double t0, double tfinal, const fvb_integrate_options *opt,
                  int64_t max_states, double *ts, double *us_free, double *heads_nodes, int64_t *n_states,
                  int64_t *n_solves, int64_t *n_cg_iterations, int64_t *n_attempts);

/* Scatter a free-row slot to owned nodes (freenodes2nodes, src/transient.jl:172). */
int fvb_vec_to_nodes(fvb_handle h, int slot, double *head_nodes);

/* ---- SpMV storage format ---------------------------------------------------------------
 * After assembly the library checks whether every entry of A lies on at most 4 symmetric
 * diagonals (regulargrid numbering, src/grid.jl:60); if so the solver streams an index-free
 * symmetric-diagonal copy (8*(K+1)+16 bytes per row) instead of the CSR arrays
 * (12*nnz_row+20).  The CSR arrays stay resident either way (fvb_get_csr).
 * Two kernels serve the diagonal copy: a persistent TMA pipeline (cp.async.bulk slices of every
 * operand into shared memory, dia_tma.cuh) when almost all row tiles are interior to the owned
 * range, else a per-thread-load kernel (dia.cuh).  All three kernels give bit-identical products.
 *   fmt: 0 = automatic (default), 1 = always CSR, 2 = diagonal with per-thread loads only,
 *        3 = diagonal through the TMA pipeline whatever the size (2, 3: CSR if the pattern does
 *        not qualify).  The environment variable FVB_SPMV_FORMAT presets it at fvb_create.
 *   active: 1 = CSR, 2 = diagonal (per-thread loads), 3 = diagonal (TMA pipeline). */
int fvb_set_spmv_format(fvb_handle h, int fmt);
int fvb_get_spmv_format(fvb_handle h, int *active, int *n_offsets);

/* ---- symmetric Jacobi scaling of the steady solve ------------------------------------------
 * Jacobi-PCG on A x = b (src/FiniteVolume.jl:160-161 with Pl = diag(A)) generates, in exact
 * arithmetic, the same iterates as plain CG on A^ x^ = b^ with A^ = D^-1/2 A D^-1/2 (unit
 * diagonal), x^ = D^1/2 x, b^ = D^-1/2 b.  When the diagonal format is active the library keeps
 * such a scaled copy and runs cold-started steady solves (fvb_solve without x0) on it: no D^-1
 * and no diag(A) reads inside the iteration (112 instead of 128 bytes per row).  The residual
 * norm tested and recorded is still the reference's ||b - A x||_2.  Heads agree with the unscaled
 * recurrence to solver tolerance; iteration counts can differ by rounding (+-1).
 *   mode: 0 = automatic (default), 1 = never (always the unscaled recurrence).
 * fvb_get_pcg_scaling reports whether the LAST fvb_solve on this handle ran scaled. */
int fvb_set_pcg_scaling(fvb_handle h, int mode);
int fvb_get_pcg_scaling(fvb_handle h, int *last_solve_scaled);

/* ---- device-side grid helpers (src/grid.jl:56-110, :14-33) ---------------------------------
 * Same ordering and bit-identical values as the reference's serial loops, produced directly in
 * device memory (the host never holds the neighbor list).  Buffers are plain device allocations
 * owned by the caller through fvb_device_alloc/free; they can be passed to fvb_assemble as is.
 *   plane_lo/hi  1-based inclusive x-planes this rank owns (1, ns[0] for the whole grid); the output
 *                is every face with an endpoint in those planes, in global face order
 *   n_faces      out; call with neighbors == NULL first to query it
 *   volumes      [ (plane_hi-plane_lo+1)*ns[1]*ns[2] ] or NULL */
int fvb_device_alloc(fvb_handle h, int64_t bytes, void **dev_ptr);
int fvb_device_free(fvb_handle h, void *dev_ptr);
int fvb_device_copy(fvb_handle h, void *dst, const void *src, int64_t bytes); /* any direction */
int fvb_regulargrid(fvb_handle h, const double mins[3], const double maxs[3], const int64_t ns[3],
                    int64_t plane_lo, int64_t plane_hi, int64_t *n_faces, int64_t *neighbors,
                    double *aol, double *volumes);
/* nodehycos: values of nodes node_lo..node_lo+n_have-1 (1-based, node order; host or device);
 * out[i] = logmean ? (k1+k2)/2 : sqrt(k1*k2) for the n_faces device-resident neighbor pairs. */
int fvb_nodehycos2neighborhycos(fvb_handle h, int64_t n_faces, const int64_t *neighbors_dev,
                                const double *nodehycos, int64_t node_lo, int64_t n_have, int logmean,
                                double *out_dev);

/* ---- preconditioner ----------------------------------------------------------------------
 * kind 0: Jacobi (default; north_star).  kind 1: aggregation-multigrid V-cycle (the reference
 * preconditions with Ruge-Stueben AMG, src/FiniteVolume.jl:160):
 *   - box-structured matrices (diagonal format, 3 offsets): geometric 2x2x2 aggregation on the
 *     diagonal copy (csrc/mg.cuh), also across slab ranks;            active_kind = 1
 *   - any other matrix, single rank (fracture networks, irregular Dirichlet sets): algebraic
 *     double-pairwise aggregation on the CSR rows (csrc/amg.cuh);      active_kind = 2
 * nu = smoothing sweeps per side (>=1), omega = Jacobi damping, oc = coarse-correction scaling;
 * pass 0 for the defaults (2, 0.8, 1.5).
 * Called after fvb_assemble it fails with FVB_ERR_BAD_INPUT when no hierarchy can be built;
 * set before, fvb_assemble falls back to Jacobi silently -- check fvb_get_preconditioner.
 * The transient step (fvb_step) always uses Jacobi. */
int fvb_set_preconditioner(fvb_handle h, int kind, int nu, double omega, double oc);
int fvb_get_preconditioner(fvb_handle h, int *active_kind, int *n_levels);

/* ---- measurement hooks (CUDA events on the handle's own stream) ------------------- */
/* Average device time of one SpMV launch (K5) / one PCG iteration over `reps` launches
 * on resident data, after `warmup` untimed ones. */
int fvb_time_spmv(fvb_handle h, int warmup, int reps, double *ms_avg);
/* Device times of the phases of the last fvb_assemble / fvb_solve on this handle. */
typedef struct {
  double h2d_ms, assemble_ms, solve_ms, d2h_ms;
  double spmv_ms_total;    /* summed device time of the SpMV launches sampled in the last solve */
  int64_t spmv_samples;    /* how many launches that sum covers (0 when profiling is off)      */
  int64_t kernel_launches; /* kernels launched by the library since fvb_create */
} fvb_timings;
/* Bracket every `stride`-th SpMV launch of subsequent solves with CUDA events (at most 64
 * samples per solve); 0 switches it off.  Costs two event records per sampled launch. */
int fvb_set_profiling(fvb_handle h, int stride);
int fvb_get_timings(fvb_handle h, fvb_timings *out);
int fvb_sync(fvb_handle h);

/* ---- single-process multi-GPU front end -------------------------------------------------------
 * The reference is ONE single-threaded Julia process: a drop-in lets that one thread make one call and
 * have the solve run on all the GPUs of the box -- no torchrun, no MPI.  An fvb_multi owns one handle
 * per device and drives them from worker threads it creates per call (the caller's thread blocks, like
 * any ccall).  The whole problem is passed exactly as to solvediffusion (src/FiniteVolume.jl:157), as
 * HOST arrays; the library partitions it into contiguous node ranges (whole x-planes balanced by free
 * planes when the face list is regulargrid's -- each device then gets exactly regulargrid's slab list,
 * one contiguous block copied straight from the caller's arrays -- otherwise equal node counts with the
 * faces filtered on the host), builds the halo plan itself and maps the devices' vectors by
 * cudaDeviceEnablePeerAccess (halo planes and CG scalars travel over NVLink peer memory, as in the
 * one-process-per-GPU mode; no CUDA IPC).  ndev = 1 is the plain single-GPU path.
 *   device_ids   NULL for 0..ndev-1; a device may appear only once
 *   head_nodes   all N nodes (host); x_free all Nf free rows (host); either may be NULL
 * fvb_multi_get_csr returns the WHOLE matrix (ptr[Nf+1], idx/val[nnz], 1-based) = SparseMatrixCSC A. */
typedef struct fvb_multi_s *fvb_multi;
int fvb_multi_create(int ndev, const int *device_ids, fvb_multi *out);
int fvb_multi_destroy(fvb_multi m);
int fvb_multi_set_preconditioner(fvb_multi m, int kind, int nu, double omega, double oc);
int fvb_multi_assemble(fvb_multi m, int64_t n_nodes, int64_t n_faces, const int64_t *neighbors, const double *aol,
                       const double *cond, int64_t n_cond, const int64_t *metaindex, int logk,
                       const double *sources, int64_t n_dirichlet, const int64_t *dnodes, const double *dheads);
/* grid-implicit variant (see fvb_assemble_regulargrid): nodehycos = all N node values, sources all N or NULL */
int fvb_multi_assemble_regulargrid(fvb_multi m, const double mins[3], const double maxs[3], const int64_t ns[3],
                                   const double *nodehycos, int logmean, int logk, const double *sources,
                                   int64_t n_dirichlet, const int64_t *dnodes, const double *dheads);
/* node_lo/node_hi: ndev entries each (1-based inclusive range of every device), or NULL */
int fvb_multi_sizes(fvb_multi m, int64_t *nf_global, int64_t *nnz_global, int *ndev, int64_t *node_lo,
                    int64_t *node_hi);
int fvb_multi_solve(fvb_multi m, double rtol, int64_t maxiter, double *head_nodes, double *x_free,
                    int64_t *iters, int *converged, double *resnorm_hist, int64_t hist_cap);
int fvb_multi_get_csr(fvb_multi m, int64_t *ptr, int64_t *idx, double *val);
int fvb_multi_get_b(fvb_multi m, double *b);
int fvb_multi_get_freenode(fvb_multi m, uint8_t *freenode);
/* the per-device handle (owned by m) for anything else: timings, formats, values-only updates */
int fvb_multi_device_handle(fvb_multi m, int i, fvb_handle *out);

#ifdef __cplusplus
}
#endif
#endif /* FVB200_H */
