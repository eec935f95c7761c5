"""ctypes face of oracle/fv_oracle.c plus the host-level restatement of the
reference's drivers (solvediffusion, the backward-Euler stepper, the adjoint).

TEST INFRASTRUCTURE ONLY -- see the header of fv_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (finitevolume.jl_b200/) never imports this module.

Every function cites the reference lines it restates (paths relative to the
reference checkout).  Index conventions follow Julia: int64, 1-based.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile fv_oracle.c with the committed Makefile (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "libfv_oracle.so")
    src = os.path.join(_HERE, "fv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libfv_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libfv_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.fvo_getfreenodes.restype = C.c_int64
        L.fvo_getnodei2dirichleti.restype = C.c_int64
        L.fvo_assembleA_coo.restype = C.c_int64
        L.fvo_sparse.restype = C.c_int64
        L.fvo_regulargrid.restype = C.c_int64
        L.fvo_pcg.restype = C.c_int64
        L.fvo_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def as_pairs(neighbors) -> np.ndarray:
    """Accept [(n1,n2),...] / (F,2) array / flat 2F array -> flat int64[2F]."""
    a = np.ascontiguousarray(np.asarray(neighbors, dtype=np.int64))
    return a.reshape(-1)


def materialize_metaindex(metaindex, F):
    """The reference takes an arbitrary callable (FiniteVolume.jl:75); the C level
    takes the table metaindex(1..F), or None for the identity."""
    if metaindex is None:
        return None
    if callable(metaindex):
        return np.array([metaindex(i) for i in range(1, F + 1)], dtype=np.int64)
    return np.ascontiguousarray(metaindex, dtype=np.int64)


# --------------------------------------------------------------------------------------
def getfreenodes(n, dirichletnodes):
    """src/FiniteVolume.jl:32-44."""
    dn = np.ascontiguousarray(dirichletnodes, dtype=np.int64)
    freenode = np.empty(n, np.uint8)
    n2f = np.empty(n, np.int64)
    lib().fvo_getfreenodes(C.c_int64(n), _p(dn), C.c_int64(dn.size), _p(freenode), _p(n2f))
    return freenode.astype(bool), n2f


def getnodei2dirichleti(sources, dirichletnodes):
    """src/FiniteVolume.jl:20-30 (raises like the reference's error())."""
    src = np.ascontiguousarray(sources, dtype=np.float64)
    dn = np.ascontiguousarray(dirichletnodes, dtype=np.int64)
    out = np.empty(src.size, np.int64)
    bad = lib().fvo_getnodei2dirichleti(C.c_int64(src.size), _p(src), _p(dn), C.c_int64(dn.size), _p(out))
    if bad:
        raise ValueError(f"There cannot be a source at a Dirichlet node, but node {bad} is a "
                         "Dirichlet node where a source is located.")
    return out


@dataclass
class CSC:
    """A SparseMatrixCSC{Float64,Int64} look-alike: 1-based colptr/rowval + nzval."""
    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    def toscipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))


def sparse(I, J, V, m, n) -> CSC:
    """SparseArrays.sparse(I,J,V,m,n,+) as used at src/FiniteVolume.jl:107."""
    I = np.ascontiguousarray(I, np.int64); J = np.ascontiguousarray(J, np.int64)
    V = np.ascontiguousarray(V, np.float64)
    colptr = np.empty(n + 1, np.int64)
    rowval = np.empty(max(I.size, 1), np.int64)
    nzval = np.empty(max(I.size, 1), np.float64)
    nnz = lib().fvo_sparse(C.c_int64(m), C.c_int64(n), C.c_int64(I.size), _p(I), _p(J), _p(V),
                           _p(colptr), _p(rowval), _p(nzval))
    return CSC(m, n, colptr, rowval[:nnz].copy(), nzval[:nnz].copy())


def assembleA(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
              metaindex=None, logtransformconductivity=False) -> CSC:
    """src/FiniteVolume.jl:75-108."""
    nb = as_pairs(neighbors)
    F = nb.size // 2
    aol = np.ascontiguousarray(areasoverlengths, np.float64)
    cond = np.ascontiguousarray(conductivities, np.float64)
    meta = materialize_metaindex(metaindex, F)
    freenode, n2f = getfreenodes(len(sources), dirichletnodes)
    fn8 = freenode.astype(np.uint8)
    I = np.empty(4 * F + 1, np.int64); J = np.empty(4 * F + 1, np.int64); V = np.empty(4 * F + 1, np.float64)
    m = lib().fvo_assembleA_coo(C.c_int64(F), _p(nb), _p(aol), _p(cond), _p(meta),
                                C.c_int(int(logtransformconductivity)), _p(fn8), _p(n2f), _p(I), _p(J), _p(V))
    nf = int(freenode.sum())
    return sparse(I[:m], J[:m], V[:m], nf, nf)


def assembleb(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
              metaindex=None, logtransformconductivity=False) -> np.ndarray:
    """src/FiniteVolume.jl:110-139."""
    nb = as_pairs(neighbors)
    F = nb.size // 2
    aol = np.ascontiguousarray(areasoverlengths, np.float64)
    cond = np.ascontiguousarray(conductivities, np.float64)
    src = np.ascontiguousarray(sources, np.float64)
    dh = np.ascontiguousarray(dirichletheads, np.float64)
    meta = materialize_metaindex(metaindex, F)
    n2d = getnodei2dirichleti(src, dirichletnodes)
    freenode, n2f = getfreenodes(src.size, dirichletnodes)
    fn8 = freenode.astype(np.uint8)
    b = np.empty(int(freenode.sum()), np.float64)
    lib().fvo_assembleb(C.c_int64(src.size), C.c_int64(F), _p(nb), _p(aol), _p(cond), _p(meta),
                        C.c_int(int(logtransformconductivity)), _p(src), _p(dh), _p(fn8), _p(n2f), _p(n2d), _p(b))
    return b


def freenodes2nodes(result, sources, dirichletnodes, dirichletheads):
    """src/FiniteVolume.jl:141-155."""
    src = np.ascontiguousarray(sources, np.float64)
    n2d = getnodei2dirichleti(src, dirichletnodes)
    freenode, n2f = getfreenodes(src.size, dirichletnodes)
    head = np.empty(src.size, np.float64)
    res = np.ascontiguousarray(result, np.float64)
    dh = np.ascontiguousarray(dirichletheads, np.float64)
    lib().fvo_freenodes2nodes(C.c_int64(src.size), _p(res), _p(dh), _p(freenode.astype(np.uint8)), _p(n2d), _p(head))
    return head, freenode, n2f


@dataclass
class ConvergenceHistory:
    """The fields of IterativeSolvers.ConvergenceHistory that callers of the
    reference use (examples/box_model/ex_piml_data.jl:46, examples/waffle/ex.jl:25)."""
    isconverged: bool
    iters: int
    data: dict = field(default_factory=dict)


def cg(A: CSC, b, x0=None, Pl="jacobi", tol=np.sqrt(np.finfo(np.float64).eps), maxiter=None,
       threaded=False):
    """IterativeSolvers.cg / cg! 0.8.1 (call sites src/FiniteVolume.jl:161,
    src/transient.jl:52,55) with Pl = Jacobi or identity."""
    n = A.n
    b = np.ascontiguousarray(b, np.float64)
    maxiter = n if maxiter is None else int(maxiter)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    hist = np.empty(max(maxiter, 1), np.float64)
    conv = C.c_int(0)
    it = lib().fvo_pcg(C.c_int64(n), _p(A.colptr), _p(A.rowval), _p(A.nzval), _p(b), _p(x),
                       C.c_int(0 if x0 is None else 1), C.c_int(1 if Pl == "jacobi" else 0),
                       C.c_double(tol), C.c_int64(maxiter), C.c_int(1 if threaded else 0),
                       _p(hist), C.c_int64(hist.size), C.byref(conv))
    return x, ConvergenceHistory(bool(conv.value), int(it), {"resnorm": hist[:it].copy()})


def spmv(A: CSC, x):
    y = np.empty(A.m)
    x = np.ascontiguousarray(x, np.float64)
    lib().fvo_spmv_csc(C.c_int64(A.n), _p(A.colptr), _p(A.rowval), _p(A.nzval), _p(x), _p(y))
    return y


def solvediffusion(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                   maxiter=400, tol=np.sqrt(np.finfo(np.float64).eps), metaindex=None,
                   logtransformconductivity=False, threaded=False):
    """src/FiniteVolume.jl:157-165 with the AMG preconditioner replaced by Jacobi
    (north_star); maxiter therefore counts Jacobi-PCG iterations."""
    A = assembleA(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                  metaindex, logtransformconductivity)
    b = assembleb(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                  metaindex, logtransformconductivity)
    result, ch = cg(A, b, Pl="jacobi", tol=tol, maxiter=maxiter, threaded=threaded)
    head, freenode, _ = freenodes2nodes(result, sources, dirichletnodes, dirichletheads)
    return head, ch, A, b, freenode


def regulargrid(mins, maxs, ns, want_coords=True):
    """src/grid.jl:56-110."""
    ns = np.asarray(ns, np.int64)
    if len(mins) != 3:
        raise ValueError("only 3 dimensions supported")
    N = int(np.prod(ns))
    F = int(3 * N - ns[0] * ns[1] - ns[0] * ns[2] - ns[1] * ns[2])
    coords = np.empty((N, 3), np.float64) if want_coords else None
    nb = np.empty(2 * F, np.int64); aol = np.empty(F, np.float64); vol = np.empty(N, np.float64)
    mins_ = np.ascontiguousarray(mins, np.float64); maxs_ = np.ascontiguousarray(maxs, np.float64)
    got = lib().fvo_regulargrid(_p(mins_), _p(maxs_), _p(ns), _p(coords), _p(nb), _p(aol), _p(vol))
    assert got == F
    return (coords.T if want_coords else None), nb.reshape(F, 2), aol, vol


def nodehycos2neighborhycos(neighbors, nodehycos, logtransformhyco=False):
    """src/grid.jl:14-33; nodehycos shaped (n3,n2,n1) column-major == node order."""
    nb = as_pairs(neighbors)
    flat = np.ascontiguousarray(np.asarray(nodehycos, np.float64).reshape(-1, order="F"))
    out = np.empty(nb.size // 2, np.float64)
    lib().fvo_nodehycos2neighborhycos(C.c_int64(out.size), _p(nb), _p(flat), C.c_int(int(logtransformhyco)), _p(out))
    return out


# --------------------------------------------------------------------------------------
# Transient driver: src/transient.jl.  The step controller is restated line for line;
# the linear solves default to a direct sparse solve (scipy) so that what is pinned by
# the Theis / one-node tests is the controller + assembly, independent of any
# iterative-solver tolerance.  (Signature deviation: linearsolver(A, dt, rhs, x0) solves
# (A + I/dt) x = rhs, so a factorisation can be cached per dt; the reference's hook gets
# the already shifted matrix, src/transient.jl:72-73.)
# --------------------------------------------------------------------------------------
class DirectSolver:
    """linearsolver(A, dt, rhs, x0): exact solve of (A + I/dt) x = rhs, with the sparse LU
    cached per dt (the step-doubling ladder revisits the same few dt values thousands of times)."""

    def __init__(self):
        self.cache = {}
        self.solves = 0

    def __call__(self, A, dt, rhs, x0):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        self.solves += 1
        if A.shape[0] <= 8:
            M = (A.toarray() if sp.issparse(A) else np.asarray(A)) + np.eye(A.shape[0]) / dt
            return np.linalg.solve(M, rhs)
        key = (id(A), dt)
        if key not in self.cache:
            if len(self.cache) > 64:
                self.cache.clear()
            self.cache[key] = spla.splu((A + sp.identity(A.shape[0], format="csc") * (1 / dt)).tocsc())
        return self.cache[key].solve(rhs)


def backwardeuleronestep(A, b, u_k, dt, linearsolver):
    """src/transient.jl:65-76: (A + I/dt) u = b + u_k/dt."""
    if dt <= 0:
        raise ValueError("time step must be positive")
    rhs = b + u_k / dt
    return linearsolver(A, dt, rhs, u_k)


def backwardeulertwostep(A, getb, u_k, t, dt, linearsolver, atol, onestep=None):
    """src/transient.jl:78-87."""
    if onestep is None:
        onestep = backwardeuleronestep(A, getb(t), u_k, dt, linearsolver)
    twostep1 = backwardeuleronestep(A, getb(t), u_k, 0.5 * dt, linearsolver)
    twostep = backwardeuleronestep(A, getb(t + 0.5 * dt), twostep1, 0.5 * dt, linearsolver)
    err = np.linalg.norm(onestep - twostep)
    if err < atol:
        return twostep, dt, err < atol / 4
    return twostep1, 0.5 * dt, False


def adaptivebackwardeulerstep(A, getb, u_k, t, dt, linearsolver, atol, callback):
    """src/transient.jl:89-121."""
    callback(t, dt)
    u_new, last, inc = backwardeulertwostep(A, getb, u_k, t, dt, linearsolver, atol)
    if last < dt:
        failed = True
        elapsed = 0.0
        u_el = u_k
        target = last
        while elapsed < dt:
            callback(t, dt)
            if failed:
                u_new, last, inc = backwardeulertwostep(A, getb, u_el, t + elapsed, target, linearsolver, atol, u_new)
            else:
                u_new, last, inc = backwardeulertwostep(A, getb, u_el, t + elapsed, target, linearsolver, atol)
            if last == target:
                elapsed += last
                u_el = u_new
                if inc:
                    target = 2 * last
                failed = False
            elif last < target:
                target = last
                failed = True
            else:
                raise RuntimeError("Code is broken -- laststeptime should never be greater than targetdt")
            target = min(target, dt - elapsed)
    return u_new, last, inc


def fixedbackwardeulerstep(A, getb, u_k, t, dt, linearsolver, atol, callback):
    """src/transient.jl:130-134."""
    callback(t, dt)
    return backwardeuleronestep(A, getb(t), u_k, dt, linearsolver), dt, False


def backwardeulerintegrate_core(u0, A, getb, dt0, t0, tfinal, stepper=adaptivebackwardeulerstep,
                                linearsolver=None, atol=1e-4, callback=lambda t, dt: None):
    """src/transient.jl:136-154 (A: scipy sparse, already volume-scaled)."""
    if not callable(getb):
        bconst = getb
        getb = lambda t: bconst  # noqa: E731  (src/transient.jl:123-128)
    if linearsolver is None:
        linearsolver = DirectSolver()
    us = [np.array(u0, dtype=np.float64)]
    ts = [t0]
    dt = min(dt0, tfinal - t0)
    while ts[-1] < tfinal:
        sol, last, inc = stepper(A, getb, us[-1], ts[-1], dt, linearsolver, atol, callback)
        us.append(sol)
        ts.append(ts[-1] + dt)
        dt = min(tfinal - ts[-1], 2 * last) if inc else min(tfinal - ts[-1], last)
    return us, ts


def backwardeulerintegrate(u0, tspan, Ss, volumes, neighbors, areasoverlengths, conductivities, sources,
                           dirichletnodes, dirichletheads, metaindex=None, logtransformconductivity=False,
                           dt0=1.0, getb=None, **kw):
    """src/transient.jl:156-174: assemble, scale rows by 1/(Ss*vol), integrate, scatter."""
    import scipy.sparse as sp
    u0 = np.asarray(u0, np.float64)
    freenode, n2f = getfreenodes(u0.size, dirichletnodes)
    D = (Ss * np.asarray(volumes, np.float64))[freenode]
    if getb is None:
        b = assembleb(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                      metaindex, logtransformconductivity) / D
        getb = lambda t: b  # noqa: E731
    A = assembleA(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                  metaindex, logtransformconductivity).toscipy().tocsr()
    A = sp.diags(1.0 / D) @ A
    us, ts = backwardeulerintegrate_core(u0[freenode], A, getb, dt0, tspan[0], tspan[1], **kw)
    us = [freenodes2nodes(x, sources, dirichletnodes, dirichletheads)[0] for x in us]
    return us, ts


def adjointintegrate(getdgdu, tspan, Ss, volumes, neighbors, areasoverlengths, conductivities, sources,
                     dirichletnodes, dirichletheads, metaindex=None, logtransformconductivity=False,
                     dt0=1.0, **kw):
    """src/transient.jl:188-205: gamma' = (D^-1 A)^T gamma + dg/du(T-t), gamma(0)=0,
    returned reversed as lambda(t) = gamma(T-t)."""
    import scipy.sparse as sp
    volumes = np.asarray(volumes, np.float64)
    freenode, n2f = getfreenodes(volumes.size, dirichletnodes)
    D = (Ss * volumes)[freenode]
    A = assembleA(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                  metaindex, logtransformconductivity).toscipy().tocsr()
    At = (sp.diags(1.0 / D) @ A).T.tocsr()
    gam, tsg = backwardeulerintegrate_core(np.zeros(At.shape[1]), At, lambda t: getdgdu(tspan[1] - t), dt0,
                                           tspan[0], tspan[1], **kw)
    return gam[::-1], [tspan[1] - t for t in tsg][::-1]


def num_threads() -> int:
    return int(lib().fvo_num_threads())


def set_num_threads(n: int) -> None:
    lib().fvo_set_num_threads(C.c_int(int(n)))
