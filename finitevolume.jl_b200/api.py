"""Host-side mirror of the reference's call surface for the assemble -> solve path.

Function names, argument order and meaning, return tuples and error behaviour follow
madsjulia/FiniteVolume.jl (file:line cites are relative to the reference checkout), so the
parity tests read like the reference's own tests.  Every function here drives
libfvb200.so through ctypes; there is no CPU implementation behind any of them.

Index conventions are Julia's: node ids, Dirichlet nodes, metaindex values and the
returned colptr/rowval are 1-based int64.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import check, f64, i64, lib, ptr

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)  # IterativeSolvers' default relative tolerance
DEFAULT_MAXITER = 100_000


def _pairs(neighbors) -> np.ndarray:
    """[(n1,n2),...] / (F,2) / flat 2F -> flat int64[2F] (the memory image of
    Array{Pair{Int,Int},1})."""
    return i64(np.asarray(neighbors, dtype=np.int64)).reshape(-1)


def _metaindex_table(metaindex, F):
    """The reference accepts any callable i -> index (src/FiniteVolume.jl:75); the C ABI
    takes its table.  None means the identity."""
    if metaindex is None:
        return None
    if callable(metaindex):
        return np.fromiter((metaindex(i) for i in range(1, F + 1)), dtype=np.int64, count=F)
    return i64(metaindex)


@dataclass
class ConvergenceHistory:
    """What callers of the reference read from IterativeSolvers.ConvergenceHistory
    (examples/box_model/ex_piml_data.jl:46, examples/waffle/ex.jl:25)."""
    isconverged: bool
    iters: int
    data: dict = field(default_factory=dict)


class DeviceArray:
    """A plain device allocation owned through the library (fvb_device_alloc/free): what the
    device-side grid helpers return and what System.assemble accepts in place of numpy arrays."""

    def __init__(self, system, shape, dtype):
        self.system, self.shape, self.dtype = system, tuple(int(v) for v in np.atleast_1d(shape)), np.dtype(dtype)
        self.size = int(np.prod(self.shape))
        self.nbytes = self.size * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().fvb_device_alloc(system._h, C.c_int64(self.nbytes), C.byref(p)))
        self.ptr = p.value or 0

    def to_host(self):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            check(lib().fvb_device_copy(self.system._h, ptr(out), C.c_void_p(self.ptr), C.c_int64(self.nbytes)))
        return out

    def free(self):
        if self.ptr and self.system._h:
            lib().fvb_device_free(self.system._h, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _arg(a, dtype):
    """-> (void*, element count, keep-alive) for a numpy-like input or a DeviceArray."""
    if isinstance(a, DeviceArray):
        if a.dtype != np.dtype(dtype):
            raise TypeError(f"device array has dtype {a.dtype}, expected {np.dtype(dtype)}")
        return C.c_void_p(a.ptr), a.size, a
    arr = np.ascontiguousarray(np.asarray(a, dtype=dtype)).reshape(-1)
    return ptr(arr), arr.size, arr


class System:
    """One assembled problem resident on one GPU (an fvb_handle)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fvb_create(C.c_int(device), C.byref(self._h)))
        self.device = device
        self.node_lo = 1
        self.node_hi = 0
        self.nranks = 1
        self.rank = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().fvb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU ---------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.UNIQUE_ID_BYTES)()
        check(lib().fvb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * _lib.UNIQUE_ID_BYTES).from_buffer_copy(uid)
        check(lib().fvb_comm_init(self._h, C.c_int(nranks), C.c_int(rank), buf))
        self.nranks, self.rank = nranks, rank

    # ---- assembly ----------------------------------------------------------------------
    def assemble(self, neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                 metaindex=None, logtransformconductivity=False, n_nodes=None, node_range=None):
        """assembleA + assembleb (src/FiniteVolume.jl:75-139).  With node_range=(lo,hi)
        (1-based inclusive) `sources` is the owned slice and `neighbors` etc. are the faces
        touching owned nodes, in global face order."""
        nb_p, nb_n, k0 = _arg(neighbors, np.int64)
        F = nb_n // 2
        aol_p, aol_n, k1 = _arg(areasoverlengths, np.float64)
        cond_p, cond_n, k2 = _arg(conductivities, np.float64)
        src_p, src_n, k3 = _arg(sources, np.float64)
        dn = i64(dirichletnodes)
        dh = f64(dirichletheads)
        if aol_n != F:
            raise ValueError("areasoverlengths and neighbors differ in length")
        if dn.size != dh.size:
            raise ValueError("dirichletnodes and dirichletheads differ in length")
        meta = _metaindex_table(metaindex, F)
        if meta is None and cond_n < F:
            raise IndexError("conductivities is shorter than neighbors")  # Julia: BoundsError
        N = int(src_n if n_nodes is None else n_nodes)
        lo, hi = (1, N) if node_range is None else (int(node_range[0]), int(node_range[1]))
        if src_n != hi - lo + 1:
            raise ValueError("sources must cover exactly the owned node range")
        check(lib().fvb_assemble(self._h, C.c_int64(N), C.c_int64(lo), C.c_int64(hi), C.c_int64(F), nb_p, aol_p,
                                 cond_p, C.c_int64(cond_n), ptr(meta), C.c_int(int(bool(logtransformconductivity))),
                                 src_p, C.c_int64(dn.size), ptr(dn), ptr(dh)))
        self.node_lo, self.node_hi = lo, hi
        self._logk = bool(logtransformconductivity)
        return self

    def assemble_raw(self, n_nodes, node_lo, node_hi, n_faces, neighbors_ptr, aol_ptr, cond_ptr, n_cond, meta_ptr,
                     logk, sources_ptr, n_dirichlet, dnodes_ptr, dheads_ptr):
        """fvb_assemble on raw addresses (host, pinned or device -- e.g. torch tensors' data_ptr());
        nothing is copied or converted on the Python side."""
        vp = lambda a: C.c_void_p(a) if a else None  # noqa: E731
        check(lib().fvb_assemble(self._h, C.c_int64(n_nodes), C.c_int64(node_lo), C.c_int64(node_hi),
                                 C.c_int64(n_faces), vp(neighbors_ptr), vp(aol_ptr), vp(cond_ptr), C.c_int64(n_cond),
                                 vp(meta_ptr), C.c_int(int(bool(logk))), vp(sources_ptr), C.c_int64(n_dirichlet),
                                 vp(dnodes_ptr), vp(dheads_ptr)))
        self.node_lo, self.node_hi = int(node_lo), int(node_hi)
        self._logk = bool(logk)
        return self

    def solve_raw(self, rtol, maxiter, head_ptr=0, x_ptr=0, x0_ptr=0):
        """fvb_solve on raw addresses; returns (iters, converged)."""
        vp = lambda a: C.c_void_p(a) if a else None  # noqa: E731
        iters = C.c_int64()
        conv = C.c_int()
        check(lib().fvb_solve(self._h, C.c_double(rtol), C.c_int64(int(maxiter)), vp(x0_ptr), vp(head_ptr), vp(x_ptr),
                              C.byref(iters), C.byref(conv), None, C.c_int64(0)))
        return int(iters.value), bool(conv.value)

    # ---- device-side grid helpers (src/grid.jl) ---------------------------------------------
    def device_regulargrid(self, mins, maxs, ns, planes=None, want_volumes=True):
        """regulargrid (src/grid.jl:56-110) generated on the GPU -> (neighbors, areasoverlengths, volumes)
        as DeviceArrays, bit-identical to the host builder; planes=(lo,hi) restricts to a slab."""
        if len(mins) != 3 or len(maxs) != 3 or len(ns) != 3:
            raise ValueError("only 3 dimensions supported")
        mn, mx, nn = f64(mins), f64(maxs), i64(ns)
        lo, hi = (1, int(nn[0])) if planes is None else (int(planes[0]), int(planes[1]))
        F = C.c_int64()
        check(lib().fvb_regulargrid(self._h, ptr(mn), ptr(mx), ptr(nn), C.c_int64(lo), C.c_int64(hi), C.byref(F), None,
                                    None, None))
        nb = DeviceArray(self, (F.value, 2), np.int64)
        aol = DeviceArray(self, (F.value,), np.float64)
        vol = DeviceArray(self, ((hi - lo + 1) * int(nn[1]) * int(nn[2]),), np.float64) if want_volumes else None
        check(lib().fvb_regulargrid(self._h, ptr(mn), ptr(mx), ptr(nn), C.c_int64(lo), C.c_int64(hi), C.byref(F),
                                    C.c_void_p(nb.ptr), C.c_void_p(aol.ptr), C.c_void_p(vol.ptr) if vol else None))
        return nb, aol, vol

    def device_nodehycos2neighborhycos(self, neighbors_dev, nodehycos, logtransformhyco=False, node_lo=1):
        """nodehycos2neighborhycos (src/grid.jl:14-33) on the GPU; nodehycos covers nodes node_lo.. in node order."""
        k = np.asarray(nodehycos, np.float64)
        flat = np.ascontiguousarray(k.reshape(-1, order="F") if k.ndim == 3 else k.reshape(-1))
        out = DeviceArray(self, (neighbors_dev.shape[0],), np.float64)
        check(lib().fvb_nodehycos2neighborhycos(self._h, C.c_int64(neighbors_dev.shape[0]), C.c_void_p(neighbors_dev.ptr),
                                                ptr(flat), C.c_int64(int(node_lo)), C.c_int64(flat.size),
                                                C.c_int(int(bool(logtransformhyco))), C.c_void_p(out.ptr)))
        return out

    def assemble_regulargrid(self, mins, maxs, ns, nodehycos, sources, dirichletnodes, dirichletheads,
                             logmean=True, logtransformconductivity=True, planes=None):
        """Grid-implicit assembly (fvb_assemble_regulargrid): regulargrid + nodehycos2neighborhycos + assembleA/b
        without any per-face array.  nodehycos: node values (node order) of planes max(1,lo-1)..min(n1,hi+1)
        (numpy, a DeviceArray or an int device address); sources: owned nodes or None (zeros)."""
        mn, mx, nn = f64(mins), f64(maxs), i64(ns)
        lo, hi = (1, int(nn[0])) if planes is None else (int(planes[0]), int(planes[1]))
        if isinstance(nodehycos, int):
            k_p, keep_k = C.c_void_p(nodehycos), None
        else:
            k_p, _, keep_k = _arg(nodehycos, np.float64)
        src_p, keep_s = (None, None) if sources is None else _arg(sources, np.float64)[::2]
        dn, dh = i64(dirichletnodes), f64(dirichletheads)
        check(lib().fvb_assemble_regulargrid(self._h, ptr(mn), ptr(mx), ptr(nn), C.c_int64(lo), C.c_int64(hi), k_p,
                                             C.c_int(int(bool(logmean))), C.c_int(int(bool(logtransformconductivity))),
                                             src_p, C.c_int64(dn.size), ptr(dn), ptr(dh)))
        plane = int(nn[1]) * int(nn[2])
        self.node_lo, self.node_hi = (lo - 1) * plane + 1, hi * plane
        self._logk = bool(logtransformconductivity)
        return self

    def set_assembly(self, mode):
        """0 = automatic (closed-form rows for regulargrid-ordered face lists), 1 = always the general path."""
        check(lib().fvb_set_assembly(self._h, C.c_int(int(mode))))

    def assembly(self):
        """-> "general" | "box" (closed form from the caller's arrays) | "implicit" (no face arrays)."""
        a = C.c_int()
        check(lib().fvb_get_assembly(self._h, C.byref(a)))
        return {0: "general", 1: "box", 2: "implicit"}[a.value]

    def update_values(self, conductivities, sources=None, dirichletheads=None, logtransformconductivity=None):
        cond = f64(conductivities)
        logk = self._logk if logtransformconductivity is None else bool(logtransformconductivity)
        src = None if sources is None else f64(sources)
        dh = None if dirichletheads is None else f64(dirichletheads)
        check(lib().fvb_update_values(self._h, ptr(cond), C.c_int64(cond.size), C.c_int(int(logk)), ptr(src), ptr(dh)))
        return self

    def sizes(self):
        v = [C.c_int64() for _ in range(5)]
        check(lib().fvb_sizes(self._h, *[C.byref(x) for x in v]))
        return dict(nf_local=v[0].value, nnz_local=v[1].value, row_start=v[2].value, nf_global=v[3].value,
                    n_halo=v[4].value)

    @property
    def n_own_nodes(self):
        return self.node_hi - self.node_lo + 1

    def csr(self):
        s = self.sizes()
        p = np.empty(s["nf_local"] + 1, np.int64)
        idx = np.empty(s["nnz_local"], np.int64)
        val = np.empty(s["nnz_local"], np.float64)
        check(lib().fvb_get_csr(self._h, ptr(p), ptr(idx), ptr(val)))
        return p, idx, val

    def b(self):
        out = np.empty(self.sizes()["nf_local"], np.float64)
        check(lib().fvb_get_b(self._h, ptr(out)))
        return out

    def diag(self):
        out = np.empty(self.sizes()["nf_local"], np.float64)
        check(lib().fvb_get_diag(self._h, ptr(out)))
        return out

    def freenode(self):
        out = np.empty(self.n_own_nodes, np.uint8)
        check(lib().fvb_get_freenode(self._h, ptr(out)))
        return out.astype(bool)

    def nodei2freenodei(self):
        out = np.empty(self.n_own_nodes, np.int64)
        check(lib().fvb_get_nodei2freenodei(self._h, ptr(out)))
        return out

    def halo_cols(self):
        out = np.empty(self.sizes()["n_halo"], np.int64)
        check(lib().fvb_get_halo_cols(self._h, ptr(out)))
        return out

    def set_halo_plan(self, peer_ranks, send_counts, send_rows, recv_counts):
        pr = np.ascontiguousarray(peer_ranks, np.int32)
        sc = i64(send_counts)
        sr = np.ascontiguousarray(send_rows, np.int32)
        rc = i64(recv_counts)
        check(lib().fvb_set_halo_plan(self._h, C.c_int(pr.size), ptr(pr), ptr(sc), ptr(sr), ptr(rc)))

    def peer_export(self) -> bytes:
        buf = (C.c_uint8 * _lib.PEER_BLOB_BYTES)()
        check(lib().fvb_peer_export(self._h, buf))
        return bytes(buf)

    def peer_import(self, blobs_by_rank, send_dst_index):
        raw = b"".join(blobs_by_rank)
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        dst = i64(send_dst_index)
        check(lib().fvb_peer_import(self._h, buf, ptr(dst)))

    # ---- solve ---------------------------------------------------------------------------
    def solve(self, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER, x0=None, want_head=True, want_x=False, hist_cap=None,
              head_out=None):
        s = self.sizes()
        x0a = None if x0 is None else f64(x0)
        head = None
        if want_head:
            head = head_out if head_out is not None else np.empty(self.n_own_nodes, np.float64)
        x = np.empty(s["nf_local"], np.float64) if want_x else None
        cap = int(min(maxiter, 1 << 24) if hist_cap is None else hist_cap)
        hist = np.empty(max(cap, 1), np.float64)
        iters = C.c_int64()
        conv = C.c_int()
        check(lib().fvb_solve(self._h, C.c_double(rtol), C.c_int64(int(maxiter)), ptr(x0a), ptr(head), ptr(x),
                              C.byref(iters), C.byref(conv), ptr(hist), C.c_int64(cap)))
        ch = ConvergenceHistory(bool(conv.value), int(iters.value), {"resnorm": hist[:min(iters.value, cap)].copy()})
        return head, x, ch

    def spmv(self, x, alpha=1.0, beta=0.0, y=None):
        xa = f64(x)
        ya = np.zeros_like(xa) if y is None else f64(y).copy()
        check(lib().fvb_spmv(self._h, C.c_double(alpha), ptr(xa), C.c_double(beta), ptr(ya)))
        return ya

    # ---- transient, device-resident vectors --------------------------------------------------
    def vec_upload(self, slot, v):
        check(lib().fvb_vec_upload(self._h, C.c_int(slot), ptr(f64(v))))

    def vec_download(self, slot):
        out = np.empty(self.sizes()["nf_local"], np.float64)
        check(lib().fvb_vec_download(self._h, C.c_int(slot), ptr(out)))
        return out

    def vec_copy(self, dst, src):
        check(lib().fvb_vec_copy(self._h, C.c_int(dst), C.c_int(src)))

    def vec_load_b(self, slot):
        check(lib().fvb_vec_load_b(self._h, C.c_int(slot)))

    def vec_diffnorm(self, a, b):
        out = C.c_double()
        check(lib().fvb_vec_diffnorm(self._h, C.c_int(a), C.c_int(b), C.byref(out)))
        return out.value

    def set_storage(self, Ss, volumes_owned):
        v = None if volumes_owned is None else f64(volumes_owned)
        check(lib().fvb_set_storage(self._h, C.c_double(Ss), ptr(v)))

    def step(self, rhs_slot, u_slot, dt, out_slot, adjoint=False, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER):
        iters = C.c_int64()
        conv = C.c_int()
        check(lib().fvb_step(self._h, C.c_int(rhs_slot), C.c_int(u_slot), C.c_double(dt), C.c_int(out_slot),
                             C.c_int(int(adjoint)), C.c_double(rtol), C.c_int64(int(maxiter)), C.byref(iters),
                             C.byref(conv)))
        return int(iters.value), bool(conv.value)

    def integrate(self, u0_free, t0, tfinal, dt0=1.0, atol=1e-4, fixed_step=False, adjoint=False, rtol=SQRT_EPS,
                  maxiter=DEFAULT_MAXITER, getb=None, callback=None, want="free", max_states=4096):
        """fvb_integrate: the reference's whole backwardeulerintegrate loop (src/transient.jl:78-154) behind one call.
        getb(t) -> UNSCALED right-hand side on the free rows (None: the assembled b); callback(t, dt) per attempted
        step.  want: "free" (states on free rows), "heads" (through freenodes2nodes) or None (times only).
        -> (states [n_states, ...] or None, ts, stats)."""
        nf = self.sizes()["nf_local"]
        u0 = f64(u0_free).reshape(-1)
        if u0.size != nf:
            raise ValueError("u0 must hold the free rows")
        keep = []
        opt = _lib.IntegrateOptions(atol=float(atol), dt0=float(dt0), fixed_step=int(bool(fixed_step)), adjoint=int(bool(adjoint)),
                                    rtol=float(rtol), maxiter=int(maxiter))
        if getb is not None:
            def _getb(t, out, _ctx):
                np.ctypeslib.as_array(out, shape=(max(nf, 1),))[:nf] = np.asarray(getb(t), np.float64).reshape(-1)
            opt.getb = _lib.GETB_FN(_getb)
            keep.append(opt.getb)
        if callback is not None:
            opt.callback = _lib.STEP_CALLBACK_FN(lambda t, dt, _ctx: callback(t, dt))
            keep.append(opt.callback)
        while True:
            ts = np.empty(max_states, np.float64)
            width = nf if want == "free" else (self.n_own_nodes if want == "heads" else 0)
            out = np.empty((max_states, width), np.float64) if width else None
            ns, nsol, nit, natt = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
            status = lib().fvb_integrate(self._h, ptr(u0), C.c_double(t0), C.c_double(tfinal), C.byref(opt), C.c_int64(max_states),
                                         ptr(ts), ptr(out) if want == "free" else None, ptr(out) if want == "heads" else None,
                                         C.byref(ns), C.byref(nsol), C.byref(nit), C.byref(natt))
            if status == 5 and b"max_states" in lib().fvb_last_error() and callback is None:
                max_states *= 4  # more accepted steps than room: rerun with larger buffers
                continue
            check(status)
            break
        k = ns.value
        stats = dict(steps=k - 1, linear_solves=nsol.value, cg_iterations=nit.value, attempts=natt.value)
        return (out[:k] if out is not None else None), ts[:k], stats

    def solve_shifted(self, rhs_slot, x0_slot, sigma, out_slot, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER):
        """(A + sigma D) x = rhs from x0, slot -> slot: the reference's linearsolver(A, rhs, x0) hook
        (src/transient.jl:136) on the resident matrix."""
        iters, conv = C.c_int64(), C.c_int()
        check(lib().fvb_solve_shifted(self._h, C.c_int(rhs_slot), C.c_int(x0_slot), C.c_double(sigma), C.c_int(out_slot),
                                      C.c_double(rtol), C.c_int64(int(maxiter)), C.byref(iters), C.byref(conv)))
        return int(iters.value), bool(conv.value)

    # ---- adjoint gradient gather -------------------------------------------------------------------
    def gradient_begin(self, neighbors):
        nb_p, _, keep = _arg(neighbors, np.int64)
        check(lib().fvb_gradient_begin(self._h, nb_p))

    def gradient_accumulate(self, u_slot, lambda_slot, weight):
        check(lib().fvb_gradient_accumulate(self._h, C.c_int(u_slot), C.c_int(lambda_slot), C.c_double(weight)))

    def gradient_end(self, n_faces):
        """-> (per-face d/dk contributions, per-face d/dh contributions, 1-based Dirichlet slot per face or -1,
        per-row d/dsources contributions)."""
        gk = np.empty(n_faces, np.float64)
        gh = np.empty(n_faces, np.float64)
        sl = np.empty(n_faces, np.int64)
        gs = np.empty(self.sizes()["nf_local"], np.float64)
        check(lib().fvb_gradient_end(self._h, ptr(gk), ptr(gh), ptr(sl), ptr(gs)))
        return gk, gh, sl, gs

    def vec_to_nodes(self, slot):
        out = np.empty(self.n_own_nodes, np.float64)
        check(lib().fvb_vec_to_nodes(self._h, C.c_int(slot), ptr(out)))
        return out

    # ---- measurement ------------------------------------------------------------------------
    def time_spmv(self, warmup=3, reps=20):
        out = C.c_double()
        check(lib().fvb_time_spmv(self._h, C.c_int(warmup), C.c_int(reps), C.byref(out)))
        return out.value

    def set_preconditioner(self, kind="jacobi", nu=0, omega=0.0, oc=0.0):
        """"jacobi" (north_star default) or "mg" (aggregation multigrid V-cycle; box-structured
        matrices only -- raises FVBError otherwise when called after assemble)."""
        k = {"jacobi": 0, "mg": 1}[kind]
        check(lib().fvb_set_preconditioner(self._h, C.c_int(k), C.c_int(int(nu)), C.c_double(omega), C.c_double(oc)))

    def preconditioner(self):
        a, l = C.c_int(), C.c_int()
        check(lib().fvb_get_preconditioner(self._h, C.byref(a), C.byref(l)))
        return {0: "jacobi", 1: "mg", 2: "amg"}[a.value], l.value

    def set_spmv_format(self, fmt):
        """0 = automatic (diagonal copy when the pattern allows), 1 = always CSR, 2 = diagonal copy with the
        per-thread-load kernel only, 3 = diagonal copy through the TMA pipeline whatever the size."""
        check(lib().fvb_set_spmv_format(self._h, C.c_int(int(fmt))))

    def spmv_format(self):
        """-> ("csr" | "dia", number of positive offsets)."""
        a, k = C.c_int(), C.c_int()
        check(lib().fvb_get_spmv_format(self._h, C.byref(a), C.byref(k)))
        return ("dia" if a.value in (2, 3) else "csr"), k.value

    def spmv_kernel(self):
        """-> "csr" | "dia" (per-thread loads) | "dia_tma" (TMA pipeline): the kernel the next product uses."""
        a, k = C.c_int(), C.c_int()
        check(lib().fvb_get_spmv_format(self._h, C.byref(a), C.byref(k)))
        return {1: "csr", 2: "dia", 3: "dia_tma"}[a.value]

    def set_pcg_scaling(self, mode):
        """0 = automatic (cold-started steady Jacobi solves on the diagonal format run the symmetrically
        scaled recurrence, include/fvb200.h), 1 = never."""
        check(lib().fvb_set_pcg_scaling(self._h, C.c_int(int(mode))))

    def pcg_scaling(self):
        """-> True when the last solve on this system ran the scaled recurrence."""
        a = C.c_int()
        check(lib().fvb_get_pcg_scaling(self._h, C.byref(a)))
        return bool(a.value)

    def set_profiling(self, stride):
        check(lib().fvb_set_profiling(self._h, C.c_int(int(stride))))

    def timings(self):
        t = _lib.Timings()
        check(lib().fvb_get_timings(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in t._fields_}

    def sync(self):
        check(lib().fvb_sync(self._h))


class MultiSystem:
    """One problem spread over several GPUs of the box from ONE process and one calling thread (an fvb_multi):
    the whole-problem arrays go in exactly as into solvediffusion, the library partitions, connects the devices
    over NVLink peer memory and solves; see include/fvb200.h "single-process multi-GPU front end"."""

    def __init__(self, devices=None, ndev=None):
        if devices is None:
            if ndev is None:
                n = C.c_int()
                check(lib().fvb_device_count(C.byref(n)))
                ndev = n.value
            devices = list(range(int(ndev)))
        self.devices = [int(d) for d in devices]
        ids = (C.c_int * len(self.devices))(*self.devices)
        self._m = C.c_void_p()
        check(lib().fvb_multi_create(C.c_int(len(self.devices)), ids, C.byref(self._m)))
        self.n_nodes = 0

    def close(self):
        if getattr(self, "_m", None) is not None and self._m:
            lib().fvb_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_preconditioner(self, kind="jacobi", nu=0, omega=0.0, oc=0.0):
        check(lib().fvb_multi_set_preconditioner(self._m, C.c_int({"jacobi": 0, "mg": 1}[kind]), C.c_int(int(nu)),
                                                 C.c_double(omega), C.c_double(oc)))

    def assemble(self, neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                 metaindex=None, logtransformconductivity=False):
        nb = _pairs(neighbors)
        F = nb.size // 2
        aol, cond, src = f64(areasoverlengths).reshape(-1), f64(conductivities).reshape(-1), f64(sources).reshape(-1)
        dn, dh = i64(dirichletnodes), f64(dirichletheads)
        if aol.size != F:
            raise ValueError("areasoverlengths and neighbors differ in length")
        if dn.size != dh.size:
            raise ValueError("dirichletnodes and dirichletheads differ in length")
        meta = _metaindex_table(metaindex, F)
        if meta is None and cond.size < F:
            raise IndexError("conductivities is shorter than neighbors")
        check(lib().fvb_multi_assemble(self._m, C.c_int64(src.size), C.c_int64(F), ptr(nb), ptr(aol), ptr(cond),
                                       C.c_int64(cond.size), ptr(meta), C.c_int(int(bool(logtransformconductivity))),
                                       ptr(src), C.c_int64(dn.size), ptr(dn), ptr(dh)))
        self.n_nodes = src.size
        return self

    def assemble_regulargrid(self, mins, maxs, ns, nodehycos, sources, dirichletnodes, dirichletheads, logmean=True,
                             logtransformconductivity=True):
        mn, mx, nn = f64(mins), f64(maxs), i64(ns)
        k = f64(nodehycos).reshape(-1)
        src = None if sources is None else f64(sources).reshape(-1)
        dn, dh = i64(dirichletnodes), f64(dirichletheads)
        if k.size != int(np.prod(nn)):
            raise ValueError("nodehycos must hold one value per node")
        check(lib().fvb_multi_assemble_regulargrid(self._m, ptr(mn), ptr(mx), ptr(nn), ptr(k), C.c_int(int(bool(logmean))),
                                                   C.c_int(int(bool(logtransformconductivity))), ptr(src),
                                                   C.c_int64(dn.size), ptr(dn), ptr(dh)))
        self.n_nodes = k.size
        return self

    def sizes(self):
        nf, nnz, nd = C.c_int64(), C.c_int64(), C.c_int()
        lo = (C.c_int64 * len(self.devices))()
        hi = (C.c_int64 * len(self.devices))()
        check(lib().fvb_multi_sizes(self._m, C.byref(nf), C.byref(nnz), C.byref(nd), lo, hi))
        return dict(nf_global=nf.value, nnz_global=nnz.value, ndev=nd.value, node_ranges=list(zip(list(lo), list(hi))))

    def device_system(self, i):
        """A non-owning System view of device i's handle (formats, timings, assembly kind ...)."""
        hnd = C.c_void_p()
        check(lib().fvb_multi_device_handle(self._m, C.c_int(int(i)), C.byref(hnd)))
        s = System.__new__(System)
        s._h, s.device, s.nranks, s.rank = hnd, self.devices[i], len(self.devices), i
        s.node_lo, s.node_hi = self.sizes()["node_ranges"][i]
        s.close = lambda: None  # owned by the MultiSystem
        return s

    def solve(self, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER, want_x=False, hist_cap=None):
        sz = self.sizes()
        head = np.empty(self.n_nodes, np.float64)
        x = np.empty(sz["nf_global"], np.float64) if want_x else None
        cap = int(min(maxiter, 1 << 24) if hist_cap is None else hist_cap)
        hist = np.empty(max(cap, 1), np.float64)
        iters, conv = C.c_int64(), C.c_int()
        check(lib().fvb_multi_solve(self._m, C.c_double(rtol), C.c_int64(int(maxiter)), ptr(head), ptr(x), C.byref(iters),
                                    C.byref(conv), ptr(hist), C.c_int64(cap)))
        ch = ConvergenceHistory(bool(conv.value), int(iters.value), {"resnorm": hist[:min(iters.value, cap)].copy()})
        return head, x, ch

    def csr(self):
        sz = self.sizes()
        p = np.empty(sz["nf_global"] + 1, np.int64)
        idx = np.empty(sz["nnz_global"], np.int64)
        val = np.empty(sz["nnz_global"], np.float64)
        check(lib().fvb_multi_get_csr(self._m, ptr(p), ptr(idx), ptr(val)))
        return p, idx, val

    def b(self):
        out = np.empty(self.sizes()["nf_global"], np.float64)
        check(lib().fvb_multi_get_b(self._m, ptr(out)))
        return out

    def freenode(self):
        out = np.empty(self.n_nodes, np.uint8)
        check(lib().fvb_multi_get_freenode(self._m, ptr(out)))
        return out.astype(bool)


class SparseMatrixCSC:
    """The `A` of the reference's return tuples: a SparseMatrixCSC{Float64,Int64} image whose
    arrays are fetched from the GPU on first access (A is symmetric, so the CSR arrays kept
    on the device are its CSC arrays; src/FiniteVolume.jl:107)."""

    def __init__(self, system):
        self._sys = system  # a System or a MultiSystem
        s = system.sizes()
        self.m = self.n = s["nf_global"]
        self._arrays = None

    def _fetch(self):
        if self._arrays is None:
            self._arrays = self._sys.csr()
        return self._arrays

    @property
    def colptr(self):
        return self._fetch()[0]

    @property
    def rowval(self):
        return self._fetch()[1]

    @property
    def nzval(self):
        return self._fetch()[2]

    @property
    def shape(self):
        return (self.m, self.n)

    def toscipy(self):
        import scipy.sparse as sp
        p, i, v = self._fetch()
        return sp.csc_matrix((v, i - 1, p - 1), shape=(self.m, self.n))

    def __matmul__(self, x):
        """A * x on the GPU (examples/waffle/ex.jl:16 computes A*head[freenode]-b)."""
        return self._sys.spmv(x)


# ======================================================================================
# Reference-named functions
# ======================================================================================
def getfreenodes(n, dirichletnodes, device=0):
    """src/FiniteVolume.jl:32-44 -> (freenode::Vector{Bool}, nodei2freenodei::Vector{Int})."""
    s = System(device)
    try:
        s.assemble(np.empty((0, 2), np.int64), [], [], np.zeros(n), dirichletnodes, np.zeros(len(dirichletnodes)))
        return s.freenode(), s.nodei2freenodei()
    finally:
        s.close()


def assembleA(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
              metaindex=None, logtransformconductivity=False, device=0) -> SparseMatrixCSC:
    """src/FiniteVolume.jl:75-108."""
    s = System(device).assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                                metaindex, logtransformconductivity)
    return SparseMatrixCSC(s)


def assembleb(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
              metaindex=None, logtransformconductivity=False, device=0) -> np.ndarray:
    """src/FiniteVolume.jl:110-139."""
    s = System(device)
    try:
        s.assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                   metaindex, logtransformconductivity)
        return s.b()
    finally:
        s.close()


def freenodes2nodes(result, sources, dirichletnodes, dirichletheads, device=0):
    """src/FiniteVolume.jl:141-155 -> (head, freenode, nodei2freenodei)."""
    s = System(device)
    try:
        s.assemble(np.empty((0, 2), np.int64), [], [], sources, dirichletnodes, dirichletheads)
        s.vec_upload(0, result)
        return s.vec_to_nodes(0), s.freenode(), s.nodei2freenodei()
    finally:
        s.close()


def solvediffusion(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                   maxiter=DEFAULT_MAXITER, rtol=SQRT_EPS, metaindex=None, logtransformconductivity=False, device=0,
                   precond="jacobi", devices=None):
    """src/FiniteVolume.jl:157-165 -> (head, ch, A, b, freenode).

    Differences from the reference, both mandated by north_star: the preconditioner is
    Jacobi instead of Ruge-Stueben AMG, so `maxiter` counts Jacobi-PCG iterations (default
    raised from 400 accordingly); `rtol` exposes IterativeSolvers' `tol` (same default).
    precond="mg" selects the aggregation-multigrid V-cycle (box-structured grids),
    "auto" uses it when the matrix qualifies and Jacobi otherwise."""
    if devices is not None and len(devices) > 1:
        # one process, several GPUs (fvb_multi): same call, same return tuple
        ms = MultiSystem(devices)
        if precond in ("mg", "auto"):
            ms.set_preconditioner("mg")
        ms.assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex,
                    logtransformconductivity)
        head, _, ch = ms.solve(rtol=rtol, maxiter=maxiter)
        return head, ch, SparseMatrixCSC(ms), ms.b(), ms.freenode()
    s = System(device if devices is None else devices[0])
    if precond in ("mg", "auto"):
        s.set_preconditioner("mg")  # before assemble: silently stays on Jacobi when the matrix does not qualify
    s.assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
               metaindex, logtransformconductivity)
    if precond == "mg" and s.preconditioner()[0] not in ("mg", "amg"):
        raise _lib.FVBError(1, "no multigrid hierarchy could be built for this matrix; use precond='auto' or 'jacobi'")
    head, _, ch = s.solve(rtol=rtol, maxiter=maxiter)
    return head, ch, SparseMatrixCSC(s), s.b(), s.freenode()
