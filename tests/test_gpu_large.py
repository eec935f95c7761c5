"""BASELINE config 2 at full size (256^3 lognormal, single B200; 16.6 M unknowns, 116 M nonzeros): a direct
comparison with the CPU oracle (structure bit-exact, values <= 1e-14, heads <= 1e-8 at rtol 1e-12 on both sides;
the threaded oracle needs about a minute) and size-independent properties:
determinism, symmetry, row sums, the counts of SURVEY 8a, residual of the returned heads, maximum principle,
linearity in the Dirichlet data, agreement of both SpMV formats and both preconditioners, slab rows == whole rows."""
import os
import ctypes
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(fv):
    n = 256
    s = fv.System()
    nb, aol, _ = s.device_regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], want_volumes=False)
    N = n ** 3
    lnk = math.log(1e-5) + np.random.default_rng(0).standard_normal(N)
    kf = s.device_nodehycos2neighborhycos(nb, lnk, True)
    plane = n * n
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.ones(plane), np.zeros(plane)])
    src = np.zeros(N)
    s.assemble(nb, aol, kf, src, dn, dh, None, True)
    return dict(n=n, N=N, s=s, nb=nb, aol=aol, kf=kf, lnk=lnk, dn=dn, dh=dh, src=src, plane=plane)


def test_counts_and_determinism(fv, big):
    s = big["s"]
    sz = s.sizes()
    assert sz["nf_local"] == 16646144 and sz["nnz_local"] == 116131840
    assert s.spmv_format() == ("dia", 3)
    p, i, v = s.csr()
    s2 = fv.System().assemble(big["nb"], big["aol"], big["kf"], big["src"], big["dn"], big["dh"], None, True)
    p2, i2, v2 = s2.csr()
    assert np.array_equal(p, p2) and np.array_equal(i, i2) and np.array_equal(v, v2)  # bitwise, run to run
    assert np.array_equal(s.b(), s2.b())
    # structure: columns strictly ascending inside every row, diagonal present and positive
    assert np.all(np.diff(p) >= 4) and np.all(np.diff(p) <= 7)
    rows = np.repeat(np.arange(1, sz["nf_local"] + 1), np.diff(p))
    interior = np.ones(i.size, bool)
    interior[p[1:-1] - 1] = False  # first entry of each row
    assert np.all(np.diff(i)[interior[1:]] > 0)
    d = s.diag()
    assert np.array_equal(v[i == rows], d) and d.min() > 0
    # row sums: zero except next to the Dirichlet planes, where they equal the eliminated coupling (>0)
    rs = np.add.reduceat(v, p[:-1] - 1)
    pl = big["plane"]
    assert np.max(np.abs(rs[pl:-pl])) <= 1e-12 * d.max() and rs[:pl].min() > 0 and rs[-pl:].min() > 0
    # b is exactly that coupling times head 1 on the left plane, 0 elsewhere
    b = s.b()
    assert np.allclose(b[:pl], rs[:pl], rtol=1e-13) and not b[pl:].any()


def test_oracle_parity_256(fv, orc, big):
    """north_star's correctness bar at BASELINE config 2: CSR structure bit-exact, assembled values within 1e-14
    relative, heads within 1e-8 relative at matched residual tolerance (1e-12), against the oracle's serial
    restatement of src/FiniteVolume.jl:75-165 on the same seeded inputs."""
    n, N, s = big["n"], big["N"], big["s"]
    try:
        orc.set_num_threads(max(orc.num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        pass
    _, nb, aol, _ = orc.regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], want_coords=False)
    kf = orc.nodehycos2neighborhycos(nb, big["lnk"], True)
    # the device-side grid helpers produced the very same inputs
    assert np.array_equal(big["kf"].to_host(), kf) and np.array_equal(big["aol"].to_host(), aol)
    Ao = orc.assembleA(nb, aol, kf, big["src"], big["dn"], big["dh"], None, True)
    bo = orc.assembleb(nb, aol, kf, big["src"], big["dn"], big["dh"], None, True)
    p, i, v = s.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval)           # structure: bit-exact
    assert np.max(np.abs(v - Ao.nzval) / np.abs(Ao.nzval)) <= 1e-14                # values (log K: exp <= 1 ulp)
    b = s.b()
    assert np.max(np.abs(b - bo)) <= 1e-14 * np.max(np.abs(bo))
    del nb, aol, kf, p, i, v
    xo, cho = orc.cg(Ao, bo, Pl="jacobi", tol=1e-12, maxiter=100000, threaded=True)
    ho, _, _ = orc.freenodes2nodes(xo, big["src"], big["dn"], big["dh"])
    head, _, ch = s.solve(rtol=1e-12)
    assert ch.isconverged and cho.isconverged
    assert np.max(np.abs(head - ho)) <= 1e-8 * np.max(np.abs(ho))                  # heads at matched tolerance
    assert abs(ch.iters - cho.iters) <= max(5, cho.iters // 100)
    # the multigrid-preconditioned solve meets the same bar
    s.set_preconditioner("mg")
    head_mg, _, ch_mg = s.solve(rtol=1e-12)
    s.set_preconditioner("jacobi")
    assert ch_mg.isconverged and np.max(np.abs(head_mg - ho)) <= 1e-8 * np.max(np.abs(ho))


def test_symmetry_and_formats(fv, big):
    s = big["s"]
    nf = s.sizes()["nf_local"]
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(nf), rng.standard_normal(nf)
    Ax, Ay = s.spmv(x), s.spmv(y)
    assert abs(y @ Ax - x @ Ay) <= 1e-12 * (np.linalg.norm(x) * np.linalg.norm(Ay))
    s.set_spmv_format(1)
    Ax_csr = s.spmv(x)
    s.set_spmv_format(0)
    assert np.array_equal(Ax, Ax_csr)
    # linearity of the operator
    assert np.allclose(s.spmv(2.5 * x - y), 2.5 * Ax - Ay, rtol=1e-12, atol=1e-18)


def test_solve_properties(fv, big):
    s = big["s"]
    head, x, ch = s.solve(rtol=1e-10, want_x=True)
    assert ch.isconverged
    b = s.b()
    assert np.linalg.norm(s.spmv(x) - b) <= 2e-10 * np.linalg.norm(b)
    assert head.min() >= -1e-9 and head.max() <= 1 + 1e-9  # examples/box_model/ex_piml_data.jl:49-51
    fn = s.freenode()
    assert np.array_equal(head[fn], x) and np.array_equal(head[~fn], big["dh"])
    # multigrid-preconditioned CG: same heads, ~50x fewer iterations
    s.set_preconditioner("mg")
    head_mg, _, ch_mg = s.solve(rtol=1e-10)
    s.set_preconditioner("jacobi")
    assert ch_mg.isconverged and ch_mg.iters * 20 < ch.iters
    assert np.max(np.abs(head_mg - head)) <= 1e-8
    # linearity in the boundary data: heads (3, 1) = 1 + 2 * heads (1, 0)
    s.update_values(big["kf"].to_host(), dirichletheads=1 + 2 * big["dh"])
    head2, _, ch2 = s.solve(rtol=1e-10)
    s.update_values(big["kf"].to_host(), dirichletheads=big["dh"])
    assert ch2.isconverged and np.max(np.abs(head2 - (1 + 2 * head))) <= 1e-7


def test_slab_rows_equal_whole_rows(fv, big):
    import importlib
    dist = importlib.import_module("fvb200.distributed")
    n, N = big["n"], big["N"]
    p, i, v = big["s"].csr()
    pl = dist.slab_planes(n, 8)[3]
    lo, hi = dist.node_range_of_planes(pl, n, n)
    sl = fv.System()
    nbs, aols, _ = sl.device_regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], planes=pl, want_volumes=False)
    k_lo, k_hi = lo - big["plane"], hi + big["plane"]
    kfs = sl.device_nodehycos2neighborhycos(nbs, big["lnk"][k_lo - 1:k_hi], True, node_lo=k_lo)
    sl.assemble(nbs, aols, kfs, big["src"][lo - 1:hi], big["dn"], big["dh"], None, True, n_nodes=N, node_range=(lo, hi))
    sz = sl.sizes()
    ps, is_, vs = sl.csr()
    r0 = sz["row_start"] - 1
    a, b = p[r0] - 1, p[r0 + sz["nf_local"]] - 1
    assert np.array_equal(ps - 1 + a, p[r0:r0 + sz["nf_local"] + 1] - 1)
    assert np.array_equal(is_, i[a:b]) and np.array_equal(vs, v[a:b])
    assert sz["n_halo"] == 2 * big["plane"] and sl.spmv_format() == ("dia", 3)


def test_size_guards(fv):
    """Per-GPU parts beyond 32-bit local indexing are refused before any memory is touched."""
    L = fv._lib.lib()
    s = fv.System()
    one = np.ones(4)
    nbp = np.array([1, 2], np.int64)
    st = L.fvb_assemble(s._h, ctypes.c_int64(4), ctypes.c_int64(1), ctypes.c_int64(4), ctypes.c_int64(2 ** 30),
                        fv._lib.ptr(nbp), fv._lib.ptr(one), fv._lib.ptr(one), ctypes.c_int64(4), None, ctypes.c_int(0),
                        fv._lib.ptr(one), ctypes.c_int64(0), None, None)
    assert st == 1 and b"use more ranks" in L.fvb_last_error()
    st = L.fvb_assemble(s._h, ctypes.c_int64(2 ** 31), ctypes.c_int64(1), ctypes.c_int64(2 ** 31), ctypes.c_int64(1),
                        fv._lib.ptr(nbp), fv._lib.ptr(one), fv._lib.ptr(one), ctypes.c_int64(4), None, ctypes.c_int(0),
                        fv._lib.ptr(one), ctypes.c_int64(0), None, None)
    assert st == 1
