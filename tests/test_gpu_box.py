"""The closed-form ("box") assembly of regulargrid-ordered problems (csrc/box.cuh) against the general
adjacency/CSR path and the CPU oracle: rows written straight into the diagonal format must give bit-identical
b, diag(A), products and -- through the lazily built CSR image -- bit-identical colptr/rowval/nzval; anything that is
not exactly regulargrid's list, or whose Dirichlet set leaves the diagonals, must fall back to the general path."""
import importlib
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def problem(fv, ns, sigma=1.0, seed=0, mins=None, maxs=None):
    mins = [0, 0, 0] if mins is None else mins
    maxs = [n - 1 for n in ns] if maxs is None else maxs
    _, nb, aol, vol = fv.regulargrid(mins, maxs, ns, want_coords=False)
    N = int(np.prod(ns))
    lnk = math.log(1e-5) + sigma * np.random.default_rng(seed).standard_normal(N)
    kf = fv.nodehycos2neighborhycos(nb, lnk, True)
    plane = ns[1] * ns[2]
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.ones(plane), np.zeros(plane)])
    src = 1e-7 * np.random.default_rng(seed + 1).standard_normal(N)
    src[dn - 1] = 0
    return dict(nb=nb, aol=aol, kf=kf, src=src, dn=dn, dh=dh, vol=vol, lnk=lnk, N=N, plane=plane)


def both_paths(fv, P, logk=True, cond=None, **kw):
    cond = P["kf"] if cond is None else cond
    sb = fv.System().assemble(P["nb"], P["aol"], cond, P["src"], P["dn"], P["dh"], None, logk, **kw)
    sg = fv.System()
    sg.set_assembly(1)
    sg.assemble(P["nb"], P["aol"], cond, P["src"], P["dn"], P["dh"], None, logk, **kw)
    return sb, sg


def assert_same_system(sb, sg):
    assert sb.sizes() == sg.sizes()
    assert np.array_equal(sb.b(), sg.b()) and np.array_equal(sb.diag(), sg.diag())
    assert sb.spmv_format() == sg.spmv_format() == ("dia", 3)
    x = np.random.default_rng(5).standard_normal(sb.sizes()["nf_local"])
    assert np.array_equal(sb.spmv(x), sg.spmv(x))             # both on their diagonal copies
    pb, ib, vb = sb.csr()                                       # built on demand from the diagonals
    pg, ig, vg = sg.csr()
    assert np.array_equal(pb, pg) and np.array_equal(ib, ig) and np.array_equal(vb, vg)
    assert np.array_equal(sb.halo_cols(), sg.halo_cols())
    assert np.array_equal(sb.freenode(), sg.freenode()) and np.array_equal(sb.nodei2freenodei(), sg.nodei2freenodei())


@pytest.mark.parametrize("ns", [[6, 5, 4], [9, 2, 7], [40, 16, 16], [12, 33, 3]])
@pytest.mark.parametrize("logk", [True, False])
def test_box_rows_equal_general_rows(fv, ns, logk):
    P = problem(fv, ns)
    cond = P["kf"] if logk else np.exp(P["kf"])
    sb, sg = both_paths(fv, P, logk, cond)
    assert sg.assembly() == "general"
    assert sb.assembly() == "box"
    assert_same_system(sb, sg)
    # forced CSR product on the box problem (lazy CSR) == diagonal product
    x = np.random.default_rng(6).standard_normal(sb.sizes()["nf_local"])
    y = sb.spmv(x)
    sb.set_spmv_format(1)
    assert np.array_equal(sb.spmv(x), y)
    sb.set_spmv_format(0)
    hb, _, chb = sb.solve(rtol=1e-12)
    hg, _, chg = sg.solve(rtol=1e-12)
    assert chb.isconverged and chg.isconverged and chb.iters == chg.iters and np.array_equal(hb, hg)


@pytest.mark.parametrize("ns", [[2, 2, 2], [3, 2, 2], [3, 3, 3], [4, 2, 3]])
def test_box_tiny_grids_against_oracle(fv, orc, ns):
    """The smallest grids regulargrid accepts (every ns[k] >= 2), Dirichlet on the left plane only."""
    P = problem(fv, ns)
    dn, dh = P["dn"][:P["plane"]], np.linspace(1, 2, P["plane"])
    src = P["src"].copy(); src[dn - 1] = 0
    s = fv.System().assemble(P["nb"], P["aol"], P["kf"], src, dn, dh, None, True)
    Ao = orc.assembleA(P["nb"], P["aol"], P["kf"], src, dn, dh, None, True)
    bo = orc.assembleb(P["nb"], P["aol"], P["kf"], src, dn, dh, None, True)
    p, i, v = s.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.allclose(v, Ao.nzval, rtol=1e-14, atol=0)
    assert np.allclose(s.b(), bo, rtol=1e-14, atol=0)
    ho, cho, *_ = orc.solvediffusion(P["nb"], P["aol"], P["kf"], src, dn, dh, maxiter=1000, tol=1e-12, logtransformconductivity=True)
    h, _, ch = s.solve(rtol=1e-12)
    assert ch.isconverged and np.max(np.abs(h - ho)) <= 1e-8 * np.max(np.abs(ho))


def test_box_against_oracle_with_explicit_zeros(fv, orc):
    """Plain K with some exactly-zero conductivities: sparse(...,+) keeps the explicit zeros
    (src/FiniteVolume.jl:107), so must the CSR image built from the diagonals."""
    ns = [7, 6, 5]
    P = problem(fv, ns)
    k = np.exp(P["kf"])
    k[::7] = 0.0
    s = fv.System().assemble(P["nb"], P["aol"], k, P["src"], P["dn"], P["dh"])
    assert s.assembly() == "box"
    Ao = orc.assembleA(P["nb"], P["aol"], k, P["src"], P["dn"], P["dh"])
    bo = orc.assembleb(P["nb"], P["aol"], k, P["src"], P["dn"], P["dh"])
    p, i, v = s.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.array_equal(v, Ao.nzval)
    assert np.array_equal(s.b(), bo)
    assert s.sizes()["nnz_local"] == Ao.rowval.size and np.count_nonzero(v == 0) > 0


def test_box_falls_back(fv, orc):
    ns = [8, 6, 5]
    P = problem(fv, ns)
    k = np.exp(P["kf"])
    # (a) interior Dirichlet nodes break the uniform node -> row shift
    dn = np.concatenate([P["dn"], [P["plane"] * 3 + 7]])
    dh = np.concatenate([P["dh"], [0.5]])
    src = P["src"].copy(); src[dn - 1] = 0
    s = fv.System().assemble(P["nb"], P["aol"], k, src, dn, dh)
    assert s.assembly() == "general"
    Ao = orc.assembleA(P["nb"], P["aol"], k, src, dn, dh)
    assert np.array_equal(s.csr()[1], Ao.rowval) and np.array_equal(s.csr()[2], Ao.nzval)
    # (b) same faces in another order / one face flipped / one face dropped
    perm = np.random.default_rng(0).permutation(P["nb"].shape[0])
    s = fv.System().assemble(P["nb"][perm], P["aol"][perm], k[perm], P["src"], P["dn"], P["dh"])
    assert s.assembly() == "general"
    nb2 = P["nb"].copy(); nb2[100] = nb2[100, ::-1]
    s = fv.System().assemble(nb2, P["aol"], k, P["src"], P["dn"], P["dh"])
    assert s.assembly() == "general"
    Ao = orc.assembleA(nb2, P["aol"], k, P["src"], P["dn"], P["dh"])
    assert np.array_equal(s.csr()[1], Ao.rowval) and np.array_equal(s.csr()[2], Ao.nzval)
    s = fv.System().assemble(P["nb"][:-1], P["aol"][:-1], k[:-1], P["src"], P["dn"], P["dh"])
    assert s.assembly() == "general"
    # (c) Dirichlet prefix that is not a whole plane still qualifies (shift uniform over free-free faces)
    dn3 = np.arange(1, P["plane"] + 4)
    src3 = P["src"].copy(); src3[dn3 - 1] = 0
    s = fv.System().assemble(P["nb"], P["aol"], k, src3, dn3, np.linspace(1, 2, dn3.size))
    Ao = orc.assembleA(P["nb"], P["aol"], k, src3, dn3, np.linspace(1, 2, dn3.size))
    bo = orc.assembleb(P["nb"], P["aol"], k, src3, dn3, np.linspace(1, 2, dn3.size))
    p, i, v = s.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.array_equal(v, Ao.nzval)
    assert np.array_equal(s.b(), bo)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_box_slabs(fv, orc, nranks):
    """Every 'rank' assembles its slab by the closed form: rows, halo columns and b equal the general path's;
    stacked they equal the oracle's unpartitioned matrix."""
    dist = importlib.import_module("fvb200.distributed")
    ns = [11, 6, 5]
    P = problem(fv, ns)
    N = P["N"]
    k = np.exp(P["kf"])
    Ao = orc.assembleA(P["nb"], P["aol"], k, P["src"], P["dn"], P["dh"])
    bo = orc.assembleb(P["nb"], P["aol"], k, P["src"], P["dn"], P["dh"])
    ptr, idx, val, b = [np.array([1])], [], [], []
    for pl in dist.slab_planes(ns[0], nranks):
        lo, hi = dist.node_range_of_planes(pl, ns[1], ns[2])
        _, nbs, aols, _ = fv.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False, planes=pl)
        ks = np.exp(0.5 * (P["lnk"][nbs[:, 0] - 1] + P["lnk"][nbs[:, 1] - 1]))
        Q = dict(nb=nbs, aol=aols, kf=ks, src=P["src"][lo - 1:hi], dn=P["dn"], dh=P["dh"])
        sb, sg = both_paths(fv, Q, False, ks, n_nodes=N, node_range=(lo, hi))
        assert sb.assembly() == "box" and sg.assembly() == "general"
        assert_same_system(sb, sg)
        p, i, v = sb.csr()
        ptr.append(p[1:] - 1 + ptr[-1][-1]); idx.append(i); val.append(v); b.append(sb.b())
    assert np.array_equal(np.concatenate(ptr), Ao.colptr) and np.array_equal(np.concatenate(idx), Ao.rowval)
    assert np.allclose(np.concatenate(val), Ao.nzval, rtol=1e-15, atol=0) and np.allclose(np.concatenate(b), bo, rtol=1e-15, atol=0)


def test_box_slab_with_dirichlet_only_neighbour_plane(fv):
    """A slab whose upper neighbour plane is the Dirichlet plane (the last rank owns that plane alone): the off-rank
    heads enter b through the sorted Dirichlet table."""
    ns = [5, 4, 3]
    P = problem(fv, ns)
    N, plane = P["N"], P["plane"]
    k = np.exp(P["kf"])
    pl = (3, 4)  # planes 3..4 of 5: plane 5 (above) is entirely Dirichlet, plane 2 (below) entirely free
    lo, hi = (pl[0] - 1) * plane + 1, pl[1] * plane
    _, nbs, aols, _ = fv.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False, planes=pl)
    ks = np.exp(0.5 * (P["lnk"][nbs[:, 0] - 1] + P["lnk"][nbs[:, 1] - 1]))
    dh = np.concatenate([np.ones(plane), np.linspace(2, 3, plane)])
    Q = dict(nb=nbs, aol=aols, kf=ks, src=P["src"][lo - 1:hi], dn=P["dn"], dh=dh)
    sb, sg = both_paths(fv, Q, False, ks, n_nodes=N, node_range=(lo, hi))
    assert sb.assembly() == "box"
    assert_same_system(sb, sg)
    assert sb.sizes()["n_halo"] == plane


def test_box_update_values(fv):
    ns = [10, 7, 6]
    P = problem(fv, ns)
    s = fv.System().assemble(P["nb"], P["aol"], P["kf"], P["src"], P["dn"], P["dh"], None, True)
    assert s.assembly() == "box"
    p0, i0, v0 = s.csr()
    kf2 = P["kf"] + 0.3 * np.random.default_rng(9).standard_normal(P["kf"].size)
    dh2 = P["dh"] * 2 + 1
    s.update_values(kf2, dirichletheads=dh2)
    fresh = fv.System().assemble(P["nb"], P["aol"], kf2, P["src"], P["dn"], dh2, None, True)
    p1, i1, v1 = s.csr()
    pf, if_, vf = fresh.csr()
    assert np.array_equal(p1, pf) and np.array_equal(i1, if_) and np.array_equal(v1, vf) and not np.array_equal(v1, v0)
    assert np.array_equal(s.b(), fresh.b())
    h1, _, c1 = s.solve(rtol=1e-11)
    h2, _, c2 = fresh.solve(rtol=1e-11)
    assert c1.isconverged and np.array_equal(h1, h2) and c1.iters == c2.iters


@pytest.mark.parametrize("mins,maxs,ns", [([0, 0, 0], [9, 7, 5], [10, 8, 6]), ([-50, -50, 0], [50, 50, 10], [20, 10, 2]),
                                           ([0, 0, 0], [1, 1, 1], [3, 4, 5])])
@pytest.mark.parametrize("logmean,logk", [(True, True), (False, False)])
def test_implicit_grid_equals_explicit(fv, orc, mins, maxs, ns, logmean, logk):
    """fvb_assemble_regulargrid (no face arrays) == regulargrid + nodehycos2neighborhycos + assemble, bit for bit,
    whole grid and slabs; heads vs the oracle."""
    dist = importlib.import_module("fvb200.distributed")
    N = int(np.prod(ns))
    plane = ns[1] * ns[2]
    rng = np.random.default_rng(3)
    nodek = (math.log(1e-5) + rng.standard_normal(N)) if logk else np.exp(rng.standard_normal(N))
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.ones(plane), np.zeros(plane)])
    src = 1e-7 * rng.standard_normal(N); src[dn - 1] = 0
    _, nb, aol, _ = orc.regulargrid(mins, maxs, ns, want_coords=False)
    kf = orc.nodehycos2neighborhycos(nb, nodek, logmean)
    se = fv.System().assemble(nb, aol, kf, src, dn, dh, None, logk)
    si = fv.System().assemble_regulargrid(mins, maxs, ns, nodek, src, dn, dh, logmean=logmean, logtransformconductivity=logk)
    assert si.assembly() == "implicit" and se.assembly() == "box"
    pe, ie, ve = se.csr()
    pi, ii, vi = si.csr()
    assert np.array_equal(pe, pi) and np.array_equal(ie, ii) and np.array_equal(ve, vi)
    assert np.array_equal(se.b(), si.b()) and np.array_equal(se.diag(), si.diag())
    hi_, _, chi = si.solve(rtol=1e-12)
    ho, cho, *_ = orc.solvediffusion(nb, aol, kf, src, dn, dh, maxiter=20000, tol=1e-12, logtransformconductivity=logk)
    assert chi.isconverged and np.max(np.abs(hi_ - ho)) <= 1e-8 * np.max(np.abs(ho))
    # zero sources may be passed as None
    s0 = fv.System().assemble_regulargrid(mins, maxs, ns, nodek, None, dn, dh, logmean=logmean, logtransformconductivity=logk)
    sz = fv.System().assemble(nb, aol, kf, np.zeros(N), dn, dh, None, logk)
    assert np.array_equal(s0.b(), sz.b())
    # slabs
    if ns[0] >= 6:
        for pl in dist.slab_planes(ns[0], 3):
            lo, hi = dist.node_range_of_planes(pl, ns[1], ns[2])
            k_lo, k_hi = max(1, lo - plane), min(N, hi + plane)
            ss = fv.System().assemble_regulargrid(mins, maxs, ns, nodek[k_lo - 1:k_hi], src[lo - 1:hi], dn, dh, logmean=logmean,
                                                  logtransformconductivity=logk, planes=pl)
            sz_ = ss.sizes()
            r0 = sz_["row_start"] - 1
            ps, is_, vs = ss.csr()
            a, b_ = pe[r0] - 1, pe[r0 + sz_["nf_local"]] - 1
            assert np.array_equal(is_, ie[a:b_]) and np.array_equal(vs, ve[a:b_])
            assert np.array_equal(ss.b(), se.b()[r0:r0 + sz_["nf_local"]])
    # values-only update with new node values
    nodek2 = nodek + (0.1 if logk else 0.05)
    si.update_values(nodek2)
    kf2 = orc.nodehycos2neighborhycos(nb, nodek2, logmean)
    se2 = fv.System().assemble(nb, aol, kf2, src, dn, dh, None, logk)
    assert np.array_equal(si.csr()[2], se2.csr()[2]) and np.array_equal(si.b(), se2.b())


def test_implicit_grid_rejects_unrepresentable_dirichlet_sets(fv):
    ns = [6, 5, 4]
    N = int(np.prod(ns))
    with pytest.raises(fv.FVBError) as e:
        fv.System().assemble_regulargrid([0, 0, 0], [5, 4, 3], ns, np.zeros(N), None, [1, 2, 50], [1.0, 1.0, 0.0])
    assert e.value.status == 1 and "fvb_assemble" in str(e.value)
    with pytest.raises(fv.FVBError):  # a source on a Dirichlet node is still the reference's error
        src = np.zeros(N); src[0] = 1.0
        fv.System().assemble_regulargrid([0, 0, 0], [5, 4, 3], ns, np.zeros(N), src, [1], [1.0])
