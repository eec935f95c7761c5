"""Pins the CPU oracle against every known answer the reference's own tests hold for the
hot path.  No GPU needed.  Reference paths are relative to its checkout."""
import math

import numpy as np
import pytest


def test_runtests_chain(orc):
    """test/runtests.jl:4-16 -- 4-node chain, neighbours listed in both directions (exercises the
    duplicate-summing of sparse(...,+)): h == [1, 2/3, 1/3, 0]."""
    nb = [(1, 2), (2, 1), (2, 3), (3, 2), (3, 4), (4, 3)]
    h, ch, A, b, fn = orc.solvediffusion(nb, np.ones(6), np.ones(6), np.zeros(4), [1, 4], [1.0, 0.0])
    assert np.allclose(h, [1.0, 2 / 3, 1 / 3, 0.0], rtol=math.sqrt(np.finfo(float).eps))
    assert np.array_equal(A.toscipy().toarray(), [[4.0, -2.0], [-2.0, 4.0]])
    assert np.array_equal(b, [2.0, 0.0])
    assert list(fn) == [False, True, True, False]
    assert ch.isconverged


def test_source_on_dirichlet_node_raises(orc):
    """src/FiniteVolume.jl:25-27."""
    with pytest.raises(ValueError, match="source at a Dirichlet node"):
        orc.assembleb([(1, 2)], [1.0], [1.0], [1.0, 0.0], [1], [0.0])


def test_sparse_semantics(orc):
    """sparse(I,J,V,m,n,+): left fold in input order, explicit zeros kept, rows ascending."""
    I = [2, 1, 2, 2, 1]
    J = [1, 1, 1, 2, 1]
    V = [1e16, 1.0, -1e16, 3.0, -1.0]
    A = orc.sparse(I, J, V, 2, 2)
    assert list(A.colptr) == [1, 3, 4]
    assert list(A.rowval) == [1, 2, 2]
    assert list(A.nzval) == [0.0, 0.0, 3.0]  # (1+-1) kept as explicit zero; (1e16 + -1e16) = 0
    # order matters: 1 + 1e16 - 1e16 == 0, but 1e16 - 1e16 + 1 == 1
    B = orc.sparse([1, 1, 1], [1, 1, 1], [1.0, 1e16, -1e16], 1, 1)
    C = orc.sparse([1, 1, 1], [1, 1, 1], [1e16, -1e16, 1.0], 1, 1)
    assert B.nzval[0] == 0.0 and C.nzval[0] == 1.0


def test_assembly_against_dense_and_scipy(orc):
    rng = np.random.default_rng(7)
    N, F, ND = 300, 1500, 25
    nb = rng.integers(1, N + 1, size=(F, 2))  # duplicates, both directions and self loops occur
    aol, k = rng.random(F) + 0.1, rng.random(F) + 0.1
    dn = rng.choice(np.arange(1, N + 1), ND, replace=False)
    dh = rng.random(ND)
    src = rng.random(N)
    src[dn - 1] = 0
    A = orc.assembleA(nb, aol, k, src, dn, dh)
    b = orc.assembleb(nb, aol, k, src, dn, dh)
    fn, n2f = orc.getfreenodes(N, dn)
    D = np.zeros((A.n, A.n))
    bb = src[fn].copy()
    n2d = -np.ones(N, int)
    n2d[dn - 1] = np.arange(ND)
    for i, (a, c) in enumerate(nb - 1):
        cc = k[i] * aol[i]
        if fn[a] and fn[c]:
            r1, r2 = n2f[a] - 1, n2f[c] - 1
            D[r1, r1] += cc; D[r1, r2] -= cc; D[r2, r2] += cc; D[r2, r1] -= cc
        elif fn[a]:
            D[n2f[a] - 1, n2f[a] - 1] += cc; bb[n2f[a] - 1] += cc * dh[n2d[c]]
        elif fn[c]:
            D[n2f[c] - 1, n2f[c] - 1] += cc; bb[n2f[c] - 1] += cc * dh[n2d[a]]
    S = A.toscipy()
    assert np.allclose(S.toarray(), D, rtol=1e-13, atol=1e-15)
    assert np.allclose(b, bb, rtol=1e-13)
    assert abs(S - S.T).max() == 0.0
    for j in range(A.n):
        assert np.all(np.diff(A.rowval[A.colptr[j] - 1:A.colptr[j + 1] - 1]) > 0)


def test_fourfractures_known_counts(orc, fourfractures):
    """Nf=2076, nnz=14528, Jacobi-PCG converges; PFLOTRAN heads agree to ~1e-2."""
    m = fourfractures
    src = np.zeros(m["xs"].size)
    h, ch, A, b, fn = orc.solvediffusion(m["neighbors"], m["areasoverlengths"], m["conductivities"], src,
                                         m["dirichletnodes"], m["dirichletheads"], maxiter=5000)
    assert A.n == 2076 and A.nzval.size == 14528
    assert ch.isconverged and 100 < ch.iters < 400
    assert h.min() >= 1e6 - 1e-3 and h.max() <= 2e6 + 1e-3  # maximum principle
    assert np.max(np.abs(h - m["pflotran_h"])) / 2e6 < 1.5e-2


def W(u):
    if u <= 1:
        return (-math.log(u) - 0.57721566 + 0.99999193 * u - 0.24991055 * u**2 + 0.05519968 * u**3
                - 0.00976004 * u**4 + 0.00107857 * u**5)
    return (u**2 + 2.334733 * u + 0.250621) / (u**2 + 3.330657 * u + 1.681534) * math.exp(-u) / u


def theis_setup(orc_or_fv):
    """test/theis.jl:21-50 verbatim."""
    steadyhead, sidelength, thickness = 1e3, 50.0, 10.0
    mins, maxs, ns = [-sidelength, -sidelength, 0], [sidelength, sidelength, thickness], [101, 101, 2]
    k, Q, Ss = 1e-5, 1e-3, 0.1
    coords, nb, aol, vol = orc_or_fv.regulargrid(mins, maxs, ns)
    hycos = np.full(aol.size, k)
    N = coords.shape[1]
    src = np.zeros(N)
    center = [i for i in range(N) if coords[0, i] == 0 and coords[1, i] == 0]
    src[center[0]] = -Q / (2 * len(center) - 2)
    src[center[-1]] = -Q / (2 * len(center) - 2)
    src[center[1:-1]] = -2 * Q / (2 * len(center) - 2)
    r = np.hypot(coords[0], coords[1])
    dn = np.nonzero(r - sidelength >= 0)[0] + 1
    dh = np.full(dn.size, steadyhead)
    good = np.nonzero((coords[2] == thickness) & (coords[1] == 0) & (coords[0] > 0.1) & (coords[0] <= sidelength / 2))[0]
    return dict(coords=coords, nb=nb, aol=aol, vol=vol, hycos=hycos, src=src, dn=dn, dh=dh, good=good, k=k, Q=Q,
                Ss=Ss, thickness=thickness, sidelength=sidelength, steadyhead=steadyhead)


def theis_expected(P, t):
    rs = P["coords"][0, P["good"]]
    T = P["thickness"] * P["k"]
    S = P["Ss"] * P["thickness"]
    theis = np.array([P["Q"] * W(r**2 * S / (4 * T * t)) / (4 * math.pi * T) for r in rs])
    thiem = np.array([P["Q"] * math.log(P["sidelength"] / r) / (2 * math.pi * T) for r in rs])
    return theis, thiem


def test_thiem_theis(orc):
    """test/theis.jl:52-65 -- steady solution vs Thiem, transient (10 days, dt0=60, atol=1e-4) vs
    Theis, both with isapprox(atol=1e-4, rtol=2e-2)."""
    P = theis_setup(orc)
    assert P["dn"].size == 4752
    usteady, ch, A, b, fn = orc.solvediffusion(P["nb"], P["aol"], P["hycos"], P["src"], P["dn"], P["dh"],
                                               maxiter=20000, tol=1e-12)
    assert A.n == 15650 and A.nzval.size == 93108
    tend = 60 * 60 * 24 * 1e1
    theis, thiem = theis_expected(P, tend)
    steady_dd = P["steadyhead"] - usteady[P["good"]]
    assert np.linalg.norm(thiem - steady_dd) <= max(1e-4, 2e-2 * max(np.linalg.norm(thiem), np.linalg.norm(steady_dd)))
    u0 = np.full(P["src"].size, P["steadyhead"])
    solver = orc.DirectSolver()
    us, ts = orc.backwardeulerintegrate(u0, (0.0, tend), P["Ss"], P["vol"], P["nb"], P["aol"], P["hycos"], P["src"],
                                        P["dn"], P["dh"], atol=1e-4, dt0=60.0, linearsolver=solver)
    assert ts[-1] == tend
    assert len(ts) - 1 == 1090 and solver.solves == 3283  # SURVEY App. D trajectory
    model_dd = P["steadyhead"] - us[-1][P["good"]]
    assert np.linalg.norm(theis - model_dd) <= max(1e-4, 2e-2 * max(np.linalg.norm(theis), np.linalg.norm(model_dd)))


def onenode_problem():
    """test/onenodeadjoint.jl:13-28."""
    return dict(Ss=1.0, volumes=[1.0, 1.0], neighbors=[(1, 2)], aol=[1.0], loghycos=np.array([0.0]),
                sources=[0.0, 1.0], dn=[1], dh=[0.0], u0=[0.0, 0.0], tspan=(0.0, 1.0), atol=1e-8, dt0=1e-3)


def test_onenode_forward(orc):
    """test/onenodeadjoint.jl:29-44 -- u2(t) = 1 - exp(-t); with log K = 1: (1 - exp(-e t))/e."""
    p = onenode_problem()
    us, ts = orc.backwardeulerintegrate(p["u0"], p["tspan"], p["Ss"], p["volumes"], p["neighbors"], p["aol"],
                                        p["loghycos"], p["sources"], p["dn"], p["dh"], None, True, atol=p["atol"],
                                        dt0=p["dt0"])
    for u, t in zip(us, ts):
        assert math.isclose(u[1], 1 - math.exp(-t), rel_tol=1e-4, abs_tol=0.0) or t == 0.0
    us, ts = orc.backwardeulerintegrate(p["u0"], p["tspan"], p["Ss"], p["volumes"], p["neighbors"], p["aol"],
                                        p["loghycos"] + 1, p["sources"], p["dn"], p["dh"], None, True, atol=p["atol"],
                                        dt0=p["dt0"])
    for u, t in zip(us, ts):
        assert math.isclose(u[1], (1 - math.exp(-math.e * t)) / math.e, rel_tol=1e-4) or t == 0.0


def test_onenode_adjoint(orc):
    """test/onenodeadjoint.jl:56-64 -- lambda(t) against the closed form."""
    p = onenode_problem()
    sigma2 = 0.01**2
    u_init = lambda s: (1 - math.exp(-math.e * s)) / math.e  # noqa: E731
    u_obs = lambda s: 1 - math.exp(-s)  # noqa: E731
    f = lambda s: 2 * sigma2 * (u_init(s) - u_obs(s))  # noqa: E731
    lam, tl = orc.adjointintegrate(lambda t: np.array([f(t)]), p["tspan"], p["Ss"], p["volumes"], p["neighbors"],
                                   p["aol"], p["loghycos"] + 1, p["sources"], p["dn"], p["dh"], None, True,
                                   atol=p["atol"], dt0=p["dt0"])
    from scipy.integrate import quad
    T = p["tspan"][1]
    for lv, t in zip(lam, tl):
        s_ = T - t
        gamma = math.exp(-math.e * s_) * quad(lambda s: math.exp(math.e * s) * f(T - s), 0, s_)[0]
        assert math.isclose(lv[0], gamma, rel_tol=1e-4, abs_tol=1e-7)


def test_regulargrid_counts(orc):
    """src/grid.jl:69-70 sizes, :60 node order, +x/+y/+z emission order."""
    coords, nb, aol, vol = orc.regulargrid([0, 0, 0], [3, 2, 1], [4, 3, 2])
    assert nb.shape[0] == 3 * 24 - 12 - 8 - 6
    assert list(map(tuple, nb[:3])) == [(1, 7), (1, 3), (1, 2)]
    assert np.isclose(vol.sum(), 3 * 2 * 1)
    assert coords.shape == (3, 24) and list(coords[:, 1]) == [0.0, 0.0, 1.0]


def _scaled_cg_numpy(S, b, tol, maxiter):
    """Plain CG on A^ = D^-1/2 A D^-1/2, x^ = D^1/2 x, b^ = D^-1/2 b, stopping on the TRUE residual norm
    sqrt(sum d_i r^_i^2) -- the recurrence libfvb200 runs for cold-started steady solves (csrc/pcg.cuh, SC = true),
    restated in numpy with the same order of operations (rho = r^.r^, u^ = r^ + beta u^, c^ = A^ u^,
    alpha = rho / u^.c^, x^ += alpha u^, r^ -= alpha c^)."""
    import scipy.sparse as sp
    d = S.diagonal()
    s = 1.0 / np.sqrt(d)
    Ah = (sp.diags(s) @ S @ sp.diags(s)).tocsr()
    Ah.setdiag(1.0)
    r = s * b
    x = np.zeros_like(b)
    u = np.zeros_like(b)
    rho, rho_prev = float(r @ r), 1.0
    resid0 = float(np.sqrt(b @ b))
    hist = []
    for it in range(maxiter):
        beta = 0.0 if it == 0 else rho / rho_prev
        u = r + beta * u
        c = Ah @ u
        alpha = rho / float(u @ c)
        x += alpha * u
        r -= alpha * c
        rho_prev, rho = rho, float(r @ r)
        res = float(np.sqrt(np.sum(d * r * r)))
        hist.append(res)
        if res <= tol * resid0:
            break
    return s * x, np.array(hist)


@pytest.mark.parametrize("case", ["box_sigma1", "box_sigma3", "fourfractures"])
def test_scaled_recurrence_is_the_same_iteration(orc, fourfractures, case):
    """The claim behind fvb_set_pcg_scaling (include/fvb200.h): Jacobi-PCG on A and CG on D^-1/2 A D^-1/2 are the
    same iteration.  Checked on the CPU against the oracle's IterativeSolvers-style Jacobi-PCG: equal iteration
    counts (+-1), equal residual histories, equal heads -- on mild and strong heterogeneity and on the irregular
    fracture graph, at the default and at a tight tolerance."""
    if case == "fourfractures":
        m = fourfractures
        nb, aol, k, dn, dh = m["neighbors"], m["areasoverlengths"], m["conductivities"], m["dirichletnodes"], m["dirichletheads"]
        src, logk = np.zeros(m["xs"].size), False
    else:
        ns = [14, 12, 10]
        _, nb, aol, _ = orc.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
        N = int(np.prod(ns))
        sigma = 1.0 if case == "box_sigma1" else 3.0
        lnk = np.log(1e-5) + sigma * np.random.default_rng(0).standard_normal(N)
        k = orc.nodehycos2neighborhycos(nb, lnk, True)
        plane = ns[1] * ns[2]
        dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
        dh = np.concatenate([np.ones(plane), np.zeros(plane)])
        src, logk = np.zeros(N), True
    A = orc.assembleA(nb, aol, k, src, dn, dh, None, logk)
    b = orc.assembleb(nb, aol, k, src, dn, dh, None, logk)
    S = A.toscipy().tocsr()
    for tol in (np.sqrt(np.finfo(float).eps), 1e-12):
        xo, cho = orc.cg(A, b, tol=tol, maxiter=100000)
        xs, hist = _scaled_cg_numpy(S, b, tol, 100000)
        assert cho.isconverged and abs(len(hist) - cho.iters) <= 1
        n = min(len(hist), cho.iters) - 1
        assert np.allclose(hist[:n], cho.data["resnorm"][:n], rtol=1e-6)
        assert np.max(np.abs(xs - xo)) <= 1e-8 * np.max(np.abs(xo))


def test_amg_oracle_on_fourfractures(orc, fourfractures):
    """oracle/amg_oracle.py (numpy restatement of csrc/amg.cuh, the checker of the GPU hierarchy): on the reference's
    fracture fixture the aggregation hierarchy coarsens by ~5x per level and AMG-PCG needs an order of magnitude fewer
    iterations than Jacobi-PCG (172) for the same heads; the cycle is a symmetric operator."""
    from oracle import amg_oracle
    m = fourfractures
    args = (m["neighbors"], m["areasoverlengths"], m["conductivities"], np.zeros(m["xs"].size), m["dirichletnodes"],
            m["dirichletheads"])
    A = orc.assembleA(*args).toscipy().tocsr()
    b = orc.assembleb(*args)
    H = amg_oracle.Hierarchy(A)
    sizes = H.sizes()
    assert sizes[0] == 2076 and len(sizes) >= 2 and all(c * 3 < f for f, c in zip(sizes[:-1], sizes[1:]))
    x, it, ok = amg_oracle.pcg(A, b, H.apply, math.sqrt(np.finfo(float).eps), 500)
    xj, chj = orc.cg(orc.assembleA(*args), b, Pl="jacobi")
    assert ok and chj.isconverged and it * 5 < chj.iters and chj.iters == 172
    assert np.max(np.abs(x - xj)) <= 1e-6 * np.max(np.abs(xj))
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    assert abs(u @ H.apply(v) - v @ H.apply(u)) <= 1e-10 * abs(u @ H.apply(v))
