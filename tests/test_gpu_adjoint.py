"""Transient adjoint gradient (BASELINE config 4: "... plus transientadjoint gradient"): the per-face gather
kernel + exact time quadrature against central finite differences of the objective, as the reference's own
tests do (test/onenodeadjoint.jl:67-75 rtol 1e-2, test/theisadjoint.jl:74-84 rtol 1e-3)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def objective_and_gradient(fv, prob, p, uobs, obsfree, sigma, want_gradient=True):
    """G(p) = int sum_obs sigma^2 (u - uobs)^2 dt through the GPU forward solve; dG/dp through the GPU adjoint."""
    nk, N = prob["nk"], prob["N"]
    k, src, dh = p[:nk], p[nk:nk + N], p[nk + N:]
    kw = dict(atol=prob["atol"], dt0=prob["dt0"], rtol=1e-13)
    us, ts = fv.backwardeulerintegrate(prob["u0"], prob["tspan"], prob["Ss"], prob["vol"], prob["nb"], prob["aol"], k,
                                       src, prob["dn"], dh, prob.get("meta"), prob["logk"], **kw)
    uc = fv.getcontinuoussolution(us, ts)
    g, dgdu = fv.getadjointfunctions(sigma, obsfree, uobs["uc"], prob["dn"], N)
    G = fv.integrate_g(g, uc, [ts, uobs["ts"]], prob["tspan"])
    if not want_gradient:
        return G, None
    lam, tl = fv.adjointintegrate(lambda t: dgdu(uc, t), prob["tspan"], prob["Ss"], prob["vol"], prob["nb"],
                                  prob["aol"], k, src, prob["dn"], dh, prob.get("meta"), prob["logk"], **kw)
    grad = fv.integratedfdplambda(us, ts, lam, tl, prob["tspan"], prob["Ss"], prob["vol"], prob["nb"], prob["aol"], k,
                                  src, prob["dn"], dh, prob.get("meta"), prob["logk"])
    return G, fv.gradientintegrate(lam[0], None, 0.0, grad)


def check_fd(fv, prob, p0, uobs, obsfree, sigma, indices, deltap, rtol):
    G0, grad = objective_and_gradient(fv, prob, p0, uobs, obsfree, sigma)
    assert grad.shape == p0.shape
    scale = np.max(np.abs(grad))
    for i in indices:
        pp, pm = p0.copy(), p0.copy()
        pp[i] += deltap
        pm[i] -= deltap
        fd = (objective_and_gradient(fv, prob, pp, uobs, obsfree, sigma, False)[0]
              - objective_and_gradient(fv, prob, pm, uobs, obsfree, sigma, False)[0]) / (2 * deltap)
        assert math.isclose(fd, grad[i], rel_tol=rtol, abs_tol=1e-6 * scale), (i, fd, grad[i])
    return grad


def test_onenode_gradient(fv):
    """test/onenodeadjoint.jl:13-28,46-75: two nodes, log K; parameters 1 (K), 3 (source at node 2), 4 (Dirichlet head)."""
    prob = dict(Ss=1.0, vol=[1.0, 1.0], nb=[(1, 2)], aol=[1.0], dn=[1], u0=[0.0, 0.0], tspan=(0.0, 1.0), atol=1e-7,
                dt0=1e-3, logk=True, nk=1, N=2)  # (the reference integrates with atol=1e-8; 1e-7 keeps the test short)
    sigma = lambda i, t: 0.01  # noqa: E731
    us, ts = fv.backwardeulerintegrate(prob["u0"], prob["tspan"], 1.0, prob["vol"], prob["nb"], prob["aol"], [0.0],
                                       [0.0, 1.0], [1], [0.0], None, True, atol=1e-7, dt0=1e-3, rtol=1e-13)
    uobs = dict(uc=fv.getcontinuoussolution(us, ts), ts=ts)
    p0 = np.array([1.0, 0.0, 1.0, 0.0])
    # continuous adjoint vs FD of the discretised objective: they meet as atol -> 0 (1e-2 at the reference's 1e-8)
    grad = check_fd(fv, prob, p0, uobs, [1], sigma, [0, 2, 3], 1e-4, 3e-2)
    assert grad[1] == 0.0  # a source on the Dirichlet node is not a parameter of the free system


@pytest.mark.parametrize("logk", [True, False])
def test_box_gradient_all_parameter_blocks(fv, logk):
    """Heterogeneous 5x4x3 box with non-uniform volumes, sources and three Dirichlet planes' worth of heads:
    conductivity, source and Dirichlet-head entries of the gradient against central differences; a metaindex
    table shares conductivities between faces."""
    ns = [5, 4, 3]
    _, nb, aol, vol = fv.regulargrid([0, 0, 0], [4, 3, 2], ns, want_coords=False)
    N = int(np.prod(ns))
    rng = np.random.default_rng(11)
    nk = 9
    meta = rng.integers(1, nk + 1, size=nb.shape[0])
    k = rng.standard_normal(nk) * 0.5 if logk else np.exp(rng.standard_normal(nk) * 0.5)
    plane = ns[1] * ns[2]
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.full(plane, 1.0), np.full(plane, 0.2)]) + 0.05 * rng.random(2 * plane)
    src = 0.3 * rng.standard_normal(N)
    src[dn - 1] = 0
    prob = dict(Ss=0.7, vol=vol, nb=nb, aol=aol, dn=dn, u0=np.full(N, 0.5), tspan=(0.0, 0.8), atol=1e-6, dt0=2e-2,
                logk=logk, nk=nk, N=N, meta=meta)
    # observations: the solution for perturbed parameters
    k_true = k + 0.3 * rng.standard_normal(nk) * (1 if logk else 0.2 * k)
    us, ts = fv.backwardeulerintegrate(prob["u0"], prob["tspan"], prob["Ss"], vol, nb, aol, k_true, src, dn, dh, meta, logk,
                                       atol=1e-6, dt0=2e-2, rtol=1e-13)
    uobs = dict(uc=fv.getcontinuoussolution(us, ts), ts=ts)
    freenode, n2f = fv.getfreenodes(N, dn)
    obsfree = [int(n2f[n]) for n in (plane + 2, 2 * plane + 5, 3 * plane + 1)]
    sigma = lambda i, t: 1.0 + 0.1 * i / N  # noqa: E731
    p0 = np.concatenate([k, src, dh])
    free_nodes = np.nonzero(freenode)[0]
    idx = [0, nk - 1, nk + free_nodes[7], nk + N + 1, nk + N + plane + 3]
    grad = check_fd(fv, prob, p0, uobs, obsfree, sigma, idx, 1e-4, 1e-2)
    assert np.all(grad[nk:nk + N][~freenode] == 0.0) and np.any(grad[:nk] != 0) and np.any(grad[nk + N:] != 0)
