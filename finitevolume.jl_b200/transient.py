"""Transient / adjoint drivers mirroring src/transient.jl.

The step controller (step doubling, rejection, growth) is host logic in the reference and
stays host logic here; it is restated from src/transient.jl:78-154 function by function.
What changed is where vectors live: every state vector is a *device-resident slot* of the
System, a backward-Euler solve is one `fvb_step` call, and the controller only ever pulls
one scalar (||onestep - twostep||, :81) across PCIe per attempted step.

Linear algebra of one step.  The reference scales rows, At = D^-1 A with D = Ss*volumes
(scalebyvolume!, :7-22), and solves the non-symmetric (At + I/dt) u+ = D^-1 b + u/dt (:71-73)
by CG with an AMG fallback (:50-58).  We solve the equivalent SPD system
(A + D/dt) u+ = b + D u/dt with Jacobi-PCG; the adjoint step
(At^T + I/dt) g+ = f + g/dt (:193,:203) becomes (A + D/dt) w = f + g/dt, g+ = D w.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .api import DEFAULT_MAXITER, SQRT_EPS, System


class _Pool:
    """Hands out the FVB_NSLOT device vector slots; a DVec returns its slot when dropped."""

    def __init__(self, system: System, reserved=()):
        self.sys = system
        self.free = [s for s in range(_lib.NSLOT) if s not in reserved]

    def new(self):
        if not self.free:
            raise RuntimeError("out of device vector slots")
        return DVec(self, self.free.pop())


class DVec:
    def __init__(self, pool, slot):
        self.pool, self.slot = pool, slot

    def __del__(self):
        try:
            self.pool.free.append(self.slot)
        except Exception:
            pass


class DeviceStepper:
    """backwardeuleronestep! (src/transient.jl:65-76) on device slots."""

    def __init__(self, system: System, getb, adjoint=False, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER):
        self.sys, self.getb, self.adjoint, self.rtol, self.maxiter = system, getb, adjoint, rtol, maxiter
        self.pool = _Pool(system, reserved=(0,))
        self.linear_solves = 0
        self.cg_iterations = 0
        self._b_key = object()

    def _load_b(self, t):
        b = self.getb(t)
        if b is None:  # constant assembled b already resident
            if self._b_key is not None:
                self.sys.vec_load_b(0)
                self._b_key = None
        else:
            self.sys.vec_upload(0, b)
            self._b_key = object()

    def onestep(self, u: DVec, t, dt) -> DVec:
        if dt <= 0:
            raise ValueError("time step must be positive")  # src/transient.jl:68-70
        self._load_b(t)
        out = self.pool.new()
        it, _ = self.sys.step(0, u.slot, dt, out.slot, adjoint=self.adjoint, rtol=self.rtol, maxiter=self.maxiter)
        self.linear_solves += 1
        self.cg_iterations += it
        return out

    def diffnorm(self, a: DVec, b: DVec) -> float:
        return self.sys.vec_diffnorm(a.slot, b.slot)

    def upload(self, v) -> DVec:
        d = self.pool.new()
        self.sys.vec_upload(d.slot, v)
        return d

    def download(self, d: DVec):
        return self.sys.vec_download(d.slot)


class HostStepper:
    """The reference's pluggable `linearsolver(A, rhs, x0)` hook (src/transient.jl:136, used
    with `A \\ b` in test/ode.jl:36) for a user-supplied matrix.  Only the controller runs
    here; the solve is entirely the caller's function -- there is no built-in CPU solver."""

    def __init__(self, A, getb, linearsolver):
        if linearsolver is None:
            raise RuntimeError("a generic (non finite-volume) matrix needs an explicit linearsolver(A, rhs, x0); "
                               "finitevolume.jl_b200 has no CPU solver to fall back to")
        self.A, self.getb, self.linearsolver = A, getb, linearsolver
        self.linear_solves = 0

    def onestep(self, u, t, dt):
        if dt <= 0:
            raise ValueError("time step must be positive")
        rhs = self.getb(t) + u / dt
        n = self.A.shape[0]
        try:
            import scipy.sparse as sp
            shifted = self.A + (sp.identity(n, format="csr") / dt if sp.issparse(self.A) else np.eye(n) / dt)
        except ImportError:  # dense only
            shifted = self.A + np.eye(n) / dt
        self.linear_solves += 1
        return np.asarray(self.linearsolver(shifted, rhs, u)).reshape(-1)

    def diffnorm(self, a, b):
        return float(np.linalg.norm(a - b))

    def upload(self, v):
        return np.array(v, dtype=np.float64)

    def download(self, d):
        return d


# ---- the controller: src/transient.jl:78-154 ----------------------------------------------------------
def backwardeulertwostep(S, u_k, t, dt, atol, onestep=None):
    """src/transient.jl:78-87."""
    if onestep is None:
        onestep = S.onestep(u_k, t, dt)
    twostep1 = S.onestep(u_k, t, 0.5 * dt)
    twostep = S.onestep(twostep1, t + 0.5 * dt, 0.5 * dt)
    err = S.diffnorm(onestep, twostep)
    if err < atol:
        return twostep, dt, err < atol / 4
    return twostep1, 0.5 * dt, False


def adaptivebackwardeulerstep(S, u_k, t, dt, atol, callback):
    """src/transient.jl:89-121."""
    callback(t, dt)
    u_new, laststeptime, increasestepsize = backwardeulertwostep(S, u_k, t, dt, atol)
    if laststeptime < dt:
        laststepfailed = True
        elapsedtime = 0.0
        u_elapsedtime = u_k
        targetdt = laststeptime
        while elapsedtime < dt:
            callback(t, dt)
            if laststepfailed:
                u_new, laststeptime, increasestepsize = backwardeulertwostep(S, u_elapsedtime, t + elapsedtime, targetdt,
                                                                             atol, u_new)
            else:
                u_new, laststeptime, increasestepsize = backwardeulertwostep(S, u_elapsedtime, t + elapsedtime, targetdt,
                                                                             atol)
            if laststeptime == targetdt:
                elapsedtime += laststeptime
                u_elapsedtime = u_new
                if increasestepsize:
                    targetdt = 2 * laststeptime
                laststepfailed = False
            elif laststeptime < targetdt:
                targetdt = laststeptime
                laststepfailed = True
            else:
                raise RuntimeError("Code is broken -- laststeptime should never be greater than targetdt")
            targetdt = min(targetdt, dt - elapsedtime)
    return u_new, laststeptime, increasestepsize


def fixedbackwardeulerstep(S, u_k, t, dt, atol, callback):
    """src/transient.jl:130-134."""
    callback(t, dt)
    return S.onestep(u_k, t, dt), dt, False


def _integrate(S, u0, dt0, t0, tfinal, stepper, atol, callback, keep):
    """src/transient.jl:136-154.  `keep(dvec)` converts an accepted state for storage."""
    u = S.upload(u0)
    us = [keep(u)]
    ts = [t0]
    dt = min(dt0, tfinal - t0)
    while ts[-1] < tfinal:
        u, laststeptime, increasestepsize = stepper(S, u, ts[-1], dt, atol, callback)
        us.append(keep(u))
        ts.append(ts[-1] + dt)
        dt = min(tfinal - ts[-1], 2 * laststeptime) if increasestepsize else min(tfinal - ts[-1], laststeptime)
    return us, ts


def backwardeulerintegrate_generic(u0, A, b, dt0, t0, tfinal, stepper=adaptivebackwardeulerstep, linearsolver=None,
                                   atol=1e-4, callback=lambda t, dt: None):
    """backwardeulerintegrate(u0, A, b|getb, dt0, t0, tfinal; ...) for a caller-supplied matrix
    (src/transient.jl:123-128, :136-154): controller here, solves in the caller's linearsolver."""
    getb = b if callable(b) else (lambda t, _b=np.asarray(b, dtype=np.float64): _b)
    S = HostStepper(A, getb, linearsolver)
    return _integrate(S, u0, dt0, t0, tfinal, stepper, atol, callback, keep=lambda v: np.array(v))


def backwardeulerintegrate(u0, tspan, Ss, volumes, neighbors, areasoverlengths, conductivities, sources,
                           dirichletnodes, dirichletheads, metaindex=None, logtransformconductivity=False, *,
                           dt0=1.0, stepper=adaptivebackwardeulerstep, atol=1e-4, callback=None,
                           getb=None, rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER, device=0, stats=None, controller="device"):
    """Model-level entry, src/transient.jl:156-174 -> (us, ts); every us[i] is a length-N head
    vector (free rows scattered, Dirichlet heads filled in, :172).

    getb(t), when given, must return the UNSCALED right-hand side b(t) on the free rows (the
    reference's getb returns D^-1 b; multiply by Ss*volumes[free] to convert).

    controller="device" (default): the step controller runs inside the library (fvb_integrate; the two stock
    steppers of the reference).  controller="host", or any other `stepper` callable: the Python restatement of
    src/transient.jl:78-154 above drives one fvb_step per solve -- kept as the cross-check of the device one."""
    u0 = np.asarray(u0, np.float64)
    sysm = System(device).assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                                   metaindex, logtransformconductivity)
    try:
        sysm.set_storage(float(Ss), np.asarray(volumes, np.float64))
        freenode = sysm.freenode()
        if controller == "device" and stepper in (adaptivebackwardeulerstep, fixedbackwardeulerstep):
            heads, ts, st = sysm.integrate(u0[freenode], tspan[0], tspan[1], dt0=dt0, atol=atol,
                                           fixed_step=stepper is fixedbackwardeulerstep, rtol=rtol, maxiter=maxiter,
                                           getb=getb, callback=callback, want="heads")
            if stats is not None:
                stats.update(st)
            return [heads[k] for k in range(heads.shape[0])], [float(t) for t in ts]
        if callback is None:
            callback = lambda t, dt: None  # noqa: E731
        S = DeviceStepper(sysm, getb if getb is not None else (lambda t: None), adjoint=False, rtol=rtol, maxiter=maxiter)
        us, ts = _integrate(S, u0[freenode], dt0, tspan[0], tspan[1], stepper, atol, callback,
                            keep=lambda d: sysm.vec_to_nodes(d.slot))
        # the initial entry is the caller's u0 restricted to free nodes + Dirichlet heads (:170-172)
        if stats is not None:
            stats.update(linear_solves=S.linear_solves, cg_iterations=S.cg_iterations, steps=len(ts) - 1)
        return us, ts
    finally:
        sysm.close()


def adjointintegrate(getdgdu, tspan, Ss, volumes, neighbors, areasoverlengths, conductivities, sources,
                     dirichletnodes, dirichletheads, metaindex=None, logtransformconductivity=False, *,
                     dt0=1.0, stepper=adaptivebackwardeulerstep, atol=1e-4, callback=None,
                     rtol=SQRT_EPS, maxiter=DEFAULT_MAXITER, device=0, stats=None, controller="device"):
    """src/transient.jl:188-205 -> (lambdas, ts_lambda): integrates
    dgamma/dt = -(D^-1 A)^T gamma + dg/du(T - t), gamma(0) = 0 and returns it reversed in time,
    lambda(t) = gamma(T - t), on the free rows."""
    sysm = System(device).assemble(neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads,
                                   metaindex, logtransformconductivity)
    try:
        sysm.set_storage(float(Ss), np.asarray(volumes, np.float64))
        nf = sysm.sizes()["nf_local"]
        if controller == "device" and stepper in (adaptivebackwardeulerstep, fixedbackwardeulerstep):
            gam, tsg, st = sysm.integrate(np.zeros(nf), tspan[0], tspan[1], dt0=dt0, atol=atol,
                                          fixed_step=stepper is fixedbackwardeulerstep, adjoint=True, rtol=rtol, maxiter=maxiter,
                                          getb=lambda t: np.asarray(getdgdu(tspan[1] - t), np.float64), callback=callback,
                                          want="free")
            if stats is not None:
                stats.update(st)
            return [gam[k] for k in range(gam.shape[0])][::-1], [tspan[1] - float(t) for t in tsg][::-1]
        if callback is None:
            callback = lambda t, dt: None  # noqa: E731
        S = DeviceStepper(sysm, lambda t: np.asarray(getdgdu(tspan[1] - t), np.float64), adjoint=True, rtol=rtol,
                          maxiter=maxiter)
        gammas, tsg = _integrate(S, np.zeros(nf), dt0, tspan[0], tspan[1], stepper, atol, callback,
                                 keep=lambda d: S.download(d))
        if stats is not None:
            stats.update(linear_solves=S.linear_solves, cg_iterations=S.cg_iterations, steps=len(tsg) - 1)
        return gammas[::-1], [tspan[1] - t for t in tsg][::-1]
    finally:
        sysm.close()


def adjointintegrate_generic(A, getdgdu, tspan, dt0=1.0, **kwargs):
    """adjointintegrate(A, getdgdu, tspan; dt0, kwargs...) for a caller-supplied matrix (src/transient.jl:199-203;
    test/odeadjoint.jl:36): integrates dgamma/dt = -A gamma + dg/du(T - t), gamma(0) = 0 with the generic
    integrator (the solves are the caller's `linearsolver`) and returns (lambdas, ts_lambda), lambda(t) = gamma(T - t).
    As in the reference, A is the matrix of the ADJOINT equation (pass the transpose of the forward one)."""
    gamma0 = np.zeros(A.shape[1])
    gammas, tsg = backwardeulerintegrate_generic(gamma0, A, lambda t: np.asarray(getdgdu(tspan[1] - t), np.float64), dt0,
                                                 tspan[0], tspan[1], **kwargs)
    return gammas[::-1], [tspan[1] - t for t in tsg][::-1]


def gradientintegrate_generic(lambdac, du0dp, dgdp, dfdp, tspan, grids):
    """gradientintegrate(lambdac::Function, du0dp, dgdp::Function, dfdp::Function, tspan) of src/transient.jl:207-216:
    dG/dp = du0dp * lambda(0) + int dg/dp dt + int dfdp(t) * lambda(t) dt.  The reference integrates with adaptive
    QuadGK; here `grids` (the time grids the piecewise-linear u and lambda live on) give a Simpson rule on the merged
    grid, exact for products of two piecewise-linear functions."""
    ts, ws = _simpson_nodes(grids, tspan[0], tspan[1])
    i1 = sum(w * np.asarray(dgdp(t), np.float64) for t, w in zip(ts, ws))
    i2 = sum(w * (np.asarray(dfdp(t), np.float64) @ np.asarray(lambdac(t), np.float64)) for t, w in zip(ts, ws))
    return gradientintegrate(lambdac(tspan[0]), du0dp, i1, i2)


# ======================================================================================================
# Adjoint gradient: src/transientadjointutils.jl + gradientintegrate (src/transient.jl:207-216)
# ======================================================================================================
def getcontinuoussolution(us, ts):
    """src/transient.jl:176-180: piecewise-linear interpolant t -> u(t) of the stored states
    (Interpolations.Gridded(Linear()))."""
    ts = np.asarray(ts, np.float64)
    asc = ts if ts[0] <= ts[-1] else ts[::-1]
    vals = us if ts[0] <= ts[-1] else us[::-1]

    def uc(t):
        j = int(np.clip(np.searchsorted(asc, t, side="right") - 1, 0, len(asc) - 2))
        w = (t - asc[j]) / (asc[j + 1] - asc[j])
        return (1 - w) * np.asarray(vals[j]) + w * np.asarray(vals[j + 1])
    return uc


def _simpson_nodes(grids, t0, t1):
    """Quadrature nodes/weights that integrate products of functions piecewise linear on the given
    grids exactly over [t0, t1] (Simpson on the merged grid; the reference uses adaptive QuadGK)."""
    pts = np.unique(np.concatenate([np.asarray(g, np.float64) for g in grids] + [[t0, t1]]))
    pts = pts[(pts >= t0) & (pts <= t1)]
    w = {}
    for a, b in zip(pts[:-1], pts[1:]):
        h = b - a
        for t, c in ((a, h / 6), (0.5 * (a + b), 4 * h / 6), (b, h / 6)):
            w[t] = w.get(t, 0.0) + c
    ts = sorted(w)
    return ts, [w[t] for t in ts]


def getadjointfunctions(sigma, obsfreenodes, uobs, dirichletnodes, n_nodes, device=0):
    """g and dg/du of src/transientadjointutils.jl:4-21: g(u,t) = sum_i sigma(i,t)^2 (u_i - uobs_i)^2 over the
    observed free rows i (u, uobs: callables t -> length-N node vectors).  Returns (g, dgdu)."""
    from .api import getfreenodes
    freenode, n2f = getfreenodes(n_nodes, dirichletnodes, device=device)
    f2n = np.nonzero(freenode)[0]  # 0-based node of free row (1-based row r -> f2n[r-1])
    nf = int(freenode.sum())

    def g(u, t):
        ue, uo = u(t), uobs(t)
        return float(sum(sigma(i, t) ** 2 * (ue[f2n[i - 1]] - uo[f2n[i - 1]]) ** 2 for i in obsfreenodes))

    def dgdu(u, t):
        ue, uo = u(t), uobs(t)
        out = np.zeros(nf)
        for i in obsfreenodes:
            out[i - 1] = 2 * sigma(i, t) ** 2 * (ue[f2n[i - 1]] - uo[f2n[i - 1]])
        return out
    return g, dgdu


def integrate_g(g, u, grids, tspan):
    """G = int g(u,t) dt (the `G` of getadjointfunctions, :42-54), exact for piecewise-linear u/uobs."""
    ts, ws = _simpson_nodes(grids, tspan[0], tspan[1])
    return float(sum(w * g(u, t) for t, w in zip(ts, ws)))


def integratedfdplambda(us, ts, lambdas, ts_lambda, tspan, Ss, volumes, neighbors, areasoverlengths, conductivities,
                        sources, dirichletnodes, dirichletheads, metaindex=None, logtransformconductivity=False,
                        device=0):
    """int (df/dp)^T lambda dt for p = [conductivities; sources; dirichletheads] (route A of the reference:
    `dfdp(t) * lambda(t)` of src/transientadjointutils.jl:22-32 integrated as in src/transient.jl:207-209), with
    f = D^-1 (b - A u).  us/ts: forward states on all N nodes; lambdas/ts_lambda: adjoint states on the free
    rows (as returned by backwardeulerintegrate / adjointintegrate).  The per-face gather runs on the GPU
    (fvb_gradient_*); u and lambda are piecewise linear in t, so Simpson on the merged grid is exact."""
    from .api import _metaindex_table
    nb = np.asarray(neighbors, np.int64).reshape(-1, 2)
    F = nb.shape[0]
    cond = np.asarray(conductivities, np.float64)
    N, ND = len(sources), len(dirichletheads)
    sysm = System(device).assemble(nb, areasoverlengths, cond, sources, dirichletnodes, dirichletheads, metaindex,
                                   logtransformconductivity)
    try:
        sysm.set_storage(float(Ss), np.asarray(volumes, np.float64))
        freenode = sysm.freenode()
        uc = getcontinuoussolution([np.asarray(u)[freenode] for u in us], ts)
        lc = getcontinuoussolution(lambdas, ts_lambda)
        qt, qw = _simpson_nodes([ts, ts_lambda], tspan[0], tspan[1])
        sysm.gradient_begin(nb)
        for t, w in zip(qt, qw):
            sysm.vec_upload(0, uc(t))
            sysm.vec_upload(1, lc(t))
            sysm.gradient_accumulate(0, 1, w)
        gk, gh, slot, gs = sysm.gradient_end(F)
        meta = _metaindex_table(metaindex, F)
        idx = np.arange(F) if meta is None else meta - 1
        out = np.zeros(cond.size + N + ND)
        out[:cond.size] = np.bincount(idx, weights=gk, minlength=cond.size)            # faces -> conductivity index
        out[cond.size:cond.size + N][freenode] = gs                                     # free rows -> nodes
        has = slot > 0
        out[cond.size + N:] = np.bincount(slot[has] - 1, weights=gh[has], minlength=ND)  # faces -> Dirichlet head
        return out
    finally:
        sysm.close()


def gradientintegrate(lambda0, du0dp, integrated_dgdp, integrateddfdplambda_):
    """src/transient.jl:212-216: dG/dp = du0dp * lambda(0) + int dg/dp dt + int (df/dp)^T lambda dt."""
    first = 0.0 if du0dp is None else du0dp @ np.asarray(lambda0)
    return first + integrated_dgdp + integrateddfdplambda_
