"""CPU restatement (numpy/scipy) of the aggregation-AMG preconditioner of finitevolume.jl_b200/csrc/amg.cuh.

TEST INFRASTRUCTURE ONLY (see oracle/fv_oracle.c): tests compare the GPU hierarchy (level sizes) and one V-cycle
application against this restatement; nothing under finitevolume.jl_b200/ imports it.  The reference's own
preconditioner is AlgebraicMultigrid.ruge_stuben (src/FiniteVolume.jl:160), an un-vendored upstream package; this is
NOT a restatement of Ruge-Stueben but of the aggregation scheme the B200 build uses in its place (same class:
Galerkin hierarchy + V-cycle inside CG), so it pins our own kernels, not the reference's iteration counts.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

MAX_LEVELS, COARSEST, COARSE_SWEEPS, MAX_ROW, ROUNDS, STALL = 24, 512, 40, 128, 6, 0.9


def edge_hash(a, b):
    lo = np.minimum(a, b).astype(np.uint32)
    hi = np.maximum(a, b).astype(np.uint32)
    with np.errstate(over="ignore"):
        h = (lo * np.uint32(0x9E3779B1)) ^ ((hi + np.uint32(0x7F4A7C15)) * np.uint32(0x85EBCA77))
        h ^= h >> np.uint32(15); h *= np.uint32(0x2C1B3C6D); h ^= h >> np.uint32(12); h *= np.uint32(0x297A2D39); h ^= h >> np.uint32(15)
    return h


def pairwise(A):
    """Handshake matching (k_amg_propose / k_amg_accept), ROUNDS rounds -> agg (row -> coarse row), nc."""
    A = A.tocsr()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    cols, w = A.indices, -A.data
    hk = edge_hash(rows, cols)
    match = np.full(n, -1, np.int64)
    for _ in range(ROUNDS):
        ok = (rows != cols) & (w > 0) & (match[rows] < 0) & (match[cols] < 0)
        r, c, ww, hh = rows[ok], cols[ok], w[ok], hk[ok]
        # per row: maximum by (weight, hash); entries are in row order, the first maximum in that order wins ties
        order = np.lexsort((-np.arange(r.size), hh, ww, r))  # ascending: last of each row group is the best
        r_s = r[order]
        last = np.flatnonzero(np.r_[r_s[1:] != r_s[:-1], True]) if r_s.size else np.empty(0, np.int64)
        prop = np.full(n, -1, np.int64)
        prop[r_s[last]] = c[order][last]
        i = np.flatnonzero((match < 0) & (prop >= 0))
        mutual = i[prop[prop[i]] == i]
        match[mutual] = prop[mutual]
    # leftovers join the pair of their strongest paired neighbour (k_amg_attach)
    ok = (rows != cols) & (w > 0) & (match[rows] < 0) & (match[cols] >= 0)
    r, c, ww, hh = rows[ok], cols[ok], w[ok], hk[ok]
    order = np.lexsort((-np.arange(r.size), hh, ww, r))
    r_s = r[order]
    last = np.flatnonzero(np.r_[r_s[1:] != r_s[:-1], True]) if r_s.size else np.empty(0, np.int64)
    attach = np.full(n, -1, np.int64)
    attach[r_s[last]] = c[order][last]
    idx = np.arange(n)
    a_safe = np.where(attach >= 0, attach, 0)
    rootid = np.where(match >= 0, np.minimum(idx, match), np.where(attach >= 0, np.minimum(a_safe, match[a_safe]), idx))
    isroot = rootid == idx
    cid = np.cumsum(isroot) - 1
    return cid[rootid], int(isroot.sum())


def galerkin(A, agg, nc):
    n = A.shape[0]
    P = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
    Ac = (P.T @ A @ P).tocsr()
    Ac.sort_indices()
    return Ac


class Hierarchy:
    def __init__(self, A, nu=2, omega=0.8, oc=1.5):
        self.nu, self.omega, self.oc = nu, omega, oc
        self.levels = []
        A = A.tocsr()
        while True:
            n = A.shape[0]
            lev = dict(A=A, dinv=1.0 / A.diagonal(), agg=None)
            self.levels.append(lev)
            if n <= COARSEST or len(self.levels) >= MAX_LEVELS:
                break
            a1, n1 = pairwise(A)
            A1 = galerkin(A, a1, n1)
            a2, n2 = pairwise(A1)
            agg = a2[a1]
            if n2 > int(STALL * n) or n2 < 1:
                break
            Ac = galerkin(A, agg, n2)
            if np.max(np.diff(Ac.indptr)) > MAX_ROW:
                break
            lev["agg"] = agg
            lev["nc"] = n2
            A = Ac

    def sizes(self):
        return [L["A"].shape[0] for L in self.levels]

    def _smooth(self, L, r, x):
        return x + self.omega * L["dinv"] * (r - L["A"] @ x)

    def apply(self, r):
        """z = M^-1 r: one V(nu,nu) cycle from a zero initial guess (csrc/amg.cuh, fvb200.cu: amg_vcycle)."""
        return self._cycle(0, np.asarray(r, np.float64))

    def _cycle(self, l, r):
        L = self.levels[l]
        if l == len(self.levels) - 1 or L["agg"] is None:
            x = self.omega * L["dinv"] * r
            for _ in range(1, COARSE_SWEEPS):
                x = self._smooth(L, r, x)
            return x
        x = self.omega * L["dinv"] * r
        for _ in range(1, self.nu):
            x = self._smooth(L, r, x)
        rc = np.bincount(L["agg"], weights=r - L["A"] @ x, minlength=L["nc"])
        x = x + self.oc * self._cycle(l + 1, rc)[L["agg"]]
        for _ in range(self.nu):
            x = self._smooth(L, r, x)
        return x


def pcg(A, b, M, tol, maxiter):
    x = np.zeros_like(b)
    r = b.copy()
    res0 = np.linalg.norm(r)
    z = M(r)
    p = z.copy()
    rz = r @ z
    for it in range(1, maxiter + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= tol * res0:
            return x, it, True
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxiter, False
