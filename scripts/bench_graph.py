#!/usr/bin/env python
"""Irregular-graph throughput (SURVEY 8d config c3 "plus a synthetic large irregular graph"; workload of
examples/fractures/ex.jl:9-15): the general adjacency/CSR assembly and the CSR SpMV / Jacobi-PCG on a synthetic
discrete-fracture-network-like graph at the sizes of the reference's published timings
(examples/fractures/plotscaling.jl:3-47: homogenous-10m 7 758 411 nodes 143-265 s, pl_alpha_1.6 9 384 128 nodes 174-317 s,
25L_network_x2 18 663 887 nodes 129-242 s, AMG-PCG on one CPU core; the real meshes are not in the repository).

The generator: `nfrac` planar fractures, each an m x m triangulated lattice (nodes joined to their +1, +m, +m+1
neighbours: interior degree 6, like the fixture's mean degree 6.0), numbered fracture after fracture, row-major inside a
fracture; every fracture is tied to `links` random other fractures along a lattice row (intersection traces:
node j of a row of fracture a -- node j of a row of fracture b), which raises degrees up to ~14 as in the fixture;
conductivity = geometric mean of the two fractures' log-normal permeabilities (setupmesh.jl:39), Dirichlet heads 2e6 on
the first lattice row of the first fractures and 1e6 on the last row of the last ones (the fixture's values).
numbering = "natural" | "shuffled" (a random renumbering of all nodes: the worst case for the x[col] gathers) |
"rcm" (reverse Cuthill-McKee of the shuffled graph: what a reordering pass buys back).

Prints one JSON object per numbering: assembly time, CSR SpMV time and GB/s against 12*nnz + 4*(Nf+1) + 16*Nf,
Jacobi-PCG iterations and time.  usage: python scripts/bench_graph.py [--nodes 7758411] [--numbering natural shuffled rcm]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PUBLISHED = {7758411: ("homogenous-10m", 143.3, 265.1), 9384128: ("pl_alpha_1.6", 174.4, 317.2),
             18663887: ("25L_network_x2", 129.2, 241.9)}


def dfn_like_graph(n_nodes_target, nfrac=64, links=3, seed=0):
    """-> dict(neighbors (F,2) int64 1-based with n1 < n2, lexicographically sorted like the fixture; areasoverlengths;
    conductivities; dirichletnodes; dirichletheads; fractureindices; N)."""
    rng = np.random.default_rng(seed)
    m = max(4, int(round(math.sqrt(n_nodes_target / nfrac))))
    per = m * m
    N = per * nfrac
    ii, jj = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    loc = (ii * m + jj).reshape(-1)
    e = []
    for di, dj in ((0, 1), (1, 0), (1, 1)):
        ok = ((ii + di < m) & (jj + dj < m)).reshape(-1)
        e.append(np.stack([loc[ok], loc[ok] + di * m + dj], 1))
    lattice = np.concatenate(e)                       # (~3 m^2, 2), local ids
    pairs = [lattice + f * per for f in range(nfrac)]
    # intersection traces
    for a in range(nfrac):
        for b in rng.choice(nfrac, size=links, replace=False):
            if a == b:
                continue
            ra, rb = rng.integers(0, m, 2)
            na = a * per + ra * m + np.arange(m)
            nb_ = b * per + rb * m + np.arange(m)
            pairs.append(np.stack([np.minimum(na, nb_), np.maximum(na, nb_)], 1))
    nb = np.concatenate(pairs).astype(np.int64)
    frac_perm = np.exp(rng.normal(math.log(1e-12), 1.0, nfrac))
    fidx = np.repeat(np.arange(nfrac), per)
    k = np.sqrt(frac_perm[fidx[nb[:, 0]]] * frac_perm[fidx[nb[:, 1]]])
    aol = rng.uniform(2.8e-9, 3.5e-5, nb.shape[0])
    nin = max(1, nfrac // 16)
    dn = np.concatenate([f * per + np.arange(m) for f in range(nin)] +
                        [f * per + (m - 1) * m + np.arange(m) for f in range(nfrac - nin, nfrac)])
    dh = np.concatenate([np.full(nin * m, 2e6), np.full(nin * m, 1e6)])
    return dict(N=N, nb0=nb, aol=aol, k=k, dn0=dn, dh=dh, fractureindices=fidx + 1, m=m, nfrac=nfrac)


def renumber(G, how, seed=1):
    """Apply a node numbering and return 1-based, (n1 < n2)-oriented, lexicographically sorted face arrays."""
    N = G["N"]
    if how == "natural":
        new = np.arange(N)
    else:
        new = np.random.default_rng(seed).permutation(N)
        if how == "rcm":
            import scipy.sparse as sp
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            a, b = new[G["nb0"][:, 0]], new[G["nb0"][:, 1]]
            A = sp.coo_matrix((np.ones(a.size, np.int8), (a, b)), shape=(N, N)).tocsr()
            A = A + A.T
            perm = reverse_cuthill_mckee(A, symmetric_mode=True)   # perm[i] = old (shuffled) id placed at i
            inv = np.empty(N, np.int64)
            inv[perm] = np.arange(N)
            new = inv[new]
    a, b = new[G["nb0"][:, 0]], new[G["nb0"][:, 1]]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    order = np.lexsort((hi, lo))
    nb = np.stack([lo[order], hi[order]], 1).astype(np.int64) + 1
    return nb, G["aol"][order], G["k"][order], np.sort(new[G["dn0"]]) + 1, G["dh"][np.argsort(new[G["dn0"]], kind="stable")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=7758411)
    ap.add_argument("--nfrac", type=int, default=64)
    ap.add_argument("--numbering", nargs="+", default=["natural", "shuffled", "rcm"])
    ap.add_argument("--rtol", type=float, default=math.sqrt(np.finfo(float).eps))
    ap.add_argument("--maxiter", type=int, default=200000)
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as g
    fv = g.load_package()
    t0 = time.perf_counter()
    G = dfn_like_graph(args.nodes, args.nfrac)
    t_gen = time.perf_counter() - t0
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    pub = min(PUBLISHED, key=lambda n: abs(n - G["N"]))
    for how in args.numbering:
        t0 = time.perf_counter()
        nb, aol, k, dn, dh = renumber(G, how)
        t_ren = time.perf_counter() - t0
        src = np.zeros(G["N"])
        s = fv.System(0)
        s.assemble(nb, aol, k, src, dn, dh)          # warm-up (allocations)
        t0 = time.perf_counter()
        s.assemble(nb, aol, k, src, dn, dh)
        t_asm_wall = time.perf_counter() - t0
        tm = s.timings()
        sz = s.sizes()
        nf, nnz = sz["nf_local"], sz["nnz_local"]
        deg = np.bincount(np.concatenate([nb[:, 0], nb[:, 1]]), minlength=G["N"] + 1)[1:]
        ms = s.time_spmv(warmup=3, reps=20)
        alg = 12 * nnz + 4 * (nf + 1) + 16 * nf
        out = {"graph": f"synthetic DFN-like: {G['nfrac']} fractures of {G['m']}x{G['m']} triangulated lattices + intersection traces",
               "numbering": how, "nodes": G["N"], "faces": int(nb.shape[0]), "free_rows": nf, "nnz": nnz,
               "degree_min_mean_max": [int(deg.min()), float(deg.mean()), int(deg.max())],
               "bandwidth_of_numbering": int(np.max(nb[:, 1] - nb[:, 0])),
               "format": s.spmv_format()[0], "assembly": s.assembly(),
               "assemble_device_ms": tm["assemble_ms"], "assemble_h2d_ms": tm["h2d_ms"], "assemble_call_s": t_asm_wall,
               "assembly_algorithmic_gb": (32 * nb.shape[0] + 12 * nnz + 12 * nf) / 1e9,
               "assembly_gbs": (32 * nb.shape[0] + 12 * nnz + 12 * nf) / 1e9 / (tm["assemble_ms"] * 1e-3),
               "spmv_ms": ms, "spmv_algorithmic_bytes": alg, "spmv_gbs": alg / (ms * 1e-3) / 1e9,
               "spmv_frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
               "host_generate_s": t_gen, "host_renumber_s": t_ren}
        if not args.no_solve:
            t0 = time.perf_counter()
            head, _, ch = s.solve(rtol=args.rtol, maxiter=args.maxiter)
            out.update(solve_wall_s=time.perf_counter() - t0, solve_device_ms=s.timings()["solve_ms"], pcg_iterations=ch.iters,
                       converged=ch.isconverged, head_range=[float(head.min()), float(head.max())],
                       time_to_solution_s=t_asm_wall + (time.perf_counter() - t0),
                       published_reference={"mesh": PUBLISHED[pub][0], "nodes": pub,
                                            "solvediffusion_s_fastest_slowest_host": list(PUBLISHED[pub][1:]),
                                            "source": "examples/fractures/plotscaling.jl:3-47 (AMG-PCG, one CPU process, other "
                                                      "hardware, the real dfnWorks mesh -- context, not a like-for-like ratio)"})
            # the preconditioner class the reference actually uses on these meshes: algebraic multigrid on the CSR rows
            try:
                t0 = time.perf_counter()
                s.set_preconditioner("mg")
                t_setup = time.perf_counter() - t0
                kind, nlev = s.preconditioner()
                t0 = time.perf_counter()
                head_a, _, cha = s.solve(rtol=args.rtol, maxiter=args.maxiter)
                out["amg"] = {"active": kind, "levels": nlev, "setup_s": t_setup, "solve_wall_s": time.perf_counter() - t0,
                              "solve_device_ms": s.timings()["solve_ms"], "pcg_iterations": cha.iters, "converged": cha.isconverged,
                              "max_rel_head_difference_vs_jacobi": float(np.max(np.abs(head_a - head)) / np.max(np.abs(head))),
                              "time_to_solution_s": t_asm_wall + t_setup + (time.perf_counter() - t0)}
            except Exception as e:
                out["amg"] = {"error": repr(e)}
        print(json.dumps(out), flush=True)
        s.close()


if __name__ == "__main__":
    sys.exit(main())
