#!/usr/bin/env python
"""The reference's fracture-network example (examples/fractures/ex.jl:6-17) on the B200 path:

    xs, ys, zs, neighbors, areasoverlengths, fractureindices, dirichletnodes, dirichletheads, conductivities =
        JLD.load(joinpath(meshdir, "mesh.jld"), "xs", ..., "conductivities")
    sources = zeros(length(xs))
    h, ch, A, b, freenode = FiniteVolume.solvediffusion(neighbors, areasoverlengths, conductivities, sources,
                                                        dirichletnodes, dirichletheads)

followed by the re-solve loop of examples/fractures/ex_comparison.jl:34-42 (the conductivity of one fracture
changes, the mesh does not), which here is a values-only re-assembly on the retained structure.

    python examples/fractures.py /path/to/examples/fractures/fourfractures [--resolves 3]

The mesh is irregular, so the solve runs on the general CSR kernels.  Needs a B200 and the built library
(`python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

NAMES = ("xs", "ys", "zs", "neighbors", "areasoverlengths", "fractureindices", "dirichletnodes", "dirichletheads",
         "conductivities")


def load_problem(fv, meshdir):
    """examples/fractures/ex.jl:9-10 -> dict of arrays (neighbors as (F, 2) int64, 1-based)."""
    vals = fv.jld.load(os.path.join(meshdir, "mesh.jld"), *NAMES)
    p = dict(zip(NAMES, vals))
    p["sources"] = np.zeros(p["xs"].size)
    return p


def fracture_of_face(p):
    """connection2fracture of ex.jl:12: a face belongs to the fracture of its first node."""
    return p["fractureindices"][p["neighbors"][:, 0] - 1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("meshdir")
    ap.add_argument("--resolves", type=int, default=3)
    ap.add_argument("--rtol", type=float, default=None)
    args = ap.parse_args()
    fv = g.load_package()
    t0 = time.perf_counter()
    p = load_problem(fv, args.meshdir)
    print(f"mesh: {p['xs'].size} nodes, {p['neighbors'].shape[0]} faces, {p['dirichletnodes'].size} Dirichlet nodes "
          f"({time.perf_counter() - t0:.3f} s to read)")
    rtol = fv.SQRT_EPS if args.rtol is None else args.rtol
    s = fv.System().assemble(p["neighbors"], p["areasoverlengths"], p["conductivities"], p["sources"], p["dirichletnodes"],
                             p["dirichletheads"])
    t0 = time.perf_counter()
    h, x, ch = s.solve(rtol=rtol, want_x=True)
    b = s.b()
    print(f"solvediffusion: {ch.iters} iterations, converged={ch.isconverged}, {time.perf_counter() - t0:.4f} s, "
          f"format={s.spmv_format()[0]}, |A h - b| = {np.linalg.norm(s.spmv(x) - b):.3e}, "
          f"heads in [{h.min():.6g}, {h.max():.6g}]")
    # ex_comparison.jl:34-42: raise the conductivity inside one fracture, keep the mesh
    rng = np.random.default_rng(0)
    f2f = fracture_of_face(p)
    both = p["fractureindices"][p["neighbors"][:, 1] - 1] == f2f
    for f in rng.choice(np.unique(p["fractureindices"]), size=args.resolves):
        k = p["conductivities"].copy()
        k[(f2f == f) & both] *= 1.5
        t0 = time.perf_counter()
        s.update_values(k)
        h2, _, ch2 = s.solve(rtol=rtol)
        print(f"fracture {int(f)} x1.5: values-only re-assembly + solve {time.perf_counter() - t0:.4f} s, {ch2.iters} iterations, "
              f"max head change {np.max(np.abs(h2 - h)):.4g}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
