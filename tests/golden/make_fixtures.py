"""Regenerates the committed fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference); the GPU box never runs it.

fourfractures.npz  <- examples/fractures/fourfractures/mesh.jld and pflotran_solution.jld
    JLD 0.1.1 = HDF5 with a 512-byte user block; datasets are contiguous little-endian,
    read here at their absolute file offsets (no h5py in this image; offsets recorded in
    SURVEY.md App. C and re-checked below through the invariants the survey lists).
"""
import hashlib
import os
import sys

import numpy as np

REF = os.environ.get("FV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def rd(buf, off, dtype, count):
    return np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()


def main():
    mesh = open(os.path.join(REF, "examples/fractures/fourfractures/mesh.jld"), "rb").read()
    assert len(mesh) == 280040 and hashlib.sha256(mesh).hexdigest().startswith("2dcddfbb5998e606")
    N, F, ND = 2106, 6314, 30
    out = dict(
        xs=rd(mesh, 4456, "<f8", N), ys=rd(mesh, 23352, "<f8", N), zs=rd(mesh, 40200, "<f8", N),
        neighbors=rd(mesh, 59096, "<i8", 2 * F).reshape(F, 2),
        areasoverlengths=rd(mesh, 160120, "<f8", F),
        fractureindices=rd(mesh, 210632, "<i8", N),
        conductivities=rd(mesh, 229528, "<f8", F),
        dirichletnodes=rd(mesh, 57964, "<i8", ND),
        dirichletheads=rd(mesh, 228100, "<f8", ND),
    )
    nb = out["neighbors"]
    assert nb.min() == 1 and nb.max() == N and (nb[:, 0] < nb[:, 1]).all()
    assert abs(out["areasoverlengths"].sum() - 0.04095143598578653) < 1e-15
    assert (out["conductivities"] == 1e-12).all()
    assert set(out["fractureindices"]) == {1, 2, 3, 4}
    assert list(out["dirichletnodes"]) == list(range(1, 11)) + list(range(1676, 1686)) + list(range(2097, 2107))
    assert (out["dirichletheads"][:10] == 2e6).all() and (out["dirichletheads"][10:] == 1e6).all()
    pf = open(os.path.join(REF, "examples/fractures/fourfractures/pflotran_solution.jld"), "rb").read()
    out["pflotran_h"] = rd(pf, 4456, "<f8", N)
    assert 1e6 - 1 <= out["pflotran_h"].min() and out["pflotran_h"].max() <= 2e6 + 1
    np.savez_compressed(os.path.join(HERE, "fourfractures.npz"), **out)
    print("wrote fourfractures.npz")


if __name__ == "__main__":
    sys.exit(main())
