"""Child process of tests/test_gpu_multi.py for the single-process multi-GPU front end (fvb_multi_*): runs one case
and prints "inproc ok".  usage: python tests/multi_inproc_worker.py regular|irregular <ndev> <n1,n2,n3|->"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from oracle import fv_oracle as orc  # noqa: E402

fv = g.load_package()
orc.build()


def box(ns, sigma=1.0):
    _, nb, aol, vol = fv.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
    N = int(np.prod(ns))
    lnk = math.log(1e-5) + sigma * np.random.default_rng(0).standard_normal(N)
    kf = fv.nodehycos2neighborhycos(nb, lnk, True)
    plane = ns[1] * ns[2]
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.ones(plane), np.zeros(plane)])
    src = 1e-7 * np.random.default_rng(1).standard_normal(N)
    src[dn - 1] = 0
    return nb, aol, kf, src, dn, dh, lnk


def regular(ndev, ns):
    nb, aol, kf, src, dn, dh, lnk = box(ns)
    ms = fv.MultiSystem(list(range(ndev)))
    ms.assemble(nb, aol, kf, src, dn, dh, None, True)
    sz = ms.sizes()
    assert sz["node_ranges"][0][0] == 1 and sz["node_ranges"][-1][1] == src.size
    Ao = orc.assembleA(nb, aol, kf, src, dn, dh, None, True)
    bo = orc.assembleb(nb, aol, kf, src, dn, dh, None, True)
    p, i, v = ms.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.allclose(v, Ao.nzval, rtol=1e-14, atol=0)
    assert np.allclose(ms.b(), bo, rtol=1e-14, atol=0)
    print("assembled", sz, flush=True)
    head, x, ch = ms.solve(rtol=1e-12, want_x=True)
    ho, cho, *_ = orc.solvediffusion(nb, aol, kf, src, dn, dh, maxiter=50000, tol=1e-12, logtransformconductivity=True)
    assert ch.isconverged and np.max(np.abs(head - ho)) <= 1e-8 * np.max(np.abs(ho)), np.max(np.abs(head - ho))
    assert abs(ch.iters - cho.iters) <= 3, (ch.iters, cho.iters)
    fn = ms.freenode()
    assert np.array_equal(head[fn], x) and np.array_equal(head[~fn], dh)
    kinds = [ms.device_system(r).assembly() for r in range(ndev) if ms.device_system(r).sizes()["nf_local"] >= 2]
    if ns[0] >= ndev + 2:  # (with a rank that owns no free row the front end falls back to filtered node ranges)
        assert kinds and all(k == "box" for k in kinds), kinds
    print("solved", ch.iters, kinds, flush=True)
    h2, ch2, A2, b2, fn2 = fv.solvediffusion(nb, aol, kf, src, dn, dh, rtol=1e-12, logtransformconductivity=True,
                                             devices=list(range(ndev)))
    assert np.array_equal(h2, head) and ch2.iters == ch.iters and np.array_equal(A2.rowval, Ao.rowval)
    print("solvediffusion(devices=...) ok", flush=True)
    if ns[0] >= ndev + 2:
        mi = fv.MultiSystem(list(range(ndev)))
        mi.assemble_regulargrid([0, 0, 0], [n - 1 for n in ns], ns, lnk, src, dn, dh)
        hi_, _, chi = mi.solve(rtol=1e-12)
        assert chi.isconverged and np.max(np.abs(hi_ - ho)) <= 1e-8 * np.max(np.abs(ho))
        print("implicit ok", flush=True)
        mg = fv.MultiSystem(list(range(ndev)))
        mg.set_preconditioner("mg")
        mg.assemble(nb, aol, kf, src, dn, dh, None, True)
        hm, _, chm = mg.solve(rtol=1e-12)
        assert chm.isconverged and np.max(np.abs(hm - ho)) <= 1e-8 * np.max(np.abs(ho))
        print("multigrid ok", chm.iters, mg.device_system(0).preconditioner(), flush=True)


def irregular(ndev):
    ff = dict(np.load(os.path.join(ROOT, "tests", "golden", "fourfractures.npz")))
    args = (ff["neighbors"], ff["areasoverlengths"], ff["conductivities"], np.zeros(ff["xs"].size), ff["dirichletnodes"],
            ff["dirichletheads"])
    ms = fv.MultiSystem(list(range(ndev))).assemble(*args)
    Ao = orc.assembleA(*args)
    p, i, v = ms.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.array_equal(v, Ao.nzval)
    head, _, ch = ms.solve(rtol=1e-12)
    ho, cho, *_ = orc.solvediffusion(*args, maxiter=20000, tol=1e-12)
    assert ch.isconverged and np.max(np.abs(head - ho)) <= 1e-8 * np.max(np.abs(ho))
    print("irregular solved", ch.iters, cho.iters, flush=True)


if __name__ == "__main__":
    case, ndev, arg = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    if case == "regular":
        regular(ndev, [int(v) for v in arg.split(",")])
    else:
        irregular(ndev)
    print("inproc ok", flush=True)
