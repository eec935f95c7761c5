"""One small pass through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck; one tool per
gpurun call): closed-form and general assembly, the three SpMV kernels (TMA pipelines with interior AND edge tiles),
scaled and unscaled Jacobi-PCG, both multigrids, the lazily built CSR, values-only update, the cooperative transient
attempt kernel and the launch-per-phase stepper, the adjoint gradient gather.  Every result is checked against the
CPU oracle so that a sanitizer run that "passes" on garbage is impossible.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import __graft_entry__ as g  # noqa: E402
from oracle import fv_oracle as orc  # noqa: E402

fv = g.load_package()
ns = [24, 16, 16]
_, nb, aol, vol = fv.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
N = int(np.prod(ns))
lnk = math.log(1e-5) + np.random.default_rng(0).standard_normal(N)
kf = fv.nodehycos2neighborhycos(nb, lnk, True)
plane = ns[1] * ns[2]
dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
dh = np.concatenate([np.ones(plane), np.zeros(plane)])
src = np.zeros(N)
ho, *_ = orc.solvediffusion(nb, aol, kf, src, dn, dh, maxiter=20000, tol=1e-11, logtransformconductivity=True)


def check(head, what):
    err = np.max(np.abs(head - ho))
    assert err <= 1e-7, (what, err)
    print(f"ok {what}: max|dh| = {err:.1e}", flush=True)


s = fv.System().assemble(nb, aol, kf, src, dn, dh, None, True)           # closed-form assembly
assert s.assembly() == "box"
for fmt, name in ((3, "dia_tma"), (2, "dia"), (1, "csr")):                # three SpMV kernels; 1 builds the lazy CSR
    s.set_spmv_format(fmt)
    h, _, ch = s.solve(rtol=1e-11)
    check(h, f"box + {name} (scaled={s.pcg_scaling()})")
s.set_spmv_format(3)
s.set_pcg_scaling(1)
h, _, _ = s.solve(rtol=1e-11)
check(h, "unscaled recurrence")
s.set_pcg_scaling(0)
s.set_preconditioner("mg")
h, _, ch = s.solve(rtol=1e-11)
check(h, f"geometric multigrid ({ch.iters} its)")
s.set_preconditioner("jacobi")
s.update_values(kf)
h, _, _ = s.solve(rtol=1e-11)
check(h, "values-only update")
si = fv.System().assemble_regulargrid([0, 0, 0], [n - 1 for n in ns], ns, lnk, None, dn, dh)
h, _, _ = si.solve(rtol=1e-11)
check(h, "grid-implicit assembly")
sg = fv.System()
sg.set_assembly(1)
sg.set_preconditioner("mg")
sg.set_spmv_format(1)
sg.assemble(nb, aol, kf, src, dn, dh, None, True)                         # general assembly, CSR forced => algebraic MG
h, _, ch = sg.solve(rtol=1e-11)
check(h, f"general assembly + CSR + {sg.preconditioner()[0]} ({ch.iters} its)")
# transient: cooperative attempt kernel, then the launch-per-phase stepper; adjoint gradient gather
u0 = np.full(N, 0.5)
for coop in ("0", "1"):
    os.environ["FVB_COOP_OFF"] = "" if coop == "1" else "1"
    if coop == "1":
        del os.environ["FVB_COOP_OFF"]
    st = {}
    us, ts = fv.backwardeulerintegrate(u0, (0.0, 2000.0), 0.1, vol, nb, aol, np.exp(kf), src, dn, dh, atol=1e-4, dt0=10.0,
                                       rtol=1e-10, stats=st)
    print(f"ok transient (coop={coop}): {st}", flush=True)
    if coop == "0":
        ref = us[-1]
    else:
        assert np.max(np.abs(us[-1] - ref)) <= 1e-8
s2 = fv.System().assemble(nb, aol, kf, src, dn, dh, None, True)
s2.set_storage(0.1, vol)
s2.gradient_begin(nb)
s2.vec_upload(1, np.linspace(0, 1, s2.sizes()["nf_local"]))
s2.vec_upload(2, np.ones(s2.sizes()["nf_local"]))
s2.gradient_accumulate(1, 2, 0.5)
gk, gh, sl, gs = s2.gradient_end(nb.shape[0])
assert np.all(np.isfinite(gk)) and np.all(np.isfinite(gs))
print("ok gradient gather", flush=True)
print("sanitize_case ok", flush=True)
