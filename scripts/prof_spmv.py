"""Small driver for ncu captures: assemble an n^3 lognormal box, run a few SpMV launches and a
short (maxiter-capped) PCG so every hot kernel appears a handful of times.
usage: python scripts/prof_spmv.py [n] [pcg_iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from bench import problem_inputs, spmv_bytes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fv = g.load_package()
P = problem_inputs(fv, n, 1.0)
s = fv.System(0)
t0 = time.perf_counter()
s.assemble(P["nb"], P["aol"], P["kf"], P["src"], P["dn"], P["dh"], None, True)
t1 = time.perf_counter()
sz = s.sizes()
fmt = s.spmv_format()
s.set_spmv_format(1)
ms = s.time_spmv(warmup=2, reps=5)
s.set_spmv_format(0)
ms_auto = s.time_spmv(warmup=2, reps=5)
gb = spmv_bytes(sz["nf_local"], sz["nnz_local"]) / 1e9
gb_dia = (8 * (fmt[1] + 1) + 16) * sz["nf_local"] / 1e9
print(f"format(auto)={fmt} spmv(auto)={ms_auto:.4f} ms -> {gb_dia / ms_auto * 1e3:.1f} GB/s of its own bytes, "
      f"{gb / ms_auto * 1e3:.1f} GB/s CSR-equivalent")
print(f"n={n} Nf={sz['nf_local']} nnz={sz['nnz_local']} assemble(wall incl. H2D)={t1 - t0:.3f}s "
      f"tm={s.timings()} spmv={ms:.4f} ms -> {gb / ms * 1e3:.1f} GB/s")
head, x, ch = s.solve(maxiter=iters)
tm = s.timings()
print(f"pcg {ch.iters} its: {tm['solve_ms'] / max(ch.iters, 1):.4f} ms/it")
s.set_preconditioner("mg")
print("preconditioner:", s.preconditioner())
import time as _t
for rep in range(2):
    t0 = _t.perf_counter()
    head2, x2, ch2 = s.solve(maxiter=max(iters, 200))
    tm = s.timings()
    print(f"mg-pcg: {ch2.iters} its converged={ch2.isconverged} solve={tm['solve_ms']:.1f} ms "
          f"({tm['solve_ms'] / max(ch2.iters, 1):.2f} ms/it) wall={_t.perf_counter() - t0:.3f}s resnorm[-1]={ch2.data['resnorm'][-1]:.3e}")
