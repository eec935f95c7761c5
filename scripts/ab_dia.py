"""A/B of the diagonal-format SpMV kernels and of the scaled recurrence on one GPU.
usage: python scripts/ab_dia.py [n] [pcg_iters]
Prints the alone-timed SpMV (20 launches, CUDA events) for the per-thread-load kernel (format 2) and the
TMA pipeline (format 3), and ms/iteration of maxiter-capped solves for every (format, scaling) pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from bench import problem_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fv = g.load_package()
P = problem_inputs(fv, n, 1.0)
s = fv.System(0)
s.assemble(P["nb"], P["aol"], P["kf"], P["src"], P["dn"], P["dh"], None, True)
nf = s.sizes()["nf_local"]
K = s.spmv_format()[1]
print(f"n={n} nf={nf} format={s.spmv_format()} auto kernel={s.spmv_kernel()}")
for fmt, name in ((2, "dia (per-thread loads)"), (3, "dia_tma")):
    s.set_spmv_format(fmt)
    ms = s.time_spmv(warmup=3, reps=20)
    gb = (8 * (K + 1) + 16) * nf / 1e9
    print(f"spmv alone  fmt={fmt} {name:24s} {ms:.4f} ms  {gb / ms * 1e3:.0f} GB/s (48 B/row)")
s.set_profiling(10)
for fmt in (2, 3):
    for scal in (1, 0):
        s.set_spmv_format(fmt)
        s.set_pcg_scaling(scal)
        s.solve(maxiter=20, want_head=False)
        _, _, ch = s.solve(maxiter=iters, want_head=False)
        tm = s.timings()
        bpr = 8 * (K + (0 if s.pcg_scaling() else 1)) + 16
        sp = tm["spmv_ms_total"] / max(tm["spmv_samples"], 1)
        print(f"pcg fmt={fmt} scaled={s.pcg_scaling()} {ch.iters} its: {tm['solve_ms'] / max(ch.iters, 1):.4f} ms/it; "
              f"spmv in situ {sp:.4f} ms = {bpr * nf / 1e9 / sp * 1e3:.0f} GB/s ({bpr} B/row)")
