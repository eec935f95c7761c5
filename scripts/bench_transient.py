"""BASELINE config 4 timing: test/theis.jl transient (101x101x2, 10 days, dt0=60, atol=1e-4) on the GPU
(device-resident backward-Euler solves, host step controller) next to the CPU oracle's stepper with
cached sparse-LU solves.  usage: python scripts/bench_transient.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402
from test_oracle_pins import theis_setup  # noqa: E402

fv = g.load_package()
P = theis_setup(fv)
tend = 60 * 60 * 24 * 1e1
u0 = np.full(P["src"].size, P["steadyhead"])
for rep in range(2):
    stats = {}
    t0 = time.perf_counter()
    us, ts = fv.backwardeulerintegrate(u0, (0.0, tend), P["Ss"], P["vol"], P["nb"], P["aol"], P["hycos"], P["src"],
                                       P["dn"], P["dh"], atol=1e-4, dt0=60.0, rtol=1e-10, stats=stats)
    t1 = time.perf_counter()
    print(f"gpu: {t1 - t0:.3f}s steps={stats['steps']} solves={stats['linear_solves']} cg_its={stats['cg_iterations']} "
          f"-> {(t1 - t0) / stats['linear_solves'] * 1e3:.3f} ms/solve, {(t1 - t0) / stats['cg_iterations'] * 1e6:.2f} us/cg-iteration")
if "--oracle" in sys.argv:
    from oracle import fv_oracle as orc
    t0 = time.perf_counter()
    uso, tso = orc.backwardeulerintegrate(u0, (0.0, tend), P["Ss"], P["vol"], P["nb"], P["aol"], P["hycos"], P["src"],
                                          P["dn"], P["dh"], atol=1e-4, dt0=60.0)
    print(f"oracle (cached sparse LU per dt): {time.perf_counter() - t0:.3f}s; max|du| = {np.max(np.abs(us[-1] - uso[-1])):.2e}")
