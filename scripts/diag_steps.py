"""Per-phase device timings, host wall times and free device memory over repeated assemble+solve steps
(device-resident inputs, then pinned host inputs), to see where a step spends time outside the solve.
usage: [FVB_DEBUG=1] python scripts/diag_steps.py [n] [maxiter]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from bench import problem_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fv = g.load_package()
keep = []


def pinned(shape, dt):
    import numpy as np
    t = torch.empty(shape, dtype={np.int64: torch.int64, np.float64: torch.float64}[dt], pin_memory=True)
    keep.append(t)
    return t.numpy()


P = problem_inputs(fv, n, 1.0, pin=pinned)
lo, hi = P["node_range"]
dev = {k: torch.from_numpy(P[k]).cuda() for k in ("nb", "aol", "kf", "src", "dn", "dh")}
head_dev = torch.empty(hi - lo + 1, dtype=torch.float64, device="cuda")
head_host = torch.empty(hi - lo + 1, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize()
s = fv.System(0)
dptr = {k: v.data_ptr() for k, v in dev.items()}
hptr = {k: P[k].ctypes.data for k in dev}


def step(a, head_ptr, tag):
    t0 = time.perf_counter()
    s.assemble_raw(P["N"], lo, hi, P["F"], a["nb"], a["aol"], a["kf"], P["F"], 0, True, a["src"], P["dn"].size, a["dn"], a["dh"])
    t1 = time.perf_counter()
    it, conv = s.solve_raw(1.49e-8, maxiter, head_ptr=head_ptr)
    t2 = time.perf_counter()
    s.sync()
    tm = s.timings()
    free, total = torch.cuda.mem_get_info()
    print(f"{tag}: assemble_call {1e3 * (t1 - t0):8.1f} ms  solve_call {1e3 * (t2 - t1):8.1f} ms | device: h2d {tm['h2d_ms']:7.1f} "
          f"assemble {tm['assemble_ms']:7.1f} solve {tm['solve_ms']:8.1f} d2h {tm['d2h_ms']:7.1f} | its {it} | "
          f"free {free / 2**30:6.1f} GiB", flush=True)


for i in range(4):
    step(dptr, head_dev.data_ptr(), f"dev  {i}")
for i in range(3):
    step(hptr, head_host.data_ptr(), f"host {i}")
for i in range(2):
    step(dptr, head_dev.data_ptr(), f"dev  {i + 4}")
