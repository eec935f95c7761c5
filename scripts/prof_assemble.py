"""Small driver for ncu captures of the assembly kernels: assemble an n^3 lognormal box `reps` times
(device-resident inputs) and print the device time of the assembly phase.
usage: python scripts/prof_assemble.py [n] [reps] [fmt]   (fmt: FVB_SPMV_FORMAT-style 0/1)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from bench import problem_inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fv = g.load_package()
P = problem_inputs(fv, n, 1.0)
lo, hi = P["node_range"]
dev = {k: torch.from_numpy(P[k]).cuda() for k in ("nb", "aol", "kf", "src", "dn", "dh")}
torch.cuda.synchronize()
s = fv.System(0)
if len(sys.argv) > 3:
    s.set_spmv_format(int(sys.argv[3]))
d = {k: v.data_ptr() for k, v in dev.items()}
for i in range(reps):
    s.assemble_raw(P["N"], lo, hi, P["F"], d["nb"], d["aol"], d["kf"], P["F"], 0, True, d["src"], P["dn"].size, d["dn"], d["dh"])
    tm = s.timings()
    print(f"assemble {i}: device {tm['assemble_ms']:.2f} ms (h2d/d2d {tm['h2d_ms']:.2f} ms) format={s.spmv_format()} "
          f"sizes={s.sizes()}", flush=True)
