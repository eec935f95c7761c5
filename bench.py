#!/usr/bin/env python
"""bench.py -- steady `solvediffusion` on a heterogeneous regular grid (BASELINE.json metric:
"steady solve time & SpMV GB/s, 512^3 heterogeneous grid, 1/2/4/8 B200").

One step = one pass of the hot path: assemble (per-face conductance -> deterministic CSR build
with Dirichlet elimination + b) followed by the Float64 Jacobi-PCG solve to the reference's
default tolerance (rtol = sqrt(eps), IterativeSolvers default) and the scatter to node heads.

  value      seconds per step with the inputs already resident in HBM (device CUDA events on the
             library's own stream, max over ranks)
  e2e        the same step through the public call with pinned HOST inputs: host->device copies of
             neighbors/areasoverlengths/conductivities/sources/Dirichlet lists and the device->host
             read of the heads are inside the timed region (host clock bracketed by synchronisation)
  roofline   the SpMV kernel of the timed solves (dominant), sampled in situ with CUDA events: on this
             workload the symmetric-diagonal TMA kernel on the Jacobi-scaled copy, 8*K + 16 = 40
             algorithmic bytes per row; `csr_kernel` gives the general CSR kernel timed alone against
             12*nnz + 4*(Nf+1) + 16*Nf
  cpu_baseline / --impl reference
             the CPU restatement of the reference path (oracle/, all host threads) on a bounded
             sample, extrapolated to the workload (the Julia reference itself cannot run here)

N > 1: launched by torchrun, one rank per GPU; the grid is slab-partitioned by x-plane (strong
scaling: the global problem is fixed), halo planes and CG scalars travel over NVLink peer memory
(the CG all-reduces inside the reducing kernels) or, with FVB_P2P=0, over NCCL.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=512, help="grid points per axis (BASELINE metric: 512)")
    ap.add_argument("--sigma", type=float, default=1.0, help="std of ln K (lognormal conductivity)")
    ap.add_argument("--rtol", type=float, default=SQRT_EPS)
    ap.add_argument("--maxiter", type=int, default=200000)
    ap.add_argument("--cpu-sample-n", type=int, default=192, help="grid size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--implicit", action="store_true",
                    help="grid-implicit inputs: upload node ln K only and assemble with fvb_assemble_regulargrid (no per-face "
                         "array exists anywhere); the only way to run --grid 1024 on 1-2 GPUs")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="side measurements only (e.g. 1024^3 scaling points): skip the host-buffer end-to-end leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra legs (device-grid e2e, other preconditioner, CSR kernel timing, tight tolerance)")
    ap.add_argument("--parity-n", type=int, default=96,
                    help="grid size of the oracle parity leg run after the timed region at every N (0: skip)")
    ap.add_argument("--tight-rtol", type=float, default=1e-12,
                    help="tolerance of the extra leg that reports time-to-solution at the parity tolerance (0: skip)")
    ap.add_argument("--ref-budget-s", type=float, default=900.0,
                    help="--impl reference: wall-clock budget for additional timed steps after the first full one")
    ap.add_argument("--precond", default="jacobi", choices=["jacobi", "mg"],
                    help="preconditioner of the headline numbers (north_star: jacobi); the other one is reported "
                         "in the extra object `alt_precond`")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def problem_inputs(fv, n, sigma, planes=None, pin=None):
    """SURVEY 8d workload: regulargrid([0,0,0],[n-1]*3,[n]*3) (unit cells), node ln K =
    ln(1e-5) + sigma*z with z ~ N(0,1) i.i.d. from default_rng(0) in node order, face value =
    mean of the two node logs (nodehycos2neighborhycos(..., true), src/grid.jl:27) passed with
    logtransformconductivity=true, head 1 on the x=0 plane and 0 on the x=n-1 plane, no sources.
    With planes=(lo,hi) only the faces touching those x-planes are built (one rank's slab)."""
    ns = [n, n, n]
    N = n ** 3
    plane = n * n
    if planes is None:
        planes = (1, n)
    F_cap = 3 * (planes[1] - planes[0] + 1) * plane + plane
    alloc = pin if pin is not None else (lambda shape, dt: np.empty(shape, dt))
    nb_buf = alloc((F_cap, 2), np.int64)
    aol_buf = alloc((F_cap,), np.float64)
    _, nb, aol, vol = fv.regulargrid([0, 0, 0], [n - 1] * 3, ns, want_coords=False,
                                     planes=None if planes == (1, n) else planes,
                                     out={"neighbors": nb_buf, "areasoverlengths": aol_buf})
    F = nb.shape[0]
    lnk = math.log(1e-5) + sigma * np.random.default_rng(0).standard_normal(N)
    kf = alloc((F,), np.float64)
    step = 1 << 24
    for o in range(0, F, step):  # chunked to bound temporaries
        e = min(F, o + step)
        np.add(lnk[nb[o:e, 0] - 1], lnk[nb[o:e, 1] - 1], out=kf[o:e])
    kf *= 0.5
    lo, hi = (planes[0] - 1) * plane + 1, planes[1] * plane
    # node values of the slab plus one plane on each side (what the device-side grid path uploads)
    k_lo, k_hi = max(1, lo - plane), min(N, hi + plane)
    lnk_slab = alloc((k_hi - k_lo + 1,), np.float64)
    lnk_slab[:] = lnk[k_lo - 1:k_hi]
    del lnk
    src = alloc((hi - lo + 1,), np.float64)
    src[:] = 0.0
    dn = alloc((2 * plane,), np.int64)
    dn[:plane] = np.arange(1, plane + 1)
    dn[plane:] = np.arange(N - plane + 1, N + 1)
    dh = alloc((2 * plane,), np.float64)
    dh[:plane] = 1.0
    dh[plane:] = 0.0
    return dict(N=N, F=F, node_range=(lo, hi), nb=nb, aol=aol, kf=kf, src=src, dn=dn, dh=dh, lnk_slab=lnk_slab,
                lnk_node_lo=k_lo)


def implicit_inputs(n, sigma, planes, pin, chunk=1 << 24):
    """The same workload described by node values only: ln K of the owned x-planes plus one plane on each side
    (what fvb_assemble_regulargrid takes), the two Dirichlet planes, zero sources (passed as NULL)."""
    N, plane = n ** 3, n * n
    lo, hi = (planes[0] - 1) * plane + 1, planes[1] * plane
    k_lo, k_hi = max(1, lo - plane), min(N, hi + plane)
    lnk_slab = pin((k_hi - k_lo + 1,), np.float64)
    rng = np.random.default_rng(0)
    # the global field is drawn in node order from one stream (same numbers as problem_inputs); stream through it
    pos = 0
    while pos < k_hi:
        m = min(chunk, N - pos)
        z = rng.standard_normal(m)
        a, b = max(pos, k_lo - 1), min(pos + m, k_hi)
        if a < b:
            lnk_slab[a - (k_lo - 1):b - (k_lo - 1)] = math.log(1e-5) + sigma * z[a - pos:b - pos]
        pos += m
    dn = pin((2 * plane,), np.int64)
    dn[:plane] = np.arange(1, plane + 1)
    dn[plane:] = np.arange(N - plane + 1, N + 1)
    dh = pin((2 * plane,), np.float64)
    dh[:plane] = 1.0
    dh[plane:] = 0.0
    return dict(N=N, F=0, node_range=(lo, hi), lnk_slab=lnk_slab, lnk_node_lo=k_lo, dn=dn, dh=dh)


def spmv_bytes(nf, nnz):
    """int32 column indices and row pointers, f64 values, x read once, y written once."""
    return 12 * nnz + 4 * (nf + 1) + 16 * nf


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def host_mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return None


def oracle_peak_bytes(n):
    """Peak host memory of the oracle's solvediffusion on an n^3 grid: inputs 32 B/face, the COO triples of
    src/FiniteVolume.jl:94-105 (3 x 8 B x 4 per face), sparse!'s scratch and output arrays (2 x 16 B per triple),
    the trimmed CSC copy and the solver vectors."""
    F, N = 3 * n ** 3 - 3 * n * n, n ** 3
    return 300 * F + 120 * N


def _oracle(threads=None):
    from oracle import fv_oracle as orc
    orc.build()
    # all the host threads this process may use (torchrun presets OMP_NUM_THREADS=1 for its workers, which would
    # silently turn the multi-rank launch of the reference arm into a single-threaded run)
    if threads is None:
        try:
            threads = max(orc.num_threads(), len(os.sched_getaffinity(0)))
        except (AttributeError, OSError):
            threads = orc.num_threads()
    orc.set_num_threads(threads)
    return orc


def cpu_reference_solve(n_s, sigma, rtol, threads=None, keep=False):
    """One complete step of the CPU restatement of the reference path (oracle/: serial assembly exactly as
    src/FiniteVolume.jl:75-139 + sparse!, Jacobi-PCG with the threaded row gather) on the n_s^3 instance of the
    bench workload.  Returns the measured times; nothing is extrapolated here."""
    orc = _oracle(threads)
    threads = orc.num_threads()
    ns = [n_s] * 3
    _, nb, aol, vol = orc.regulargrid([0, 0, 0], [n_s - 1] * 3, ns, want_coords=False)
    del vol
    N = n_s ** 3
    plane = n_s * n_s
    lnk = math.log(1e-5) + sigma * np.random.default_rng(0).standard_normal(N)
    kf = orc.nodehycos2neighborhycos(nb, lnk, True)
    del lnk
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    dh = np.concatenate([np.ones(plane), np.zeros(plane)])
    src = np.zeros(N)
    t1 = time.perf_counter()
    A = orc.assembleA(nb, aol, kf, src, dn, dh, None, True)
    b = orc.assembleb(nb, aol, kf, src, dn, dh, None, True)
    t2 = time.perf_counter()
    F_s = nb.shape[0]
    if not keep:
        del nb, aol, kf
    x, ch = orc.cg(A, b, Pl="jacobi", tol=rtol, maxiter=200000, threaded=threads > 1)
    head, _, _ = orc.freenodes2nodes(x, src, dn, dh)
    t3 = time.perf_counter()
    out = dict(seconds=t3 - t1, assemble_s=t2 - t1, solve_s=t3 - t2, iters=int(ch.iters),
               converged=bool(ch.isconverged), threads=threads, nf=int(A.n), faces=int(F_s), n=n_s,
               sane=bool(head.min() >= -1e-6 and head.max() <= 1 + 1e-6))
    if keep:
        out.update(A=A, b=b, head=head)
    return out


def cpu_reference_sample(n_s, sigma, rtol, target_n, target_iters):
    """cpu_baseline of the GPU arm: a BOUNDED sample (an n_s^3 instance of the same workload solved completely by
    the oracle) scaled to the workload: assembly by the number of faces, PCG by Nf * iterations with the iteration
    count the GPU arm measured on the workload itself.  The measured, un-extrapolated number is the
    `--impl reference` arm."""
    r = cpu_reference_solve(n_s, sigma, rtol)
    nf_t = target_n ** 3 - 2 * target_n ** 2
    F_t = 3 * target_n ** 3 - 3 * target_n ** 2
    asm_t = r["assemble_s"] * F_t / r["faces"]
    pcg_t = r["solve_s"] / (r["nf"] * max(r["iters"], 1)) * nf_t * target_iters
    r.update(value=asm_t + pcg_t,
             sample=f"{n_s}^3 instance of the same workload solved completely (assemble {r['assemble_s']:.2f}s + "
                    f"Jacobi-PCG {r['iters']} its {r['solve_s']:.2f}s, rtol={rtol:.3g}, {r['threads']} OpenMP threads), "
                    f"EXTRAPOLATED to {target_n}^3 by F for assembly and by Nf*iterations for PCG with the "
                    f"{target_iters} iterations the GPU arm measured at {target_n}^3; restated reference "
                    "(oracle port): Julia + RS-AMG are unavailable offline and the true reference is single-threaded; "
                    "the un-extrapolated measurement is `bench.py --impl reference`")
    return r


def run_reference(args, rank, world):
    """The reference arm: the oracle port of the reference's CPU path solving the configuration this line names,
    completely, on all host threads.  A 512^3 step takes minutes on the host, so the first full step IS the timed
    step (reported as steps=1, warmup=0) and further steps are only added while they fit --ref-budget-s."""
    if rank != 0:
        return
    n = args.n
    avail = host_mem_available_bytes()
    cands = [n] + [m for m in (448, 384, 320, 256, 192, 128, 96, 64) if m < n]
    n_run = cands[-1]
    for m in cands:
        if avail is None or oracle_peak_bytes(m) <= 0.92 * avail:
            n_run = m
            break
    t_all = time.perf_counter()
    runs = [cpu_reference_solve(n_run, args.sigma, args.rtol)]
    warm = 0
    # more steps only when a whole warm-up + K schedule of them is cheap
    while (len(runs) < args.warmup + args.steps
           and (time.perf_counter() - t_all) + 1.2 * runs[-1]["seconds"] < args.ref_budget_s):
        runs.append(cpu_reference_solve(n_run, args.sigma, args.rtol))
    if len(runs) >= args.warmup + args.steps:
        warm = args.warmup
    elif len(runs) > 1:
        warm = 1
    timed = runs[warm:]
    v = float(np.mean([r["seconds"] for r in timed]))
    r0 = timed[-1]
    # JULIA_NUM_THREADS=1-equivalent figure (the reference has no threaded code): a bounded single-thread sample
    single = None
    try:
        ns1 = min(n_run, 128)
        s1 = cpu_reference_solve(ns1, args.sigma, args.rtol, threads=1)
        nf_t, F_t = n_run ** 3 - 2 * n_run ** 2, 3 * n_run ** 3 - 3 * n_run ** 2
        single = {"sample": f"{ns1}^3 instance solved completely with 1 thread (assemble {s1['assemble_s']:.2f}s + "
                            f"{s1['iters']} its {s1['solve_s']:.2f}s)",
                  "sample_seconds": s1["seconds"], "cores": 1,
                  "extrapolated_value": s1["assemble_s"] * F_t / s1["faces"]
                  + s1["solve_s"] / (s1["nf"] * max(s1["iters"], 1)) * nf_t * r0["iters"],
                  "note": f"extrapolated to {n_run}^3 by F (assembly) and Nf*iterations (PCG, {r0['iters']} iterations "
                          "measured by this arm); JULIA_NUM_THREADS=1-equivalent, a labelled estimate, not the line's value"}
        _oracle()  # restore all threads
    except Exception as e:  # the side figure must not take the line down
        single = {"error": str(e)}
    a2 = argparse.Namespace(**vars(args))
    a2.n = n_run
    cfg = workload_config(a2, 1)
    cfg["parallelism"] = f"host CPU, {r0['threads']} OpenMP threads (PCG); assembly serial as in the reference"
    if n_run != n:
        cfg["note"] = (f"host MemAvailable {avail / 1e9:.0f} GB cannot hold the {n}^3 oracle run "
                       f"(~{oracle_peak_bytes(n) / 1e9:.0f} GB): this line measures the largest grid that fits")
    sample = (f"{n_run}^3 workload solved completely, nothing extrapolated: serial assembly {r0['assemble_s']:.1f}s + "
              f"Jacobi-PCG {r0['iters']} iterations {r0['solve_s']:.1f}s (rtol={args.rtol:.3g}, {r0['threads']} OpenMP "
              f"threads, converged={r0['converged']}); oracle port of the reference path -- Julia and RS-AMG "
              "are unavailable offline, the true reference is single-threaded AMG-PCG")
    line = {
        "impl": "reference", "metric": "steady_solvediffusion_time", "value": v, "unit": "s", "n_gpus": args.gpus,
        "steps": len(timed), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": v * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "s", "cores": r0["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pcg_iterations": r0["iters"], "converged": r0["converged"], "result_sane": r0["sane"],
        "assemble_s": r0["assemble_s"], "solve_s": r0["solve_s"], "single_thread": single,
        "wall_s_total": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    n = args.n
    return {"workload": f"{n}^3 regulargrid, lognormal K (sigma={args.sigma:g}, seed 0), left/right Dirichlet 1/0, "
                        f"steady solvediffusion = assemble + Jacobi-PCG to rtol={args.rtol:.3g}",
            "grid": [n, n, n], "nodes": n ** 3, "faces": 3 * n ** 3 - 3 * n ** 2, "free_rows": n ** 3 - 2 * n ** 2,
            "nnz": (n ** 3 - 2 * n * n) + 2 * ((n - 3) * n * n + 2 * (n - 2) * n * (n - 1)),
            "parallelism": f"slab{world}" if world > 1 else "single",
            "l2_policy": "inputs larger than L2 (CSR alone exceeds 126 MB); no explicit flush"}


def parity_leg(fv, fvd, sysm, args, rank, world, dist, kern):
    """Slab-partitioned solve of a small instance of the bench workload on the ranks of this run, checked on rank 0
    against the CPU oracle (tests/ do the same at more sizes; this puts the check into the bench record at every N).
    CSR: the concatenation of the ranks' rows must equal the oracle's CSC arrays -- structure bit for bit, values
    within 1e-14 relative (log K goes through exp, device vs libm <= 1 ulp).  Heads: <= 1e-8 relative at rtol 1e-12."""
    n = args.parity_n
    planes = fvd.slab_planes(n, world) if world > 1 else [(1, n)]
    P = problem_inputs(fv, n, args.sigma, planes=planes[rank])
    lo, hi = P["node_range"]
    forced = 3 if kern == "dia_tma" else (2 if kern == "dia" else 1)
    sysm.set_preconditioner("jacobi")
    sysm.set_spmv_format(forced)
    try:
        sysm.assemble_raw(P["N"], lo, hi, P["F"], P["nb"].ctypes.data, P["aol"].ctypes.data, P["kf"].ctypes.data, P["F"],
                          0, True, P["src"].ctypes.data, P["dn"].size, P["dn"].ctypes.data, P["dh"].ctypes.data)
        if world > 1:
            fvd.exchange_halo_plan(sysm)
        head = np.empty(hi - lo + 1)
        it, conv = sysm.solve_raw(1e-12, 200000, head_ptr=head.ctypes.data)
        used = sysm.spmv_kernel()
        scaled = sysm.pcg_scaling()
        ptr_, idx_, val_ = sysm.csr()
        b_ = sysm.b()
    finally:
        sysm.set_spmv_format(0)
        sysm.set_preconditioner(args.precond)
    mine = (head, ptr_, idx_, val_, b_, it, bool(conv), used, bool(scaled))
    if world > 1:
        got = [None] * world if rank == 0 else None
        dist.gather_object(mine, got, dst=0)
    else:
        got = [mine]
    if rank != 0:
        return None
    t0 = time.perf_counter()
    ref = cpu_reference_solve(n, args.sigma, 1e-12, keep=True)
    A, bo, ho = ref["A"], ref["b"], ref["head"]
    hg = np.concatenate([g[0] for g in got])
    idx = np.concatenate([g[2] for g in got])
    val = np.concatenate([g[3] for g in got])
    bg = np.concatenate([g[4] for g in got])
    ptrs, off = [], 0
    for g in got:
        ptrs.append(g[1][:-1] - 1 + off)
        off += int(g[1][-1] - 1)
    ptr_all = np.concatenate(ptrs + [np.array([off])]) + 1
    structure = bool(ptr_all.size == A.colptr.size and np.array_equal(ptr_all, A.colptr)
                     and idx.size == A.rowval.size and np.array_equal(idx, A.rowval))
    val_rel = float(np.max(np.abs(val - A.nzval) / np.abs(A.nzval))) if structure else None
    b_rel = float(np.max(np.abs(bg - bo) / np.maximum(np.abs(bo), 1e-300))) if bg.size == bo.size else None
    err_h = float(np.max(np.abs(hg - ho)) / np.max(np.abs(ho)))
    iters = sorted({g[5] for g in got})
    ok = bool(structure and val_rel is not None and val_rel <= 1e-14 and b_rel is not None and b_rel <= 1e-14
              and err_h <= 1e-8 and all(g[6] for g in got) and len(iters) == 1 and ref["converged"])
    return {"grid": [n, n, n], "ranks": world, "rtol": 1e-12, "csr_bitexact": structure,
            "csr_values_bitexact": bool(structure and np.array_equal(val, A.nzval)), "csr_values_max_rel": val_rel,
            "b_max_rel": b_rel, "err_head": err_h, "tol_head": 1e-8, "pcg_iterations": iters,
            "oracle_iterations": ref["iters"], "kernels": sorted({g[7] for g in got}), "pcg_scaled": all(g[8] for g in got),
            "oracle": "oracle/fv_oracle.c (CPU restatement of src/FiniteVolume.jl:75-165), same seeded inputs",
            "oracle_seconds": time.perf_counter() - t0, "ok": ok}


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        if rank == 0:
            g.build_library()
    fv = g.load_package()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    import importlib
    fvd = importlib.import_module("fvb200.distributed")

    n = args.n
    planes = fvd.slab_planes(n, world) if world > 1 else [(1, n)]
    mine = planes[rank]

    def pinned(shape, dt):
        tdt = {np.int64: torch.int64, np.float64: torch.float64}[dt]
        t = torch.empty(shape, dtype=tdt, pin_memory=True)
        pinned.keep.append(t)
        return t.numpy()
    pinned.keep = []

    implicit = bool(args.implicit)
    t_gen = time.perf_counter()
    if implicit:
        P = implicit_inputs(n, args.sigma, mine, pinned)
        in_keys = ("lnk_slab", "dn", "dh")
    else:
        P = problem_inputs(fv, n, args.sigma, planes=mine, pin=pinned)
        in_keys = ("nb", "aol", "kf", "src", "dn", "dh")
    t_gen = time.perf_counter() - t_gen
    lo, hi = P["node_range"]
    h2d_bytes = sum(P[k].nbytes for k in in_keys)
    d2h_bytes = (hi - lo + 1) * 8

    # device-resident copies for the `value` leg
    dev = {k: torch.from_numpy(P[k]).cuda(non_blocking=True) for k in in_keys}
    head_dev = torch.empty(hi - lo + 1, dtype=torch.float64, device="cuda")
    head_host = torch.empty(hi - lo + 1, dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()

    sysm = fv.System(local_rank)
    if world > 1:
        fvd.init_comm(sysm)
    sysm.set_profiling(50)
    sysm.set_preconditioner(args.precond)
    transport = "single GPU" if world == 1 else (
        "nccl (send/recv halo + all-reduce)" if os.environ.get("FVB_P2P", "1") == "0"
        else "peer (NVLink peer-memory halo stores + in-kernel all-reduce, CUDA IPC)")

    def barrier():
        sysm.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    wall_parts = {}

    def assemble_from(a):
        """One assembly from the addresses in `a` (device-resident copies or pinned host buffers)."""
        if implicit:
            sysm.assemble_regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], a["lnk_slab"], None, P["dn"], P["dh"],
                                      logmean=True, logtransformconductivity=True, planes=mine)
        else:
            sysm.assemble_raw(P["N"], lo, hi, P["F"], a["nb"], a["aol"], a["kf"], P["F"], 0, True, a["src"],
                              P["dn"].size, a["dn"], a["dh"])

    def step(src_arrays, head_ptr):
        a = src_arrays
        t0 = time.perf_counter()
        assemble_from(a)
        t1 = time.perf_counter()
        if world > 1:
            fvd.exchange_halo_plan(sysm)
        t2 = time.perf_counter()
        it, conv = sysm.solve_raw(args.rtol, args.maxiter, head_ptr=head_ptr)
        wall_parts.update(assemble_call_s=t1 - t0, halo_plan_s=t2 - t1, solve_call_s=time.perf_counter() - t2)
        return it, conv

    def step_devgrid(head_ptr):
        """Same solve, but the grid never exists on the host: upload the node ln K of the slab (+1 plane each
        side), build neighbors / areasoverlengths / face K on the device (fvb_regulargrid,
        fvb_nodehycos2neighborhycos), assemble from those device arrays."""
        t0 = time.perf_counter()
        nbd, aold, _ = sysm.device_regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], planes=mine, want_volumes=False)
        t1 = time.perf_counter()
        kfd = sysm.device_nodehycos2neighborhycos(nbd, P["lnk_slab"], True, node_lo=P["lnk_node_lo"])
        t2 = time.perf_counter()
        sysm.assemble_raw(P["N"], lo, hi, nbd.shape[0], nbd.ptr, aold.ptr, kfd.ptr, nbd.shape[0], 0, True,
                          host_ptrs["src"], P["dn"].size, host_ptrs["dn"], host_ptrs["dh"])
        t3 = time.perf_counter()
        if world > 1:
            fvd.exchange_halo_plan(sysm)
        r = sysm.solve_raw(args.rtol, args.maxiter, head_ptr=head_ptr)
        wall_parts.update(grid_s=t1 - t0, facek_s=t2 - t1, assemble_call_s=t3 - t2, solve_call_s=time.perf_counter() - t3)
        return r

    dev_ptrs = {k: v.data_ptr() for k, v in dev.items()}
    host_ptrs = {k: P[k].ctypes.data for k in in_keys}

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- warm-up --------------------------------------------------------------------------
    for _ in range(args.warmup):
        it, conv = step(dev_ptrs, head_dev.data_ptr())
    # ---- timed: device-resident inputs ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = sysm.timings()["kernel_launches"]
    dev_ms, spmv_ms, spmv_samples = 0.0, 0.0, 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        it, conv = step(dev_ptrs, head_dev.data_ptr())
        tm = sysm.timings()
        dev_ms += tm["h2d_ms"] + tm["assemble_ms"] + tm["solve_ms"] + tm["d2h_ms"]
        spmv_ms += tm["spmv_ms_total"]
        spmv_samples += tm["spmv_samples"]
    barrier()
    wall = time.perf_counter() - w0
    launches = sysm.timings()["kernel_launches"] - l0
    scaled = sysm.pcg_scaling()  # the timed solves ran the symmetrically scaled recurrence (unit-diagonal SpMV)
    kern = sysm.spmv_kernel()    # "csr" | "dia" (per-thread loads) | "dia_tma" (TMA pipeline)
    fmt, fmt_k = sysm.spmv_format()
    clocks = sampler.stop()
    last_tm = sysm.timings()
    sz = sysm.sizes()
    step_s = maxreduce(dev_ms / args.steps / 1e3)
    wall_s = maxreduce(wall / args.steps)

    # ---- timed: end to end from pinned host buffers -------------------------------------------
    # (one untimed pass first: the host-pointer call stages its inputs in device buffers the
    # device-resident passes above never allocated, and growing the memory pool is a one-off cost)
    if args.skip_e2e:
        it_e, conv_e = step(dev_ptrs, head_dev.data_ptr())
        head_host.copy_(head_dev)
        barrier()
        e2e_s = None
    else:
        step(host_ptrs, head_host.data_ptr())
        barrier()
        e0 = time.perf_counter()
        for _ in range(max(args.e2e_steps, 1)):
            it_e, conv_e = step(host_ptrs, head_host.data_ptr())
        barrier()
        e2e_s = maxreduce((time.perf_counter() - e0) / max(args.e2e_steps, 1))
    e2e_tm = sysm.timings()
    e2e_parts = dict(wall_parts)
    # ---- end to end with the device-side grid generator (extra information) ---------------------
    head_e2e = head_host.numpy().copy()
    assembly_kind = sysm.assembly()  # "box": closed-form rows from the caller's arrays; "implicit": no face arrays
    extras = not args.no_extras
    devgrid, implicit_leg = None, None
    if extras and not implicit:
        step_devgrid(head_host.data_ptr())  # warm-up of the extra allocations
        barrier()
        g0 = time.perf_counter()
        it_g, conv_g = step_devgrid(head_host.data_ptr())
        barrier()
        devgrid = {"value": maxreduce(time.perf_counter() - g0), "unit": "s", "pcg_iterations": it_g,
                   "h2d_bytes_per_step": int(P["lnk_slab"].nbytes + P["src"].nbytes + P["dn"].nbytes + P["dh"].nbytes),
                   "identical_heads": bool(np.array_equal(head_host.numpy(), head_e2e)), "host_wall": dict(wall_parts),
                   "solve_ms": sysm.timings()["solve_ms"], "assemble_ms": sysm.timings()["assemble_ms"],
                   "format": list(sysm.spmv_format()), "assembly": sysm.assembly(),
                   "note": "host uploads node ln K only; neighbors/areasoverlengths/face K generated on the device"}
        # the same problem with no face array at all (fvb_assemble_regulargrid)
        def step_implicit(head_ptr):
            sysm.assemble_regulargrid([0, 0, 0], [n - 1] * 3, [n, n, n], P["lnk_slab"].ctypes.data, None, P["dn"], P["dh"],
                                      logmean=True, logtransformconductivity=True, planes=mine)
            if world > 1:
                fvd.exchange_halo_plan(sysm)
            return sysm.solve_raw(args.rtol, args.maxiter, head_ptr=head_ptr)
        step_implicit(head_host.data_ptr())
        barrier()
        g0 = time.perf_counter()
        it_i, conv_i = step_implicit(head_host.data_ptr())
        barrier()
        implicit_leg = {"value": maxreduce(time.perf_counter() - g0), "unit": "s", "pcg_iterations": it_i,
                        "h2d_bytes_per_step": int(P["lnk_slab"].nbytes + P["dn"].nbytes + P["dh"].nbytes),
                        "identical_heads": bool(np.array_equal(head_host.numpy(), head_e2e)),
                        "solve_ms": sysm.timings()["solve_ms"], "assemble_ms": sysm.timings()["assemble_ms"],
                        "assembly": sysm.assembly(),
                        "note": "grid-implicit assembly: no neighbors/areasoverlengths/conductivities array exists, "
                                "rows computed from node ln K and the grid spacing"}

    # ---- the other preconditioner on the same resident inputs (extra information, not the headline) ----
    alt = "mg" if args.precond == "jacobi" else "jacobi"
    alt_info = None
    # The distributed V-cycle passed tests/test_gpu_multi.py at 2, 4 and 8 ranks on both transports
    # (profiles/r2_multi_pytest.log), so the leg runs at every N; FVB_BENCH_ALT=0 switches it off.
    run_alt = extras and os.environ.get("FVB_BENCH_ALT", "1") != "0"
    try:
        if not extras:
            raise RuntimeError("skipped (--no-extras)")
        if not run_alt:
            raise RuntimeError("skipped (FVB_BENCH_ALT=0)")
        head_main = head_host.numpy().copy()
        sysm.set_preconditioner(alt)
        step(dev_ptrs, head_dev.data_ptr())  # warm-up (allocates the hierarchy)
        barrier()
        a_ms = 0.0
        for _ in range(args.steps):
            it_a, conv_a = step(dev_ptrs, head_dev.data_ptr())
            tma = sysm.timings()
            a_ms += tma["h2d_ms"] + tma["assemble_ms"] + tma["solve_ms"] + tma["d2h_ms"]
        barrier()
        ea = time.perf_counter()
        it_a, conv_a = step(host_ptrs, head_host.data_ptr())
        barrier()
        ea = maxreduce(time.perf_counter() - ea)
        tma = sysm.timings()
        barrier()
        eg = None
        if not implicit:
            eg = time.perf_counter()
            step_devgrid(head_host.data_ptr())
            barrier()
            eg = maxreduce(time.perf_counter() - eg)
        diff = float(np.max(np.abs(head_host.numpy() - head_main)))
        alt_info = {"precond": alt, "active": sysm.preconditioner()[0], "value": maxreduce(a_ms / args.steps / 1e3),
                    "unit": "s", "e2e": ea, "e2e_device_grid": eg, "pcg_iterations": it_a, "converged": bool(conv_a),
                    "solve_ms": tma["solve_ms"], "assemble_ms": tma["assemble_ms"], "h2d_ms": tma["h2d_ms"],
                    "max_abs_head_difference_vs_headline": diff,
                    "note": "same inputs, same tolerance; heads differ by solver tolerance only"}
        sysm.set_preconditioner(args.precond)
        head_host.numpy()[:] = head_main
    except Exception as e:  # the alternative leg must never take the headline down
        alt_info = {"precond": alt, "error": str(e)}

    # ---- time-to-solution at the parity tolerance (report timing at sqrt(eps) AND at the tolerance
    #      where "heads within 1e-8" is a meaningful statement), same resident inputs ---------------------------
    tight = None
    if extras and args.tight_rtol and args.tight_rtol > 0:
        try:
            head_tj = torch.empty(hi - lo + 1, dtype=torch.float64, pin_memory=True)
            sysm.set_preconditioner("jacobi")
            barrier()
            assemble_from(dev_ptrs)
            if world > 1:
                fvd.exchange_halo_plan(sysm)
            it_t, conv_t = sysm.solve_raw(args.tight_rtol, args.maxiter, head_ptr=head_tj.data_ptr())
            tmt = sysm.timings()
            t_j = maxreduce((tmt["assemble_ms"] + tmt["solve_ms"] + tmt["d2h_ms"]) / 1e3)
            tight = {"rtol": args.tight_rtol, "jacobi": {"value": t_j, "unit": "s", "pcg_iterations": it_t,
                                                       "converged": bool(conv_t), "solve_ms": tmt["solve_ms"]},
                     "max_abs_head_difference_sqrt_eps_vs_tight":
                         maxreduce(float(np.max(np.abs(head_tj.numpy() - head_e2e)))),
                     "note": "value = assemble + solve + head read-back with device-resident inputs (one step)"}
            if run_alt:
                head_tm = torch.empty(hi - lo + 1, dtype=torch.float64, pin_memory=True)
                sysm.set_preconditioner("mg")
                assemble_from(dev_ptrs)
                if world > 1:
                    fvd.exchange_halo_plan(sysm)
                it_m, conv_m = sysm.solve_raw(args.tight_rtol, args.maxiter, head_ptr=head_tm.data_ptr())
                tmm = sysm.timings()
                active = sysm.preconditioner()[0]
                dj = maxreduce(float(np.max(np.abs(head_tm.numpy() - head_tj.numpy()))))
                tight["mg"] = {"value": maxreduce((tmm["assemble_ms"] + tmm["solve_ms"] + tmm["d2h_ms"]) / 1e3),
                               "unit": "s", "pcg_iterations": it_m, "converged": bool(conv_m), "active": active,
                               "solve_ms": tmm["solve_ms"]}
                tight["max_abs_head_difference_jacobi_vs_mg"] = dj
                tight["heads_agree_1e-8"] = bool(dj <= 1e-8)
                sysm.set_preconditioner(args.precond)
        except Exception as e:  # an extra leg must never take the headline down
            tight = {"error": repr(e)}
            sysm.set_preconditioner(args.precond)

    # sanity of the result that was timed: maximum principle + convergence (not a parity test)
    hh = head_host.numpy()
    ok = bool(conv and conv_e and hh.min() >= -1e-6 and hh.max() <= 1 + 1e-6)

    spmv_avg_ms = spmv_ms / max(spmv_samples, 1)
    csr_bytes = spmv_bytes(sz["nf_local"], sz["nnz_local"])
    # an index-free diagonal format is reported against ITS algorithmic bytes
    # (8 per stored diagonal entry incl. the main diagonal, x read once, y written once)
    # (the scaled copy has a unit diagonal that is not stored: 8 bytes per row fewer)
    alg_bytes = (8 * (fmt_k + (0 if scaled else 1)) + 16) * sz["nf_local"] if fmt == "dia" else csr_bytes
    achieved = alg_bytes / (spmv_avg_ms * 1e-3) / 1e9 if spmv_samples else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("grid_n") == n and world == 1:
                # per launch, from the committed ncu --set full capture
                traffic = tj.get("dia_scaled" if (fmt == "dia" and scaled) else fmt)
                traffic_src = ("static: dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from "
                               "the committed `ncu --set full` capture (" + str(tj.get("source", "profiles/")) + "), "
                               "not measured in this run")
        except Exception:
            pass

    # the general CSR kernel on the same resident matrix, timed alone (20 launches, CUDA events)
    csr_roof = None
    if world == 1 and extras and sz["nnz_local"] < 2 ** 31 - 64:
        sysm.set_spmv_format(1)
        ms_csr = sysm.time_spmv(warmup=3, reps=20)
        sysm.set_spmv_format(0)
        csr_roof = {"kernel": "k_spmv<false> (CSR, TMA-staged)", "avg_launch_ms": ms_csr,
                    "algorithmic_bytes_per_launch": int(csr_bytes), "achieved": csr_bytes / (ms_csr * 1e-3) / 1e9,
                    "frac": csr_bytes / (ms_csr * 1e-3) / 1e9 / peak, "timed": "alone, 20 launches"}
    if world > 1:
        ach_t = torch.tensor([achieved or 0.0], dtype=torch.float64)
        dist.all_reduce(ach_t, op=dist.ReduceOp.MIN)
        achieved = float(ach_t[0]) or None

    # ---- oracle parity of the N-rank path, in the driver-run record: a parity_n^3 instance of the same workload,
    #      slab-partitioned over the same ranks, same kernels (format forced to the one the timed solves used),
    #      rtol 1e-12; rank 0 compares the per-slab CSR rows and the heads with the CPU oracle -----------------
    parity = None
    if args.parity_n and args.parity_n > 0:
        try:
            parity = parity_leg(fv, fvd, sysm, args, rank, world, dist if world > 1 else None, kern)
        except Exception as e:
            parity = {"ok": False, "error": repr(e)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            info = cpu_reference_sample(args.cpu_sample_n, args.sigma, args.rtol, n, target_iters=it)
            cpu = {"value": info["value"], "unit": "s", "cores": info["threads"], "kind": "port",
                   "sample": info["sample"]}
        line = {
            "metric": "steady_solvediffusion_time", "value": step_s, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, world),
                           inputs=("grid-implicit: node ln K + grid spacing only (fvb_assemble_regulargrid)" if implicit else
                                   "the reference's per-face arrays: neighbors, areasoverlengths, conductivities (fvb_assemble)")),
            "clocks": clocks,
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "h2d_ms": e2e_tm["h2d_ms"], "assemble_ms": e2e_tm["assemble_ms"], "solve_ms": e2e_tm["solve_ms"],
                    "d2h_ms": e2e_tm["d2h_ms"], "pcg_iterations": it_e, "host_wall": e2e_parts},
            "e2e_device_grid": devgrid,
            "e2e_implicit_grid": implicit_leg,
            "assembly": assembly_kind,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": ("%s<true,%d,%s> (symmetric-diagonal SpMV%s + fused u.Au)"
                                    % ("k_spmv_dia_tma" if kern == "dia_tma" else "k_spmv_dia", fmt_k,
                                       "true" if scaled else "false",
                                       " on the Jacobi-scaled unit-diagonal copy" if scaled else "")) if fmt == "dia"
                         else "k_spmv<true> (CSR SpMV + fused u.Au)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "avg_launch_ms": spmv_avg_ms, "launches_sampled": int(spmv_samples), "format": fmt,
                         "spmv_kernel": kern,
                         "csr_equivalent_gbs": (csr_bytes / (spmv_avg_ms * 1e-3) / 1e9) if spmv_samples else None,
                         "csr_kernel": csr_roof,
                         "note": "rank-local rows; min over ranks" if world > 1 else "sampled inside the timed solves"},
            "cpu_baseline": cpu,
            "precond": args.precond, "pcg_scaled": bool(scaled),
            "pcg_bytes_per_row_per_iteration": (112 if scaled else 128) if fmt == "dia" else None,
            "alt_precond": alt_info,
            "tight_tolerance": tight,
            "parity": parity,
            "transport": transport,
            "pcg_iterations": it, "converged": bool(conv), "result_sane": ok,
            "assemble_ms": last_tm["assemble_ms"], "solve_ms": last_tm["solve_ms"],
            "wall_s_per_step": wall_s, "input_generation_s": t_gen,
            "rows_local": sz["nf_local"], "nnz_local": sz["nnz_local"],
            "pcg_iteration_ms": last_tm["solve_ms"] / max(it, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
