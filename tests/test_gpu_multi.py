"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): slab partition, halo exchange and all-reduced
CG scalars must reproduce the oracle's heads within 1e-8 relative -- one process per GPU under torchrun (both
transports) and one process driving all GPUs through the fvb_multi front end.  The cases "3,6,5" on 2 ranks and
"9,6,5" on 8 give the last rank the Dirichlet plane only: a rank without free rows."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run_group(cmd, env, timeout):
    """Run a launcher in its own process group and, on timeout, kill the WHOLE group: a killed torchrun parent would
    otherwise leave its rank processes spinning on the GPUs (and every later measurement on the box polluted)."""
    import signal
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        raise AssertionError(f"timed out after {timeout}s\n" + out[-3000:] + err[-3000:])
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("world,ns", [(2, "20,12,10"), (2, "7,9,5"), (2, "3,6,5"), (4, "23,8,6"), (8, "40,6,6"), (8, "9,6,5")])
def test_slab_solve_multi_gpu(fv, world, ns, p2p):
    """p2p=1: halo planes and CG scalars over NVLink peer memory (CUDA IPC); p2p=0: NCCL."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, FV_NS=ns, FVB_P2P=p2p, FV_EXPECT_MG="0" if ns in ("3,6,5", "9,6,5") else "1")
    port = 29600 + (os.getpid() + world) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = _run_group(cmd, env, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout


# ---- one process, one thread, several GPUs: the fvb_multi front end (include/fvb200.h) -------------------------------
def _box(fv, ns, sigma=1.0):
    import math
    import numpy as np
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


@pytest.mark.parametrize("ndev,ns", [(2, [64, 64, 64]), (2, [7, 9, 5]), (3, [20, 12, 10]), (4, [23, 8, 6]), (8, [40, 6, 6]),
                                     (8, [9, 6, 5])])
def test_single_process_multi_gpu_regular(fv, orc, ndev, ns):
    """solvediffusion's whole-problem arrays in, one call, ndev GPUs: heads <= 1e-8 of the oracle at rtol 1e-12, the global
    CSR image bit-exact in structure, every device on the closed-form path.  (9 planes on 8 devices: the last one owns
    the Dirichlet plane only -- a rank without free rows.)"""
    import numpy as np
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    nb, aol, kf, src, dn, dh, lnk = _box(fv, ns)
    ms = fv.MultiSystem(list(range(ndev)))
    ms.assemble(nb, aol, kf, src, dn, dh, None, True)
    sz = ms.sizes()
    assert sz["node_ranges"][0][0] == 1 and sz["node_ranges"][-1][1] == src.size
    Ao = orc.assembleA(nb, aol, kf, src, dn, dh, None, True)
    bo = orc.assembleb(nb, aol, kf, src, dn, dh, None, True)
    p, i, v = ms.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.allclose(v, Ao.nzval, rtol=1e-14, atol=0)
    assert np.allclose(ms.b(), bo, rtol=1e-14, atol=0)
    head, x, ch = ms.solve(rtol=1e-12, want_x=True)
    ho, cho, *_ = orc.solvediffusion(nb, aol, kf, src, dn, dh, maxiter=50000, tol=1e-12, logtransformconductivity=True)
    assert ch.isconverged and np.max(np.abs(head - ho)) <= 1e-8 * np.max(np.abs(ho))
    assert abs(ch.iters - cho.iters) <= 3
    fn = ms.freenode()
    assert np.array_equal(head[fn], x) and np.array_equal(head[~fn], dh)
    kinds = [ms.device_system(r).assembly() for r in range(ndev) if ms.device_system(r).sizes()["nf_local"] >= 2]
    assert kinds and all(k == "box" for k in kinds)
    # the reference-named call with devices=[...]
    h2, ch2, A2, b2, fn2 = fv.solvediffusion(nb, aol, kf, src, dn, dh, rtol=1e-12, logtransformconductivity=True,
                                             devices=list(range(ndev)))
    assert np.array_equal(h2, head) and ch2.iters == ch.iters and np.array_equal(A2.rowval, Ao.rowval)
    # grid-implicit variant
    mi = fv.MultiSystem(list(range(ndev)))
    mi.assemble_regulargrid([0, 0, 0], [n - 1 for n in ns], ns, lnk, src, dn, dh)
    hi_, _, chi = mi.solve(rtol=1e-12)
    assert chi.isconverged and np.max(np.abs(hi_ - ho)) <= 1e-8 * np.max(np.abs(ho))
    # multigrid-preconditioned, chosen before assembly
    mg = fv.MultiSystem(list(range(ndev)))
    mg.set_preconditioner("mg")
    mg.assemble(nb, aol, kf, src, dn, dh, None, True)
    hm, _, chm = mg.solve(rtol=1e-12)
    assert chm.isconverged and np.max(np.abs(hm - ho)) <= 1e-8 * np.max(np.abs(ho))


@pytest.mark.parametrize("ndev", [2, 4])
def test_single_process_multi_gpu_irregular(fv, orc, fourfractures, ndev):
    """An irregular graph (the fourfractures fixture): equal node ranges, faces filtered on the host, CSR kernels,
    irregular halo lists."""
    import numpy as np
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    ff = fourfractures
    args = (ff["neighbors"], ff["areasoverlengths"], ff["conductivities"], np.zeros(ff["xs"].size), ff["dirichletnodes"],
            ff["dirichletheads"])
    ms = fv.MultiSystem(list(range(ndev))).assemble(*args)
    Ao = orc.assembleA(*args)
    p, i, v = ms.csr()
    assert np.array_equal(p, Ao.colptr) and np.array_equal(i, Ao.rowval) and np.array_equal(v, Ao.nzval)
    head, _, ch = ms.solve(rtol=1e-12)
    ho, cho, *_ = orc.solvediffusion(*args, maxiter=20000, tol=1e-12)
    assert ch.isconverged and np.max(np.abs(head - ho)) <= 1e-8 * np.max(np.abs(ho))


@pytest.mark.parametrize("ndev", [2, 8])
def test_c_abi_multi_demo(fv, tmp_path, ndev):
    """examples/c_abi_multi_demo.c: the same front end from plain C99, one process and one thread."""
    import shutil
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = str(tmp_path / "c_abi_multi_demo")
    libdir = os.path.dirname(fv.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_multi_demo.c"), "-o", exe, fv.LIB_PATH, f"-Wl,-rpath,{libdir}", "-lm"],
                   check=True)
    r = _run_group([exe, str(ndev), "64"], dict(os.environ), timeout=240)
    assert r.returncode == 0 and "c_abi_multi_demo ok" in r.stdout, r.stdout + r.stderr
