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
    """Run a command in its own session with stdout/stderr in files and, on timeout, kill it AND every descendant
    (torchrun's rank processes live in sessions of their own): rank processes left spinning on the GPUs would pollute
    every later measurement on the box, and open pipes would block the test forever."""
    import signal
    import tempfile
    import psutil
    with tempfile.TemporaryFile("w+") as fo, tempfile.TemporaryFile("w+") as fe:
        p = subprocess.Popen(cmd, env=env, stdout=fo, stderr=fe, text=True, start_new_session=True)
        timed_out = False
        try:
            p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            timed_out = True
            try:
                kids = psutil.Process(p.pid).children(recursive=True)
            except psutil.Error:
                kids = []
            for k in kids:
                try:
                    k.kill()
                except psutil.Error:
                    pass
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except OSError:
                pass
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                pass
        fo.seek(0); fe.seek(0)
        out, err = fo.read(), fe.read()
    if timed_out:
        raise AssertionError(f"timed out after {timeout}s (process tree killed)\n" + out[-3000:] + err[-3000:])
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
    r = _run_group(cmd, env, timeout=150)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout


# ---- one process, one thread, several GPUs: the fvb_multi front end (include/fvb200.h) -------------------------------
# The bodies live in tests/multi_inproc_worker.py and run in a child process with a timeout: a deadlock between the
# devices of an in-process solve must cost a bounded amount of GPU time, not the whole pytest session.
def _inproc(case, ndev, arg, timeout=150):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "multi_inproc_worker.py"), case, str(ndev), arg]
    r = _run_group(cmd, dict(os.environ), timeout=timeout)
    assert r.returncode == 0 and "inproc ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("ndev,ns", [(2, "64,64,64"), (2, "7,9,5"), (3, "20,12,10"), (4, "23,8,6"), (8, "40,6,6"), (8, "9,6,5")])
def test_single_process_multi_gpu_regular(fv, ndev, ns):
    """solvediffusion's whole-problem arrays in, one call, ndev GPUs: heads <= 1e-8 of the oracle at rtol 1e-12, the global
    CSR image bit-exact in structure, every device on the closed-form path; the reference-named call with devices=[...];
    the grid-implicit variant; multigrid chosen before assembly.  (9 planes on 8 devices: a rank without free rows.)"""
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    _inproc("regular", ndev, ns)


@pytest.mark.parametrize("ndev", [2, 4])
def test_single_process_multi_gpu_irregular(fv, ndev):
    """An irregular graph (the fourfractures fixture): equal node ranges, faces filtered on the host, CSR kernels,
    irregular halo lists."""
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    _inproc("irregular", ndev, "-")


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
    r = _run_group([exe, str(ndev), "64"], dict(os.environ), timeout=150)
    assert r.returncode == 0 and "c_abi_multi_demo ok" in r.stdout, r.stdout + r.stderr
