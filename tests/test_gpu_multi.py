"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): slab partition, NCCL halo
exchange and all-reduced CG scalars must reproduce the oracle's heads within 1e-8 relative."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("world,ns", [(2, "20,12,10"), (2, "7,9,5"), (4, "23,8,6"), (8, "40,6,6")])
def test_slab_solve_multi_gpu(fv, world, ns, p2p):
    """p2p=1: halo planes and CG scalars over NVLink peer memory (CUDA IPC); p2p=0: NCCL."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, FV_NS=ns, FVB_P2P=p2p)
    port = 29600 + (os.getpid() + world) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout
