"""world_size-2 gloo run (CPU) of the collective setup code the multi-GPU path uses:
unique-id broadcast and halo-plan exchange over torch.distributed.  The System is replaced
by a stub with the same three methods, so no GPU is needed; the exchanged plan is checked
against the pure planner and for mutual consistency between the two ranks."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StubSystem:
    """sizes() / halo_cols() / set_halo_plan() of fvb200.System for a 1-D chain split in two."""

    def __init__(self, rank, n_per_rank):
        self.rank, self.n = rank, n_per_rank
        self.plan = None
        self.uid = None

    def sizes(self):
        return dict(row_start=1 + self.rank * self.n, nf_local=self.n)

    def halo_cols(self):
        # rank 0 references the first two rows of rank 1; rank 1 the last row of rank 0
        return np.array([self.n + 1, self.n + 2] if self.rank == 0 else [self.n], np.int64)

    def set_halo_plan(self, peers, send_counts, send_rows, recv_counts):
        self.plan = (list(peers), list(send_counts), np.asarray(send_rows), list(recv_counts))

    @staticmethod
    def unique_id():
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        self.uid = (world, rank, uid)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    g.load_package()
    import importlib
    d = importlib.import_module("fvb200.distributed")
    s = StubSystem(rank, 6)
    d.init_comm(s)
    plan = d.exchange_halo_plan(s)
    assert plan[0] == s.plan[0]
    # same partition again: the cached plan (fixed-size record path) must be identical
    plan2 = d.exchange_halo_plan(s)
    assert plan2[0] == plan[0] and plan2[1] == plan[1] and np.array_equal(plan2[2], plan[2]) and plan2[3] == plan[3]
    # a halo with more than two runs on one rank sends every rank down the pickled-object path
    class Ragged(StubSystem):
        def halo_cols(self):
            return np.array([self.n + 1, self.n + 3, self.n + 5] if self.rank == 0 else [self.n], np.int64)
    s3 = Ragged(rank, 6)
    plan3 = d.exchange_halo_plan(s3)
    assert plan3[3] == ([3] if rank == 0 else [1]) and plan3[1] == ([1] if rank == 0 else [3])
    q.put((rank, s.plan[0], s.plan[1], s.plan[2].tolist(), s.plan[3], s.uid[2] == bytes(range(128)), s.uid[:2]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_halo_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] == [1] and r0[2] == [1] and r0[3] == [5] and r0[4] == [2]   # sends its last row, receives two
    assert r1[1] == [0] and r1[2] == [2] and r1[3] == [0, 1] and r1[4] == [1]
    assert r0[5] and r1[5] and r0[6] == (2, 0) and r1[6] == (2, 1)
