"""One-rank-per-GPU plumbing for slab-partitioned problems.

Only setup-time, host-side logic lives here: choosing contiguous node ranges, exchanging the
NCCL unique id, and turning each rank's list of referenced off-rank columns into the
send/receive lists the library needs.  It uses torch.distributed collectives on small host
objects, so it runs unchanged on gloo (CPU tests) and on a NCCL job.  The per-iteration
traffic (halo planes, CG scalars) is NCCL inside libfvb200.so, not here.
"""
from __future__ import annotations

import numpy as np


def slab_planes(n1: int, nranks: int, dirichlet_end_planes: bool = True):
    """Split x-planes 1..n1 into `nranks` contiguous slabs, balancing FREE planes: with the
    usual left/right Dirichlet faces (examples/box_model/ex.jl:27-37) planes 1 and n1 carry no
    unknowns, so the end ranks take one plane more.  Returns [(p_lo, p_hi)] 1-based inclusive."""
    if nranks < 1 or n1 < nranks:
        raise ValueError("need at least one plane per rank")
    fixed = 2 if (dirichlet_end_planes and n1 >= 2 + nranks) else 0
    free = n1 - fixed
    base, extra = divmod(free, nranks)
    counts = [base + (1 if r < extra else 0) for r in range(nranks)]
    if fixed:
        counts[0] += 1
        counts[-1] += 1
    out, lo = [], 1
    for c in counts:
        out.append((lo, lo + c - 1))
        lo += c
    assert lo == n1 + 1
    return out


def node_range_of_planes(planes, n2, n3):
    """Planes (p_lo, p_hi) -> owned node range (lo, hi), 1-based inclusive (src/grid.jl:60)."""
    return (planes[0] - 1) * n2 * n3 + 1, planes[1] * n2 * n3


def halo_plan_from_ranges(rank, row_ranges, halo_cols_by_rank):
    """Pure function (no communication): given every rank's (row_start, nf_local) (1-based
    start) and every rank's ascending list of referenced off-rank global columns, return for
    `rank`: peers, send_counts, send_rows (0-based local), recv_counts.  Halo entries arrive
    grouped by owner in ascending rank order, which is ascending column order because ranks
    own ascending row ranges."""
    starts = np.array([r[0] for r in row_ranges], np.int64)
    ends = starts + np.array([r[1] for r in row_ranges], np.int64)  # exclusive

    def owner(cols):
        o = np.searchsorted(starts, cols, side="right") - 1
        if cols.size and (np.any(o < 0) or np.any(cols >= ends[o])):
            raise ValueError("halo column outside every rank's row range")
        return o

    mine = np.asarray(halo_cols_by_rank[rank], np.int64)
    if mine.size and np.any(np.diff(mine) <= 0):
        raise ValueError("halo columns must be strictly ascending")
    own_mine = owner(mine)
    if np.any(own_mine == rank):
        raise ValueError("a halo column is owned by the requesting rank")
    send = {}
    for q, cols in enumerate(halo_cols_by_rank):
        if q == rank:
            continue
        cols = np.asarray(cols, np.int64)
        sel = cols[owner(cols) == rank]
        if sel.size:
            send[q] = (sel - starts[rank]).astype(np.int32)
    recv = {int(p): int(np.count_nonzero(own_mine == p)) for p in np.unique(own_mine)}
    peers = sorted(set(send) | set(recv))
    send_counts = [int(send[p].size) if p in send else 0 for p in peers]
    recv_counts = [recv.get(p, 0) for p in peers]
    send_rows = np.concatenate([send[p] for p in peers if p in send]) if send else np.empty(0, np.int32)
    return peers, send_counts, send_rows, recv_counts


def send_destinations(rank, peers, row_ranges, halo_cols_by_rank):
    """For each peer of `rank`'s plan: the index in the PEER's vector where this rank's first halo
    value belongs = peer's nf_local + number of the peer's halo columns owned by lower ranks
    (halo columns ascend, ranks own ascending row ranges)."""
    my_start = row_ranges[rank][0]
    return [int(row_ranges[p][1] + np.searchsorted(np.asarray(halo_cols_by_rank[p], np.int64), my_start))
            for p in peers]


def runs_of(cols):
    """Ascending integer list -> [(start, length)] of its maximal runs of consecutive values.  Slab partitions
    reference whole planes of the neighbour, so a halo list of millions of columns is one or two runs; the ranks
    exchange the runs instead of the lists (4 MB per rank at 512^3 otherwise, pickled and gathered every step)."""
    c = np.asarray(cols, np.int64)
    if c.size == 0:
        return []
    brk = np.flatnonzero(np.diff(c) != 1) + 1
    starts = np.concatenate([[0], brk])
    ends = np.concatenate([brk, [c.size]])
    return [(int(c[a]), int(b - a)) for a, b in zip(starts, ends)]


def cols_of(runs):
    """Inverse of runs_of."""
    if not runs:
        return np.empty(0, np.int64)
    return np.concatenate([np.arange(s, s + n, dtype=np.int64) for s, n in runs])


_PLAN_CACHE = {}


def exchange_halo_plan(system, group=None, peer_memory=None):
    """Collective: gather row ranges and halo lists over torch.distributed and install the plan.
    peer_memory (default: on unless FVB_P2P=0): also map the neighbours' vectors and mailboxes
    through CUDA IPC so the per-iteration exchanges run over NVLink peer memory instead of NCCL.

    Wire format: one fixed-size int64[8] record per rank (row_start, nf_local, number of halo runs, two runs as
    (start, length)) all-gathered as a tensor -- a slab's halo is one run per neighbour -- and the 128-byte peer
    blobs as a uint8 tensor; only partitions whose halo has more than two runs fall back to pickled objects.
    The plan computed from a set of records is cached per system, so re-assembling the same partition (the
    per-step path of an inverse loop or of bench.py) costs two small all-gathers and no host-side planning."""
    import os

    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    s = system.sizes()
    runs = runs_of(system.halo_cols())
    rec = [int(s["row_start"]), int(s["nf_local"]), len(runs), 0, 0, 0, 0, 1]
    if len(runs) <= 2:
        for k, (st, ln) in enumerate(runs):
            rec[3 + 2 * k], rec[4 + 2 * k] = st, ln
    else:
        rec[7] = 0
    mine_t = torch.tensor(rec, dtype=torch.int64)
    all_t = [torch.empty(8, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_t, mine_t, group=group)
    recs = [tuple(int(v) for v in t) for t in all_t]
    if all(r[7] == 1 for r in recs):
        key = tuple(recs)
        cached = _PLAN_CACHE.get(id(system))
        if cached is not None and cached[0] == key:
            ranges, halos, plan = cached[1]
        else:
            ranges = [(r[0], r[1]) for r in recs]
            halos = [cols_of([(r[3 + 2 * k], r[4 + 2 * k]) for k in range(r[2])]) for r in recs]
            plan = halo_plan_from_ranges(rank, ranges, halos)
            _PLAN_CACHE[id(system)] = (key, (ranges, halos, plan))
    else:  # irregular partition somewhere: every rank takes the object path
        gathered = [None] * world
        dist.all_gather_object(gathered, (int(s["row_start"]), int(s["nf_local"]), runs), group=group)
        ranges = [(g[0], g[1]) for g in gathered]
        halos = [cols_of(g[2]) for g in gathered]
        plan = halo_plan_from_ranges(rank, ranges, halos)
    system.set_halo_plan(*plan)
    if peer_memory is None:
        peer_memory = os.environ.get("FVB_P2P", "1") != "0"
    if peer_memory and world > 1 and hasattr(system, "peer_export"):
        blob = torch.frombuffer(bytearray(system.peer_export()), dtype=torch.uint8)
        blobs_t = [torch.empty(blob.numel(), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(blobs_t, blob, group=group)
        blobs = [bytes(t.numpy().tobytes()) for t in blobs_t]
        system.peer_import(blobs, send_destinations(rank, plan[0], ranges, halos))
    return plan


def init_comm(system, group=None):
    """Collective: rank 0 creates the NCCL unique id, everyone joins (fvb_comm_init)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return
    box = [type(system).unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    system.comm_init(world, rank, box[0])
