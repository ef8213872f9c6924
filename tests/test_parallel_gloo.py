"""Node sharding and the incumbent / dual-bound all-reduce on 2 gloo ranks (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from simple_mip_solver_b200 import parallel


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)
    assert parallel.allreduce_bounds(1.0, 2.0) == (1.0, 2.0)      # no process group: identity


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    nodes = list(range(10))
    mine = parallel.shard_items(nodes, rank, world)
    # each rank "solves" its slice: objective = node id + 0.5, nodes 3 and 8 are integral
    inc = min([k + 0.5 for k in mine if k in (3, 8)], default=float('inf'))
    low = min(k + 0.25 for k in mine)
    g_inc, g_low = parallel.allreduce_bounds(inc, low)
    t = parallel.allreduce_max(float(rank + 1))
    s = parallel.allreduce_sum([len(mine), 1.0])
    out.put((rank, mine, g_inc, g_low, t, s))
    dist.destroy_process_group()


def test_two_rank_exchange():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 1, 2, 3, 4] and got[1][1] == [5, 6, 7, 8, 9]
    for _, _, g_inc, g_low, t, s in got:
        assert g_inc == 3.5 and g_low == 0.25 and t == 2.0 and s == [10.0, 2.0]


class _FakeLP:
    """Stands in for engine.BatchLP: comm_init fails on the ranks listed in ``bad``, the
    precondition probe on the ranks listed in ``unready``."""
    def __init__(self, rank, bad, unready=()):
        self.rank, self.bad, self.unready, self.joined, self.left = rank, bad, unready, False, False
        self.entered_init = False

    def comm_probe(self):
        return self.rank not in self.unready

    def comm_init(self):
        self.entered_init = True
        if self.rank in self.bad:
            raise RuntimeError('libnccl.so.2 not found')
        self.joined = True
        return True

    def comm_destroy(self):
        self.left = True


def _comm_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    res = []
    for bad in ((), (1,), (0, 1)):
        lp = _FakeLP(rank, bad)
        msgs = []
        use = parallel.join_library_comm(lp, log=msgs.append)
        res.append((use, lp.joined, lp.left, len(msgs)))
    # a rank > 0 whose precondition fails (libnccl missing there): NOBODY may enter ncclCommInitRank
    lp = _FakeLP(rank, (), unready=(1,))
    use = parallel.join_library_comm(lp)
    res.append((use, lp.entered_init))
    out.put((rank, res))
    dist.destroy_process_group()


def test_all_ranks_or_none_use_the_library_communicator():
    assert parallel.join_library_comm(_FakeLP(0, ())) is False        # no process group: nothing to join
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank succeeds: both use it, nobody leaves
    assert got[0][0] == (True, True, False, 0) and got[1][0] == (True, True, False, 0)
    # rank 1 fails: rank 0 joined and must leave again, both fall back
    assert got[0][1] == (False, True, True, 0) and got[1][1] == (False, False, False, 1)
    # both fail
    assert got[0][2] == (False, False, False, 1) and got[1][2] == (False, False, False, 1)
    # rank 1 is not ready: no rank enters the collective initialisation
    assert got[0][3] == (False, False) and got[1][3] == (False, False)
