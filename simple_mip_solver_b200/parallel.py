"""Multi-GPU plumbing of the bound step: node sharding and the incumbent / dual-bound exchange.

Node LPs are independent (they share read-only A, c, b), so a frontier is split by node across
ranks — one process per GPU, each with its own ``engine.BatchLP`` replica of the matrix — and the
data path needs no collective. The only exchange is a 16-byte all-reduce(min) of
``[incumbent objective, smallest open-node lower bound]`` after a batch: on GPUs the library's own
``blp_allreduce_min`` (NCCL over NVLink on the handle's stream, communicator bootstrapped through the
``torch.distributed`` group), in the CPU tests ``torch.distributed`` with gloo. The reference is single
process; this module has no counterpart there.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of ``total`` nodes owned by ``rank``; sizes differ by <= 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank/world: {rank}/{world}')
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_items(items: Sequence, rank: int, world: int) -> List:
    b, e = shard_bounds(len(items), rank, world)
    return list(items[b:e])


def allreduce_bounds(incumbent: float, dual_bound: float, device=None, lp=None) -> Tuple[float, float]:
    """Global (min incumbent objective, min open-node lower bound) over all ranks.

    With ``lp`` (an ``engine.BatchLP`` whose ``comm_init()`` has been called) the exchange is the
    library's own 16-byte ncclAllReduce on its stream (``blp_allreduce_min``); otherwise it goes
    through ``torch.distributed`` (gloo in the CPU tests). Without an initialised process group
    (single GPU) the inputs are returned unchanged."""
    if lp is not None:
        return lp.allreduce_min(incumbent, dual_bound)
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(incumbent), float(dual_bound)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
            else torch.device('cpu')
    t = torch.tensor([incumbent, dual_bound], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    a, b = t.tolist()
    return float(a), float(b)


def join_library_comm(lp, device=None, log=None) -> bool:
    """Give ``lp`` (an ``engine.BatchLP``) the library's own NCCL communicator on every rank, or on none.

    ``ncclCommInitRank`` is itself a collective, so the ranks first agree — one all-reduce through the
    process group — that EVERY rank can enter it (``lp.comm_probe()``: libnccl bound, no communicator
    yet, device selectable); a rank whose precondition fails never leaves its peers waiting inside the
    initialisation. Then each rank calls ``lp.comm_init()`` and the ranks agree once more on whether
    all of them succeeded; if not, those that did leave the communicator again and everybody uses
    ``torch.distributed`` for the 16-byte exchange.
    Returns True when ``allreduce_bounds(..., lp=lp)`` may be used. Single rank: False, no-op."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    world = dist.get_world_size()
    can = False
    try:
        can = bool(lp.comm_probe())
    except Exception as e:           # noqa: BLE001
        if log:
            log(f'[rank {dist.get_rank()}] blp_comm_probe failed ({e})')
    if allreduce_sum([1.0 if can else 0.0], device=device)[0] < world:
        if log and not can:
            log(f'[rank {dist.get_rank()}] cannot join the library communicator; all ranks use torch.distributed')
        return False
    ok = False
    try:
        ok = bool(lp.comm_init())
    except Exception as e:           # noqa: BLE001 - reported, then agreed on by all ranks
        if log:
            log(f'[rank {dist.get_rank()}] blp_comm_init failed ({e}); using torch.distributed')
    agreed = allreduce_sum([1.0 if ok else 0.0], device=device)[0]
    if ok and agreed < world:
        lp.comm_destroy()
    return agreed == world


def allreduce_max(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
            else torch.device('cpu')
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(values: Sequence[float], device=None) -> List[float]:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
            else torch.device('cpu')
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
