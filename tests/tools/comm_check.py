"""2+ GPUs under torchrun: the library's own NCCL exchange (blp_comm_init / blp_allreduce_min)
against torch.distributed's all-reduce on the same values."""
import os, sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from simple_mip_solver_b200 import engine, parallel

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
lp = engine.BatchLP(sp.eye(3, format='csr'), np.zeros(3), np.ones(3), device=local)
assert lp.comm_init()
for k in range(5):
    inc = float('inf') if (rank + k) % 3 == 0 else 10.0 * rank + k
    low = -1.5 * rank - k
    got = parallel.allreduce_bounds(inc, low, device=dev, lp=lp)
    ref = parallel.allreduce_bounds(inc, low, device=dev)
    assert got == ref, (rank, k, got, ref)
# the exchange also works right after a solve on the same stream
r = lp.solve_batch(np.zeros((2, 3)), np.ones((2, 3)))
assert (r.status == 0).all()
assert lp.allreduce_min(float(rank), float(-rank)) == (0.0, float(-(world - 1)))
lp.close()
dist.barrier()
if rank == 0:
    print('comm check ok on', world, 'ranks')
dist.destroy_process_group()
