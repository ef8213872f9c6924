"""Short run of the step kernels at the C5 shape for ncu (a few PDHG periods, no CUDA graph)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip

n, m, dens, B, iters = 50000, 20000, 2e-4, int(os.environ.get('NCU_B', '2048')), int(os.environ.get('NCU_ITERS', '128'))
d = numpy_random_mip(n, m, density=dens, seed=2)
lp = engine.BatchLP(d.A, d.b, d.c)
ld = engine.leading_dim(B)
dev = torch.device('cuda', 0)
lb = torch.zeros((n, ld), dtype=torch.float64, device=dev)
ub = torch.full((n, ld), 10.0, dtype=torch.float64, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
idx = torch.randint(0, n, (16, ld), device=dev, generator=g)
ub.scatter_(0, idx, torch.randint(0, 3, (16, ld), device=dev, generator=g).double())
o = engine.default_opts(max_iters=iters, eval_every=64, use_graph=0)
r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
print('ok', r['stats'])
