"""Step-kernel time vs matrix density at the C5 shape (how much do the gathers cost?)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip

n, m, B = 50000, 20000, int(os.environ.get('SWEEP_B', '4096'))
ld = engine.leading_dim(B)
dev = torch.device('cuda', 0)
lb = torch.zeros((n, ld), dtype=torch.float64, device=dev)
ub = torch.full((n, ld), 10.0, dtype=torch.float64, device=dev)
for dens in [float(a) for a in sys.argv[1:]] or [1e-7, 5e-5, 2e-4, 1e-3]:
    d = numpy_random_mip(n, m, density=dens, seed=2)
    lp = engine.BatchLP(d.A, d.b, d.c)
    o = engine.default_opts(max_iters=128, eval_every=64, profile=1)
    r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
    r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
    s = r['stats']
    pm, dm = s['primal_kernel_ms'] / s['iterations'], s['dual_kernel_ms'] / s['iterations']
    pb = 8 * B * (3 * n + m) / 1e9
    db = 8 * B * (n + 2 * m) / 1e9
    print(f'density {dens:g} nnz {d.A.nnz}: primal {pm:.3f} ms ({pb / pm * 1e3:.0f} GB/s algo)  dual {dm:.3f} ms ({db / dm * 1e3:.0f} GB/s algo)', flush=True)
    lp.close()
