"""Parallel graph branches over node-tile groups (BLP_GRAPH_LANES): time per PDHG iteration inside
the period graph at the C5 shape for a few batch widths, and bit-equality of the results."""
import os, sys, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _tuning  # noqa: F401  (tuning build of libblp.so: reads the BLP_* variables below)
import torch
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes

d, depth, root = bench.load_instance('c5')
n, m = d.n, d.m
dev = torch.device('cuda', 0)
widths = [int(a) for a in sys.argv[1:]] or [128, 512, 1024]
lp = engine.BatchLP(d.A, d.b, d.c)
pbn, dbn = 8 * (3 * n + m), 8 * (n + 2 * m)
for B in widths:
    ld = engine.leading_dim(B)
    lbs, ubs, _ = frontier_nodes(d, root['x'], 0, min(B, 512), depth, seed=0)
    reps = (B + lbs.shape[0] - 1) // lbs.shape[0]
    lb = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    ub = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    lb[:, :B] = torch.from_numpy(np.tile(lbs, (reps, 1))[:B]).to(dev).T
    ub[:, :B] = torch.from_numpy(np.tile(ubs, (reps, 1))[:B]).to(dev).T
    x0 = torch.from_numpy(root['x']).to(dev)[:, None].expand(n, ld).contiguous()
    y0 = torch.from_numpy(root['y']).to(dev)[:, None].expand(m, ld).contiguous()
    ref = None
    for lanes in os.environ.get('SWEEP_LANES', '1,2,3,4').split(','):
        os.environ['BLP_GRAPH_LANES'] = lanes
        og = engine.default_opts(max_iters=1024, eval_every=64)
        lp.solve_batch_device(lb, ub, x0=x0, y0=y0, opts=og, want_x=False, want_y=False)
        r = lp.solve_batch_device(lb, ub, x0=x0, y0=y0, opts=og, want_x=True, want_y=False)
        g = r['stats']
        sig = hashlib.sha1(r['lower'][:B].cpu().numpy().tobytes() + r['x'][:, :B].cpu().numpy().tobytes()).hexdigest()[:12]
        ref = ref or sig
        it_us = 1e3 * g['step_kernel_ms'] / g['iterations']
        print(f'B {B} lanes {lanes}: graph {it_us:.1f} us/iter ({(pbn+dbn)*B/it_us/1e3:.0f} GB/s)  result {sig} '
              f'{"same" if sig == ref else "DIFFERENT"}', flush=True)
lp.close()
