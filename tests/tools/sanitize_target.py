"""Small workloads that touch every kernel family (written as a compute-sanitizer target; the tool is
closed on this GPU pool, so it serves as a plain coverage run: every family launches, statuses are checked):
  A  one-node-per-lane kernels (cooperative period included), continuous batching, row masks, a dense
     cut row (cooperative long-row path), warm starts, most-fractional index
  B  two-nodes-per-lane kernels in two parallel graph branches, continuous batching, compaction, a
     dense cut row, batched SpMV
Run:  python tests/tools/sanitize_target.py        (or under compute-sanitizer --tool memcheck where allowed)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes, numpy_random_mip

rng = np.random.default_rng(0)


def run(tag, n, m, dens, B, slots, cut_nnz, **kw):
    d = numpy_random_mip(n, m, density=dens, seed=3)
    lbs, ubs, _ = frontier_nodes(d, np.full(n, 0.5), 0, B, 5, seed=1, p_down=0.9)
    lp = engine.BatchLP(d.A, d.b, d.c)
    rows = np.zeros((2, n))
    rows[0, rng.choice(n, size=cut_nnz, replace=False)] = -1.0
    rows[1, rng.choice(n, size=5, replace=False)] = -1.0
    lp.append_rows(rows, np.array([-0.4 * cut_nnz, -3.0]))
    masks = (rng.random((B, 2)) < 0.5).astype(np.uint8)
    r = lp.solve_batch(lbs, ubs, row_mask=masks, integer_indices=list(range(0, n, 2)),
                       opts=engine.default_opts(max_active=slots, **kw))
    r2 = lp.solve_batch(lbs, ubs, row_mask=masks, x0=r.x, y0=r.y, opts=engine.default_opts(**{**kw, 'max_iters': 128}))
    X = torch.randn((n, engine.leading_dim(B)), dtype=torch.float64, device='cuda')
    Y = lp.spmv_device(X, transpose=False, B=B)
    lp.spmv_device(Y, transpose=True, B=B)
    print(tag, 'status', dict(zip(*np.unique(r.status, return_counts=True))), 'refills', r.stats['refills'],
          'compactions', r.stats['compactions'], 'launches', r.stats['kernel_launches'], 'warm rerun', int((r2.status == 0).sum()), flush=True)
    lp.close()


run('A', 400, 200, 0.03, 90, 32, 150)
run('B', 24000, 9000, 3e-4, 192, 128, 300, eps_rel=1e-4, max_iters=4000)
print('sanitize target done')
