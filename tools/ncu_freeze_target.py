"""Short warm-started run of the step kernels on the C5 frontier for ncu (three PDHG periods, no CUDA graph):
the launches of the second period run with the frozen sets and folded matrices the first evaluation produced."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from simple_mip_solver_b200 import engine                       # noqa: E402
from simple_mip_solver_b200.instances import frontier_nodes     # noqa: E402

B = int(os.environ.get('NCU_B', '512'))
d, depth, root = bench.load_instance(os.environ.get('NCU_WL', 'c5'))
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
o = engine.default_opts(max_iters=192, eval_every=64, use_graph=0, freeze=int(os.environ.get('NCU_FREEZE', '1')))
r = lp.solve_batch(lbs, ubs, x0=np.tile(root['x'], (B, 1)), y0=np.tile(root['y'], (B, 1)), want_x=False, want_y=False, opts=o)
print('ok', r.stats)
