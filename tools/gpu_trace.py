"""Convergence trace of a few node LPs (verbose=2): KKT errors, restarts and primal weight per evaluation."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip, frontier_nodes
import bench
name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
d, depth, root = bench.load_instance(name)
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, 4, depth, seed=0)
B = 4
o = engine.default_opts(max_iters=int(os.environ.get('ITERS', '400000')), verbose=2, eval_every=int(os.environ.get('EVAL', '64')))
r = lp.solve_batch(lbs, ubs, x0=np.tile(root['x'], (B, 1)), y0=np.tile(root['y'], (B, 1)), opts=o, want_x=False, want_y=False)
print('iters', r.iterations, 'status', r.status, 'obj', r.objective)
