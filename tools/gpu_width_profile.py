"""Where does a bench step spend its time? Period time vs batch width (verbose log of one C5 slice)."""
import os, sys, re, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d, depth, root = bench.load_instance('c5')
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, opts=engine.default_opts(verbose=1, max_active=W, eval_every=int(os.environ.get('EVAL', '64'))), want_x=False, want_y=False)
print('total_ms', r.stats['total_ms'], 'step_ms', r.stats['step_kernel_ms'], 'iters', r.stats['iterations'], 'compactions', r.stats['compactions'])
