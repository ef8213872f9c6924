"""Restart / primal-weight constants vs PDHG iterations on one C5 slice (env-tunable constants)."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _tuning  # noqa: F401  (tuning build of libblp.so: reads the BLP_* variables below)
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
B = 256
d, depth, root = bench.load_instance(sys.argv[1] if len(sys.argv) > 1 else 'c5')
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
for scale in ('1.0', '0.25', '0.5', '2.0', '4.0'):
    os.environ['BLP_OMEGA0_SCALE'] = scale          # read when the handle is created
    lp = engine.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False)
    it = r.iterations
    print('omega0 x', scale, 'mean', int(it.mean()), 'p50', int(np.median(it)), 'p90', int(np.percentile(it, 90)), 'max', int(it.max()),
          'total_ms', int(r.stats['total_ms']), 'unsolved', int((r.status == 3).sum()), flush=True)
    lp.close()
