"""Restart / primal-weight constants vs PDHG iterations on one C5 slice (env-tunable constants)."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
B = 256
d, depth, root = bench.load_instance(sys.argv[1] if len(sys.argv) > 1 else 'c5')
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
configs = [dict(BLP_OMEGA_THETA=t, BLP_BETA_ART=a) for t in ('0.0', '0.05', '0.1', '0.2', '0.3') for a in ('0.36', '0.2')]
for cfg in configs:
    for k in ('BLP_BETA_SUFF', 'BLP_BETA_NEC', 'BLP_BETA_ART', 'BLP_OMEGA_THETA'):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False)
    it = r.iterations
    print(cfg, 'mean', int(it.mean()), 'p50', int(np.median(it)), 'p90', int(np.percentile(it, 90)), 'max', int(it.max()),
          'total_ms', int(r.stats['total_ms']), 'unsolved', int((r.status == 3).sum()), flush=True)
