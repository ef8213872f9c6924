"""Restart / primal-weight constants vs PDHG iterations on one C5 slice (env knobs read per call)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _tuning  # noqa: F401  (tuning build of libblp.so: reads the BLP_* variables below)
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
B = 256
wl = sys.argv[1] if len(sys.argv) > 1 else 'c5'
d, depth, root = bench.load_instance(wl)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
lp = engine.BatchLP(d.A, d.b, d.c)
cases = [{}, {'BLP_OMEGA_THETA': '0'}, {'BLP_BETA_ART': '0.7'}, {'BLP_BETA_ART': '1.0'}, {'BLP_BETA_ART': '3.0'},
         {'BLP_OMEGA_THETA': '0', 'BLP_BETA_ART': '0.5'}, {'BLP_OMEGA_THETA': '0', 'BLP_BETA_ART': '0.7'},
         {'BLP_OMEGA_THETA': '0', 'BLP_BETA_ART': '1.0'}]
if os.environ.get('SWEEP_BALANCE'):      # residual-balancing feedback on the primal weight (k_decide)
    cases = [{}, {'BLP_OMEGA_THETA': '0'}, {'BLP_OMEGA_THETA': '0', 'BLP_OMEGA_BALANCE': '0.3'},
             {'BLP_OMEGA_THETA': '0', 'BLP_OMEGA_BALANCE': '0.1'}, {'BLP_OMEGA_BALANCE': '0.3'},
             {'BLP_OMEGA_THETA': '0.02'}]
if os.environ.get('SWEEP_BALANCE') == '2':  # dead zone of the feedback
    cases = [{'BLP_OMEGA_DEADZONE': '0'}, {'BLP_OMEGA_DEADZONE': '0.5'}, {'BLP_OMEGA_DEADZONE': '0.25'}]
if os.environ.get('SWEEP_COLD'):
    x0 = y0 = None
for env in cases:
    os.environ.update(env)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False)
    it = r.iterations
    print(env or 'defaults', 'mean', int(it.mean()), 'p50', int(np.median(it)), 'p90', int(np.percentile(it, 90)), 'max', int(it.max()),
          'total_ms', int(r.stats['total_ms']), 'unsolved', int((r.status == 3).sum()), flush=True)
    for k in env:
        os.environ.pop(k)
lp.close()
