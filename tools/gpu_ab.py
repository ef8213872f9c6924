"""A/B of two builds of libblp.so (BLP_LIB) on one C5 frontier slice: iterations, time, objectives."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
wl = sys.argv[1] if len(sys.argv) > 1 else 'c5'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
d, depth, root = bench.load_instance(wl)
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False)
it = r.iterations
s = r.stats
print(json.dumps(dict(lib=os.environ.get('BLP_LIB', 'default'), mean_it=int(it.mean()), p50=int(np.median(it)), p90=int(np.percentile(it, 90)),
                      max_it=int(it.max()), total_ms=round(s['total_ms']), step_ms=round(s['step_kernel_ms']),
                      us_per_node_iter=round(1e3 * s['step_kernel_ms'] / s['node_iterations'], 5), unsolved=int((r.status != 0).sum()))), flush=True)
np.save(os.path.join('gpurun_out', 'ab_obj_%s.npy' % os.path.basename(os.environ.get('BLP_LIB', 'default'))), r.objective)
lp.close()
