"""Does a node's dive depth predict its PDHG iteration count? (ordering heuristic for continuous batching)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
B = 512
d, depth, root = bench.load_instance('c5')
lp = engine.BatchLP(d.A, d.b, d.c)
lbs, ubs, deltas = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1)); y0 = np.tile(root['y'], (B, 1))
r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False)
it = r.iterations.astype(float)
dep = np.array([len(t) for t in deltas], float)
ups = np.array([sum(1 for (j, lo, hi) in t if lo > d.l[j]) for t in deltas], float)
# distance of the clipped root point from the root point, and objective change
move = np.array([np.abs(np.clip(root['x'], lbs[k], ubs[k]) - root['x']).sum() for k in range(B)])
dobj = r.objective - root['objective']
for name, v in (('depth', dep), ('up-branches', ups), ('clip distance', move), ('objective change', dobj)):
    print(f'{name:18s} corr with iterations {np.corrcoef(v, it)[0, 1]: .3f}   rank corr {np.corrcoef(np.argsort(np.argsort(v)), np.argsort(np.argsort(it)))[0, 1]: .3f}')
np.save('gpurun_out/iters_depth.npy', np.stack([it, dep, ups, move, dobj]))
