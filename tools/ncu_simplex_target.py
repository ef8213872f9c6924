"""Short run of the dual simplex kernel for ncu: config 3 root (one CTA) and its 128 children (128 CTAs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import grumpy_random_mip
d = grumpy_random_mip(500, 300, density=0.1, rand_seed=2)
lp = engine.BatchLP(d.A, d.b, d.c)
root = lp.simplex_batch(d.l[None], d.u[None])
x = root.x[0]
frac = np.minimum(x - np.floor(x), np.ceil(x) - x)
cand = np.argsort(-frac, kind='stable')[:64]
deltas = [[(int(j), float(d.l[j]), float(np.floor(x[j])))] for j in cand] + [[(int(j), float(np.ceil(x[j])), float(d.u[j]))] for j in cand]
k = lp.simplex_children(d.l, d.u, deltas, col_status=root.col_status[0], row_status=root.row_status[0], parent_slot=0)
print('ok', root.stats, k.stats)
