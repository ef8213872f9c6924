"""How many PDHG iterations do warm-started dive nodes need at the C4/C5 shapes? (exploration)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip, random_dive_bounds
from oracle.highs_lp import HighsLP, HIGHS_INF

n, m, dens, B, depth = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
d = numpy_random_mip(n, m, density=dens, seed=2)
lp = engine.BatchLP(d.A, d.b, d.c)
t = time.time()
r = lp.solve_batch(d.l[None], d.u[None], opts=engine.default_opts(max_iters=400000))
print('root gpu: status', r.status, 'obj', r.objective, 'iters', r.iterations, 'wall', time.time() - t, flush=True)
t = time.time()
h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u)
ref = h.solve()
print('root highs: obj', ref.objective, 'simplex iters', ref.iterations, 'wall', time.time() - t, flush=True)
print('rel err', abs(r.objective[0] - ref.objective) / abs(ref.objective))
x0, y0 = r.x[0], r.y[0]
lbs, ubs, _ = random_dive_bounds(d, x0, B, depth, seed=1)
for warm in (True, False):
    t = time.time()
    rr = lp.solve_batch(lbs, ubs, x0=np.tile(x0, (B, 1)) if warm else None, y0=np.tile(y0, (B, 1)) if warm else None,
                        opts=engine.default_opts(max_iters=200000), want_x=False, want_y=False)
    dt = time.time() - t
    it = rr.iterations
    print('warm' if warm else 'cold', 'wall', dt, 'stats', rr.stats, 'status', np.unique(rr.status, return_counts=True),
          'iters min/med/mean/max', it.min(), np.median(it), it.mean(), it.max(), flush=True)
# host simplex on the same nodes, warm started from the root basis
t = time.time()
k = min(B, 16)
objs = []
for i in range(k):
    h.set_col_bounds(lbs[i], ubs[i]); h.set_basis(ref.col_basis, ref.row_basis)
    s = h.solve(); objs.append(s.objective if s.status == 0 else np.inf)
dt = time.time() - t
print('highs warm per LP s', dt / k, 'iters last', s.iterations)
print('max rel err vs highs', max(abs(a - b) / max(1, abs(b)) for a, b in zip(rr.objective[:k], objs) if np.isfinite(b)))
