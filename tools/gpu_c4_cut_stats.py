"""Config 4: where does the time of a cut round go? Per-call stats with 0 and 32 appended dense rows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes
d, depth, root4 = bench.load_instance('c4')
B = 256
lbs, ubs, dl = frontier_nodes(d, root4['x'], 0, B, depth, seed=0)
lp = engine.BatchLP(d.A, d.b, d.c)
rng = np.random.default_rng(5)
x0 = np.tile(root4['x'], (B, 1)); y0 = np.tile(root4['y'], (B, 1))
def show(tag, r):
    s = r.stats
    print(tag, 'batch iters', s['iterations'], 'mean', int(r.iterations.mean()), 'max', int(r.iterations.max()), 'total_ms', int(s['total_ms']),
          'step_ms', int(s['step_kernel_ms']), 'us/iter', round(1e3 * s['step_kernel_ms'] / s['iterations'], 1),
          'ns per node-iter', round(1e6 * s['step_kernel_ms'] / s['node_iterations'], 1), 'compactions', s['compactions'], flush=True)
res = lp.solve_batch(lbs, ubs, x0=x0, y0=y0)
res = lp.solve_batch(lbs, ubs, x0=x0, y0=y0); show('no cuts', res)
prof = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, opts=engine.default_opts(profile=1, max_iters=512), want_x=False, want_y=False).stats
print('   full width per launch: primal', round(1e3 * prof['primal_kernel_ms'] / prof['iterations'], 1), 'us  dual', round(1e3 * prof['dual_kernel_ms'] / prof['iterations'], 1), 'us')
new_rows, new_rhs = [], []
for k in range(32):
    S = rng.choice(d.n, size=400, replace=False)
    row = np.zeros(d.n); row[S] = -1.0
    new_rows.append(row); new_rhs.append(-np.floor(res.x[k % B][S].sum()))
lp.append_rows(np.array(new_rows), np.array(new_rhs))
masks = (rng.random((B, 32)) < 0.25).astype(np.uint8)
y1 = np.hstack([res.y, np.zeros((B, 32))])
r2 = lp.solve_batch(lbs, ubs, row_mask=masks, x0=res.x, y0=y1); show('32 cut rows', r2)
prof = lp.solve_batch(lbs, ubs, row_mask=masks, x0=res.x, y0=y1, opts=engine.default_opts(profile=1, max_iters=512), want_x=False, want_y=False).stats
print('   full width per launch: primal', round(1e3 * prof['primal_kernel_ms'] / prof['iterations'], 1), 'us  dual', round(1e3 * prof['dual_kernel_ms'] / prof['iterations'], 1), 'us')
r3 = lp.solve_batch(lbs, ubs, row_mask=np.zeros_like(masks), x0=x0, y0=np.hstack([y0, np.zeros((B, 32))])); show('32 cut rows, all masked off, root warm start', r3)
lp.close()
