"""C4 root (10 000 x 5 000) on the wide dual simplex: pivots, flips, time; children from the stored factor."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip
n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 5000)
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 2e-3
d = numpy_random_mip(n, m, density=dens, seed=2)
lp = engine.BatchLP(d.A, d.b, d.c)
for limit in (200, 2147483647):
    t = time.perf_counter()
    r = lp.simplex_batch(d.l[None], d.u[None], max_pivots=limit)
    dt = time.perf_counter() - t
    print(f'limit {limit}: status {r.status[0]} obj {r.objective[0]:.6f} pivots {r.pivots[0]} flips {r.stats["refills"]} '
          f'kernel {r.stats["step_kernel_ms"]:.1f} ms wall {dt:.2f} s -> {1e3 * r.stats["step_kernel_ms"] / max(r.pivots[0], 1):.1f} us per pivot', flush=True)
x = r.x[0]
frac = np.minimum(x - np.floor(x), np.ceil(x) - x)
cand = np.argsort(-frac, kind='stable')[:4]
deltas = [[(int(j), float(d.l[j]), float(np.floor(x[j])))] for j in cand] + [[(int(j), float(np.ceil(x[j])), float(d.u[j]))] for j in cand]
for slot, lim in ((0, 5), (0, 2147483647), (-1, 5)):
    if slot >= 0:
        lp.simplex_batch(d.l[None], d.u[None], col_status=r.col_status, row_status=r.row_status, max_pivots=0)
    t = time.perf_counter()
    k = lp.simplex_children(d.l, d.u, deltas, col_status=r.col_status[0], row_status=r.row_status[0], parent_slot=slot, max_pivots=lim)
    print(f'8 children slot {slot} limit {lim}: status {k.status.tolist()} pivots {k.pivots.tolist()} kernel {k.stats["step_kernel_ms"]:.1f} ms wall {time.perf_counter() - t:.2f} s', flush=True)
lp.close()
