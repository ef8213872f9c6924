"""C5 stress variant of SURVEY 8d (50 000 x 20 000, density 1e-3, ~1.0 M nonzeros; reported separately).

No simplex root fixture exists for it: HiGHS dual simplex did not finish its root LP in 50 minutes on
this container's cores (tests/tools/make_bench_fixture.py c5s ...), so the frontier is branched from the
root vertex the device itself computes (cold PDHG solve), and what is reported is
  * the root solve (iterations, seconds),
  * a frontier of 512 nodes solved from that root (node-LPs/s, iterations),
  * the step kernels' per-launch time and algorithmic GB/s at that width (blp_opts.profile),
    next to the L2->SM gather traffic 16 nnz B that makes this variant gather-bound.
Writes gpurun_out/stress_variant.json.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes, numpy_random_mip

dens = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n, m = 50000, 20000
d = numpy_random_mip(n, m, density=dens, seed=2)
lp = engine.BatchLP(d.A, d.b, d.c)
t = time.perf_counter()
root = lp.solve_batch(d.l[None], d.u[None], want_y=True)
t_root = time.perf_counter() - t
out = dict(n=n, m=m, density=dens, nnz=int(d.A.nnz), root=dict(status=int(root.status[0]), objective=float(root.objective[0]),
           iterations=int(root.iterations[0]), seconds=t_root))
print(out, flush=True)
_, _, deltas = frontier_nodes(d, root.x[0], 0, B, 32, seed=0, dense=False)
t = time.perf_counter()
r = lp.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0], want_x=False, want_y=False)
dt = time.perf_counter() - t
out['frontier'] = dict(nodes=B, seconds=dt, node_lps_per_s=B / dt, status_counts=np.bincount(r.status, minlength=4).tolist(),
                       mean_iterations=float(r.iterations.mean()), max_iterations=int(r.iterations.max()),
                       us_per_node_iteration=1e3 * r.stats['step_kernel_ms'] / r.stats['node_iterations'])
print(out['frontier'], flush=True)
p = lp.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0], want_x=False, want_y=False,
                      opts=engine.default_opts(max_iters=512, profile=1)).stats
it = p['iterations']
bytes_A = 12 * d.A.nnz + 4 * (m + 1)
bytes_AT = 12 * d.A.nnz + 4 * (n + 1)
primal_ms, dual_ms = p['primal_kernel_ms'] / it, p['dual_kernel_ms'] / it
width = p['node_iterations'] / it
# active (coordinate, node) pairs per iteration: frozen coordinates are neither updated (20 bytes) nor gathered (8)
act_c = (p['node_iterations'] * n - p['skipped_col_updates']) / it
act_r = (p['node_iterations'] * m - p['skipped_row_updates']) / it
pbytes, dbytes = 20 * act_c + 8 * act_r + bytes_AT, 20 * act_r + 8 * act_c + bytes_A
peak = 6547.5
out['kernels'] = dict(width=width, k_primal_ms=primal_ms, k_dual_ms=dual_ms,
                      frozen_col_share=1 - act_c / (width * n), frozen_row_share=1 - act_r / (width * m),
                      k_primal_gbs=pbytes / primal_ms / 1e6, k_dual_gbs=dbytes / dual_ms / 1e6,
                      k_primal_frac=pbytes / primal_ms / 1e6 / peak, k_dual_frac=dbytes / dual_ms / 1e6 / peak,
                      hbm_bytes_per_iteration=pbytes + dbytes,
                      hbm_bytes_per_iteration_without_freezing=28 * (n + m) * width + bytes_A + bytes_AT, peak_gbs=peak)
print(out['kernels'], flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/stress_variant.json', 'w'), indent=1)
lp.close()
