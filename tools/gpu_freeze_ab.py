"""A/B of frozen-coordinate skipping (blp_opts.freeze) on one frontier slice: iterations, device time, skipped share.

    python tools/gpu_freeze_ab.py c5 512 0          # 512 resident nodes, no refill
    python tools/gpu_freeze_ab.py c5 1024 512       # 1024 nodes through 512 slots (the bench's step)
    python tools/gpu_freeze_ab.py c4 512 0 0 0.05:0 0.05:0.95   # variants: margin[:step_safety]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from simple_mip_solver_b200 import engine                       # noqa: E402
from simple_mip_solver_b200.instances import frontier_nodes     # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'c5'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
slots = int(sys.argv[3]) if len(sys.argv) > 3 else 0
specs = sys.argv[4:] or ['0', '0.05:0', '0.05:0.95']      # margin[:step_safety]; margin 0 = freezing off
d, depth, root = bench.load_instance(wl)
lp = engine.BatchLP(d.A, d.b, d.c)
n, m = d.A.shape[1], d.A.shape[0]
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, B, depth, seed=0)
x0 = np.tile(root['x'], (B, 1))
y0 = np.tile(root['y'], (B, 1))
ref = None
for spec in specs:
    mg, safety = (float(a) for a in (spec.split(':') + ['0.95'])[:2])
    o = engine.default_opts(max_active=slots, freeze=int(mg > 0), freeze_margin=mg if mg > 0 else 0.05, step_safety=safety)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, want_x=False, want_y=False, opts=o)
    s, it = r.stats, r.iterations
    if ref is None:
        ref = r
    print(json.dumps(dict(workload=wl, nodes=B, slots=slots, margin=mg, step_safety=safety, mean_it=int(it.mean()), p90_it=int(np.percentile(it, 90)), max_it=int(it.max()),
                          slowest=[(int(k), int(it[k])) for k in np.argsort(-it)[:4]],
                          total_ms=round(s['total_ms']), step_ms=round(s['step_kernel_ms']),
                          us_per_node_iter=round(1e3 * s['step_kernel_ms'] / s['node_iterations'], 4),
                          skipped_cols=round(s['skipped_col_updates'] / (s['node_iterations'] * n), 4),
                          skipped_rows=round(s['skipped_row_updates'] / (s['node_iterations'] * m), 4),
                          launches=s['kernel_launches'], step_resets=s['step_resets'], unsolved=int((r.status != 0).sum()),
                          status_equal=bool((r.status == ref.status).all()),
                          max_rel_obj_diff=float(np.max(np.abs(r.objective - ref.objective) / (1 + np.abs(ref.objective)))))),
          flush=True)
lp.close()
