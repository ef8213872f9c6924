"""Exploratory GPU timing of the step kernels at the C5 shape (writes gpurun_out/explore.json)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _tuning  # noqa: F401  (tuning build of libblp.so: reads the BLP_* variables below)
import torch
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import numpy_random_mip

out = {}
n, m, dens, B = [int(a) if i != 2 else float(a) for i, a in enumerate((sys.argv[1:5] + ['50000', '20000', '2e-4', '4096'])[:4])]
d = numpy_random_mip(n, m, density=dens, seed=2)
t = time.time(); lp = engine.BatchLP(d.A, d.b, d.c); out['create_s'] = time.time() - t
ld = engine.leading_dim(B)
dev = torch.device('cuda', 0)
lb = torch.zeros((n, ld), dtype=torch.float64, device=dev)
ub = torch.full((n, ld), 10.0, dtype=torch.float64, device=dev)
# a random dive: each node fixes a few random variables' upper bound
g = torch.Generator(device=dev); g.manual_seed(0)
idx = torch.randint(0, n, (16, ld), device=dev, generator=g)
ub.scatter_(0, idx, torch.randint(0, 3, (16, ld), device=dev, generator=g).double())
nnz = d.A.nnz
bytes_iter = 12 * nnz * 2 + 4 * (m + 1) + 4 * (n + 1) + 8 * B * (6 * n + 3 * m)
for rpw in (1, 2, 4, 8, 16):
    os.environ['BLP_ROWS_PER_WARP'] = str(rpw)
    for graph in (1,):
        o = engine.default_opts(max_iters=256, eval_every=64, use_graph=graph)
        r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
        r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
        s = r['stats']
        gbs = bytes_iter * s['iterations'] / (s['step_kernel_ms'] * 1e-3) / 1e9
        out[f'rpw{rpw}_graph{graph}'] = dict(stats=s, algo_GBs=gbs)
        print(rpw, graph, s, 'algo GB/s', gbs, flush=True)
os.environ['BLP_ROWS_PER_WARP'] = '4'
o = engine.default_opts(max_iters=128, eval_every=64, profile=1)
r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False)
print('profile', r['stats'])
out['profile'] = r['stats']
# how long does a full solve take?
o = engine.default_opts(max_iters=int(os.environ.get('FULL_ITERS', '20000')), verbose=1)
t = time.time(); r = lp.solve_batch_device(lb, ub, opts=o, want_x=False, want_y=False); dt = time.time() - t
st = r['status'][:B].cpu().numpy(); it = r['iters'][:B].cpu().numpy()
out['full'] = dict(stats=r['stats'], wall_s=dt, status_counts={int(k): int((st == k).sum()) for k in np.unique(st)},
                   iters_mean=float(it.mean()), iters_max=int(it.max()))
print(out['full'])
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/explore_%d_%d.json' % (n, B), 'w'), indent=1)
