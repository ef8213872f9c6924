"""Continuous batching (blp_opts.max_active) vs resident slices on the C5 frontier: node-LPs/s and
iterations for a few (nodes per call, resident slots, evaluation period) settings."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _tuning  # noqa: F401  (tuning build of libblp.so: reads the BLP_* variables below)
import torch
import bench
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import frontier_nodes

wl = sys.argv[1] if len(sys.argv) > 1 else 'c5'
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
d, depth, root = bench.load_instance(wl)
n, m = d.n, d.m
dev = torch.device('cuda', 0)
t = time.time()
lbs, ubs, _ = frontier_nodes(d, root['x'], 0, NB, depth, seed=0)
print('frontier of', NB, 'nodes generated in', round(time.time() - t, 1), 's', flush=True)
lp = engine.BatchLP(d.A, d.b, d.c)
int_idx = torch.arange(n, dtype=torch.int32, device=dev)


def run(first, count, slots, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    ld = engine.leading_dim(count)
    lb = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    ub = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    lb[:, :count] = torch.from_numpy(lbs[first:first + count]).to(dev).T
    ub[:, :count] = torch.from_numpy(ubs[first:first + count]).to(dev).T
    x0 = torch.from_numpy(root['x']).to(dev)[:, None].expand(n, ld).contiguous()
    y0 = torch.from_numpy(root['y']).to(dev)[:, None].expand(m, ld).contiguous()
    o = engine.default_opts(max_active=slots, **kw)
    r = lp.solve_batch_device(lb.contiguous(), ub.contiguous(), x0=x0, y0=y0, int_idx=int_idx, opts=o,
                              want_x=False, want_y=False)
    st = r['status'][:count].cpu().numpy()
    it = r['iters'][:count].cpu().numpy()
    s = r['stats']
    for k in (env or {}):
        os.environ.pop(k, None)
    return dict(count=count, slots=slots, env=env, kw=kw, total_ms=round(s['total_ms']), lps=round(count / (s['total_ms'] * 1e-3), 2),
                batch_iters=s['iterations'], mean_it=int(it.mean()), p90=int(np.percentile(it, 90)), max_it=int(it.max()),
                unsolved=int((st == 3).sum()), refills=s['refills'], compactions=s['compactions'],
                obj_sum=float(r['obj'][:count][torch.from_numpy(st == 0).to(dev)].sum().item()))


cases = [
    (0, 512, 0, None, {}),
    (0, 1024, 512, None, {}),
    (0, 1024, 256, None, {}),
    (0, 1024, 512, {'BLP_ADAPTIVE_EVAL': '0'}, {}),
    (0, 1024, 512, {'BLP_ADAPTIVE_EVAL': '0'}, {'eval_every': 128}),
    (0, 1024, 768, None, {}),
    (0, NB, 512, None, {}),
]
if wl != 'c5':
    cases = [(0, 256, 0, None, {}), (0, NB, 256, None, {}), (0, NB, 128, None, {})]
for c in cases:
    print(json.dumps(run(*c[:4], **c[4])), flush=True)
lp.close()
