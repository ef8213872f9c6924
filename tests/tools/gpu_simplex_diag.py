"""GPU diagnostic (not a test): dual simplex kernel vs the numpy restatement, printing every
difference instead of stopping at the first one. Run on the GPU box:
    python tests/tools/gpu_simplex_diag.py [small|c3]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.dual_simplex import dual_simplex  # noqa: E402
from simple_mip_solver_b200 import engine  # noqa: E402
from simple_mip_solver_b200.instances import grumpy_random_mip  # noqa: E402


def compare(tag, res, k, ref):
    bad = []
    if res.status[k] != ref.status:
        bad.append(f'status {res.status[k]} vs {ref.status}')
    if res.pivots[k] != ref.pivots:
        bad.append(f'pivots {res.pivots[k]} vs {ref.pivots}')
    if not np.array_equal(res.col_status[k], ref.col_status) or not np.array_equal(res.row_status[k], ref.row_status):
        bad.append('basis differs')
    if ref.status != 1:
        dx = float(np.max(np.abs(res.x[k] - ref.x)))
        if dx != 0.0:
            bad.append(f'max |dx| {dx:.3e}')
        if res.objective[k] != ref.objective:
            bad.append(f'obj {res.objective[k]!r} vs {ref.objective!r}')
        dy = float(np.max(np.abs(res.y[k] - ref.y))) if len(ref.y) else 0.0
        if dy != 0.0:
            bad.append(f'max |dy| {dy:.3e}')
    print(f'{tag}: ' + ('identical' if not bad else '; '.join(bad)), flush=True)
    return not bad


def main(which):
    gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'example_models.json')))
    ok = True
    if which in ('small', 'all'):
        for name in ('small_branch', 'no_branch', 'infeasible', 'unbounded', 'cut2', 'cut3', 'random'):
            rec = gold[name]
            A, b, c, l, u = (np.array(rec[k], float) for k in ('A', 'b', 'c', 'l', 'u'))
            lp = engine.BatchLP(A, b, c)
            res = lp.simplex_batch(l[None], u[None])
            ok &= compare(name, res, 0, dual_simplex(A, b, c, l, u))
            lp.close()
        d = grumpy_random_mip(40, 20, density=0.2, rand_seed=2)
        lp = engine.BatchLP(d.A, d.b, d.c)
        res = lp.simplex_batch(d.l[None], d.u[None])
        ref = dual_simplex(d.A.toarray(), d.b, d.c, d.l, d.u)
        ok &= compare('40x20 root', res, 0, ref)
        print('   stats', res.stats)
        lp.close()
    if which in ('c3', 'all'):
        d = grumpy_random_mip(500, 300, density=0.1, rand_seed=2)
        A = d.A.toarray()
        lp = engine.BatchLP(d.A, d.b, d.c)
        t = time.time()
        res = lp.simplex_batch(d.l[None], d.u[None])
        print(f'c3 root: {time.time() - t:.4f} s wall, stats {res.stats}')
        ref = dual_simplex(A, d.b, d.c, d.l, d.u)
        ok &= compare('c3 root', res, 0, ref)
        x = ref.x
        ints = np.asarray(d.integer_indices)
        frac = np.minimum(x[ints] - np.floor(x[ints]), np.ceil(x[ints]) - x[ints])
        cand = ints[np.argsort(-frac, kind='stable')][:64]
        deltas = []
        for j in cand:
            deltas.append([(int(j), float(d.l[j]), float(np.floor(x[j])))])
            deltas.append([(int(j), float(np.ceil(x[j])), float(d.u[j]))])
        for cached, limit in ((False, 2147483647), (True, 2147483647), (False, 5), (True, 5)):
            if cached:
                lp.simplex_batch(d.l[None], d.u[None])
            t = time.time()
            kids = lp.simplex_children(d.l, d.u, deltas, col_status=res.col_status[0], row_status=res.row_status[0],
                                       parent_slot=0 if cached else -1, max_pivots=limit)
            print(f'c3 128 children cached={cached} limit={limit}: {time.time() - t:.4f} s wall, stats {kids.stats}, '
                  f'status counts {np.bincount(kids.status, minlength=4).tolist()}', flush=True)
            for k in (0, 1, 50, 127):
                l, u = d.l.copy(), d.u.copy()
                for j, lo, hi in deltas[k]:
                    l[j], u[j] = lo, hi
                r = dual_simplex(A, d.b, d.c, l, u, col_status=ref.col_status, row_status=ref.row_status,
                                 max_pivots=limit, start=ref if cached else None)
                ok &= compare(f'   child {k}', kids, k, r)
        lp.close()
    print('ALL IDENTICAL' if ok else 'DIFFERENCES FOUND')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'all')
