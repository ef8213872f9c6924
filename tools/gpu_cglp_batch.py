"""GPU: what batching the cut generating LP buys (utils/cut_generating_lp.py).

One disjunction (the open leaves of a short branch and bound on a GrUMPy-style random MILP), K points to
cut off — the LP solutions a frontier of K nodes would present. Times K single `solve` calls against one
`solve_batch` call, both cold-started, and checks that they return the same cuts.

    python tools/gpu_cglp_batch.py [n_vars n_rows node_limit]
"""
import json
import sys
import time

import numpy as np

from simple_mip_solver_b200 import BaseNode, BranchAndBound, CutGeneratingLP, CyLPArray, MILPInstance
from simple_mip_solver_b200.instances import grumpy_random_mip


def main():
    nv, nr, limit = (int(a) for a in (sys.argv[1:4] + ['20', '10', '8'][len(sys.argv) - 1:]))
    d = grumpy_random_mip(numVars=nv, numCons=nr, density=0.4, rand_seed=2)
    model = MILPInstance(A=d.A.toarray(), b=CyLPArray(d.b), c=CyLPArray(d.c), l=CyLPArray(d.l), u=CyLPArray(d.u),
                         sense=['Min', '>='], integerIndices=list(d.integer_indices), numVars=nv)
    bb = BranchAndBound(model, BaseNode, node_limit=limit, gomory_cuts=False)
    bb.solve()
    cglp = CutGeneratingLP(bb, bb.root_node.idx)
    x = np.asarray(bb.root_node.solution, dtype=float)
    rng = np.random.default_rng(7)
    out = dict(n=nv, rows=nr, terms=len(cglp.term_ids), device_lp=list(cglp._dM.shape), runs=[])
    cold = (np.full(cglp.lp.nVariables, 3, dtype=np.int32), np.full(cglp.lp.nConstraints, 1, dtype=np.int32))
    cglp.solve(starting_basis=cold)                      # creates the device copy
    for K in (1, 16, 64, 256):
        pts = [CyLPArray(np.maximum(x * rng.uniform(.7, 1.2, nv), 0)) for _ in range(K)]
        t0 = time.perf_counter()
        single = [cglp.solve(x_star=p, starting_basis=cold) for p in pts]
        t1 = time.perf_counter()
        batch = cglp.solve_batch(pts, starting_bases=[cold] * K)
        t2 = time.perf_counter()
        same = all(np.array_equal(a[0], b_[0]) and a[1] == b_[1] for a, b_ in zip(single, batch))
        out['runs'].append(dict(points=K, single_ms=round(1e3 * (t1 - t0), 2), batch_ms=round(1e3 * (t2 - t1), 2),
                                identical=bool(same), last_pivots=cglp.lp.iteration))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
