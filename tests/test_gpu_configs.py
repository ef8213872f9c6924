"""GPU: the BASELINE.json configurations 3 and 4 at reduced or full size, through the Node API.

config 3: synthetic random MILP 500 vars x 300 rows, 10 % density; strong branching on up to 64
          candidates x 2 children handed to the kernels as ONE batch by PseudoCostBranchNode.
config 4: sparse MILP with a cutting-plane bound: cut rows appended per node batch (shared pool,
          per-node row masks), three rounds; here 2000 x 1000 so that the CPU oracle stays fast.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.highs_lp import HIGHS_INF, HighsLP
from simple_mip_solver_b200 import CyLPArray, MILPInstance, PseudoCostBranchNode
from simple_mip_solver_b200.instances import frontier_nodes, grumpy_random_mip, numpy_random_mip

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def test_config3_strong_branching_batch(blp_lib):
    d = grumpy_random_mip(500, 300, density=0.1, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=2)
    model = MILPInstance(A=d.A.toarray(), b=CyLPArray(d.b), c=CyLPArray(d.c), l=CyLPArray(d.l), u=CyLPArray(d.u),
                         sense=['Min', '>='], integerIndices=d.integer_indices, numVars=d.n)
    node = PseudoCostBranchNode(model.lp, model.integerIndices, idx=0)
    # 100 "pivots" = 51 200 PDHG iterations: enough for every child to reach the 1e-8 tolerance
    rtn = node.bound(pseudo_costs={}, strong_branch_iters=100, gomory_cuts=False)
    ref_root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    assert rel(node.objective_value, ref_root.objective) <= 1e-6
    pcs = rtn['pseudo_costs']
    sh = model.lp._shared
    assert sh.solve_calls == 2 and sh.lps_solved == 1 + 2 * len(pcs)      # root, then ALL children at once
    assert len(pcs) >= 32
    x = node.solution
    checked = 0
    for j, entry in list(pcs.items())[:24]:
        for direction in ('left', 'right'):
            l, u = d.l.copy(), d.u.copy()
            if direction == 'left':
                u[j] = np.floor(x[j])
                change = x[j] - u[j]
            else:
                l[j] = np.ceil(x[j])
                change = l[j] - x[j]
            ref = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), l, u).solve()
            want = max(ref.objective - node.objective_value, 0) / change if ref.status == 0 else 0.0
            assert entry[direction]['times'] == 1
            # a pseudo cost is a difference quotient of two objectives that are each right to 1e-6
            # relative: the admissible error is 2e-6 |obj| / (change of the variable)
            assert entry[direction]['cost'] == pytest.approx(want, rel=1e-3, abs=2e-6 * abs(ref_root.objective) / change)
            checked += 1
    assert checked == 48
    # the branching decision the reference would take from these costs (pseudo_cost.py:118-133)
    assert node._best_pseudo_costs_index(pcs) in pcs


def test_config4_cut_rounds_with_row_masks(blp_lib):
    from simple_mip_solver_b200 import engine
    d = numpy_random_mip(2000, 1000, density=0.01, seed=2)
    root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    B = 32
    lbs, ubs, _ = frontier_nodes(d, root.x, 0, B, 8, seed=3)
    lp = engine.BatchLP(d.A, d.b, d.c)
    rng = np.random.default_rng(5)
    cuts, rhs = [], []
    masks = np.zeros((B, 0), dtype=np.uint8)
    res = lp.solve_batch(lbs, ubs, x0=np.tile(root.x, (B, 1)), y0=np.tile(np.maximum(root.row_dual, 0), (B, 1)))
    prev = res.objective.copy()
    for rnd in range(3):
        # each round appends 8 dense knapsack-cover style rows  -sum_{j in S} x_j >= -floor(sum x*_S)
        # built from one node's current solution; a row is switched on for a random half of the nodes
        new_rows, new_rhs = [], []
        for k in range(8):
            S = rng.choice(d.n, size=200, replace=False)
            xs = res.x[k % B]
            row = np.zeros(d.n)
            row[S] = -1.0
            new_rows.append(row)
            new_rhs.append(-np.floor(xs[S].sum()))
        first = lp.append_rows(np.array(new_rows), np.array(new_rhs))
        assert first == d.m + 8 * rnd
        cuts += new_rows
        rhs += new_rhs
        masks = np.hstack([masks, (rng.random((B, 8)) < 0.5).astype(np.uint8)])
        res = lp.solve_batch(lbs, ubs, row_mask=masks, x0=res.x,
                             y0=np.hstack([res.y, np.zeros((B, 8))]))
        for k in range(0, B, 5):
            h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), lbs[k], ubs[k])
            for t in np.flatnonzero(masks[k]):
                h.add_row(cuts[t], rhs[t])
            ref = h.solve()
            assert res.status[k] == ref.status, (rnd, k)
            if ref.status == 0:
                assert rel(res.objective[k], ref.objective) <= 1e-6, (rnd, k, res.objective[k], ref.objective)
                assert res.objective[k] >= prev[k] - 1e-6 * abs(prev[k])      # cuts only tighten
        ok = res.status == 0
        assert (res.y[ok][:, d.m:][masks[ok] == 0] == 0).all()                 # masked rows carry no dual
        prev = np.where(ok, res.objective, prev)
    lp.truncate_rows(d.m)
    assert lp.m == d.m
    lp.close()
