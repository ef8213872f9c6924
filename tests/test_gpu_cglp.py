"""GPU: the batched cut generating LP (utils/cut_generating_lp.py).

The CGLP of a disjunction is solved for many points at once as LPs that differ only in variable
bounds (its dual form), by the batched dual simplex kernel. Checked against the numpy restatement
of that kernel bit for bit, against HiGHS on the reference's own primal CGLP model for the value,
and against the known answers of the reference's test_cut_generating_lp.py.
"""
import json
import os

import numpy as np
import pytest

from oracle.dual_simplex import dual_simplex
from oracle.highs_lp import HIGHS_INF, HighsLP
from simple_mip_solver_b200 import (BaseNode, BranchAndBound, CutGeneratingLP, CyLPArray,
                                    DisjunctiveCutBoundNode, MILPInstance)

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))


def model_from(rec):
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def partial_tree(rec, node_limit=8):
    bb = BranchAndBound(model_from(rec), BaseNode, node_limit=node_limit, gomory_cuts=False)
    bb.solve()
    return bb


def cold(cglp):
    return (np.full(cglp.lp.nVariables, 3, dtype=np.int32), np.full(cglp.lp.nConstraints, 1, dtype=np.int32))


@pytest.mark.parametrize('name', ['cut1', 'cut2', 'lift_project', 'square', 'small_branch', 'random'])
def test_batched_cglp_bit_exact_and_optimal(blp_lib, name):
    bb = partial_tree(EXAMPLES[name])
    root = bb.root_node
    cglp = CutGeneratingLP(bb, root.idx)
    x = np.asarray(root.solution, dtype=float)
    rng = np.random.default_rng(5)
    points = [CyLPArray(x)] + [CyLPArray(np.maximum(x * rng.uniform(.7, 1.2, len(x)), 0)) for _ in range(15)]
    cuts = cglp.solve_batch(points, starting_bases=[cold(cglp)] * len(points))
    assert cglp.batch_calls == 1 and cglp.points_solved == len(points)
    lo, hi = cglp._point_bounds(np.array(points))
    n = cglp.n
    dM = cglp._dM.toarray()
    fin = lambda v, big: np.where(np.isinf(v), big, v)
    for k, (p, (pi, pi0)) in enumerate(zip(points, cuts)):
        ref = dual_simplex(dM, cglp._dr, cglp._dc, lo[k], hi[k], col_status=np.full(dM.shape[1], 3, np.int8),
                           row_status=np.full(dM.shape[0], 1, np.int8))
        assert ref.status == 0 and pi is not None, (name, k)
        assert np.array_equal(pi, ref.y[n:2 * n] - ref.y[:n]) and pi0 == ref.y[2 * n] - ref.y[2 * n + 1], (name, k)
        c = np.zeros(cglp.lp.nVariables)
        c[:n], c[n] = p, -1.0
        h = HighsLP(cglp._M, c, cglp._r, np.full(cglp._M.shape[0], HIGHS_INF),
                    fin(cglp._lo, -HIGHS_INF), fin(cglp._hi, HIGHS_INF)).solve()
        assert h.status == 0 and abs(float(np.dot(pi, p)) - pi0 - h.objective) <= 1e-8, (name, k)
        for leaf in bb.tree.get_leaves(root.idx):
            if leaf.lp_feasible and leaf.solution is not None:
                assert float(np.dot(pi, np.maximum(leaf.solution, 0))) >= pi0 - 1e-6
    # the last point's basis restarts its solve without a pivot
    again = CutGeneratingLP(bb, root.idx)
    pi2, pi02 = again.solve(x_star=points[-1], starting_basis=cglp.lp.getBasisStatus())
    assert again.lp.iteration == 0 and np.allclose(pi2, cuts[-1][0]) and pi02 == pytest.approx(cuts[-1][1])
    cglp.close()
    again.close()
    bb.model.lp._shared.close()


def test_cglp_known_answers_of_the_reference_on_the_device(blp_lib):
    """test_cut_generating_lp.py:371-415."""
    bb = BranchAndBound(model_from(EXAMPLES['square']), BaseNode, gomory_cuts=False)
    bb.solve()
    pi, pi0 = CutGeneratingLP(bb, bb.root_node.idx).solve()
    np.testing.assert_allclose(pi / pi0, [0, 1] if abs(pi[1]) > abs(pi[0]) else [1, 0], atol=.01)
    assert (pi - .01 < 0).all() and pi0 - .01 < 0
    bb = partial_tree(EXAMPLES['small_branch'], node_limit=10)
    pi, pi0 = CutGeneratingLP(bb, bb.root_node.idx).solve()
    np.testing.assert_allclose(pi / pi0, [0, 0, 1], atol=.01)
    bb = partial_tree(EXAMPLES['square'], node_limit=1)
    pi, pi0 = CutGeneratingLP(bb, bb.root_node.idx).solve(x_star=CyLPArray([1.5, 2]))
    assert pi0 == pytest.approx(-.75, abs=.01)
    np.testing.assert_allclose(pi, [0, -.5], atol=.01)


def test_frontier_prefetch_batches_the_first_disjunctive_cut_on_the_device(blp_lib):
    hits = batched = 0
    for name, rec in list(SCALE1.items())[::9] + [('cut2', EXAMPLES['cut2']), ('random', EXAMPLES['random'])]:
        tree = partial_tree(rec)
        want = rec.get('mip_optimum', rec['reference']['BaseNode']['objective'])
        cglp = CutGeneratingLP(tree, tree.root_node.idx)
        bb = BranchAndBound(model_from(rec), DisjunctiveCutBoundNode, cglp=cglp, gomory_cuts=False, frontier_batch=8)
        bb.solve()
        assert bb.status == 'optimal' and abs(bb.objective_value - want) <= 1e-6 * max(1, abs(want)), name
        hits += cglp.prefetch_hits
        batched += cglp.points_solved - cglp.batch_calls
        cglp.close()
        bb.model.lp._shared.close()
        tree.model.lp._shared.close()
    assert hits > 0 and batched > 0


def test_cglp_through_the_first_order_kernels(blp_lib):
    """A CGLP above blp_simplex_batch_rows() rows is solved by PDHG; forced here on a small one and
    compared with the exact path: same optimum, valid cuts."""
    bb = partial_tree(EXAMPLES['random'])
    root = bb.root_node
    x = np.asarray(root.solution, dtype=float)
    points = [CyLPArray(x), CyLPArray(x * .9), CyLPArray(np.maximum(x - .05, 0))]
    exact = CutGeneratingLP(bb, root.idx)
    want = exact.solve_batch(points)
    first_order = CutGeneratingLP(bb, root.idx)
    first_order.method = 'pdhg'
    got = first_order.solve_batch(points)
    assert first_order.batch_calls == 1
    for p, (pi, pi0), (qi, qi0) in zip(points, want, got):
        assert qi is not None
        assert abs((float(np.dot(qi, p)) - qi0) - (float(np.dot(pi, p)) - pi0)) <= 1e-6
        for leaf in bb.tree.get_leaves(root.idx):
            if leaf.lp_feasible and leaf.solution is not None:
                assert float(np.dot(qi, np.maximum(leaf.solution, 0))) >= qi0 - 1e-5
    exact.close()
    first_order.close()
    bb.model.lp._shared.close()
