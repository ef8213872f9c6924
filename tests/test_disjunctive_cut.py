"""DisjunctiveCutBoundNode / CutGeneratingLP: host logic on exact LP answers (CPU), plus validity
properties of the generated cuts (the reference's tests use the same properties:
test_cut_generating_lp.py, test_disjunctive_cut.py)."""
import itertools
import json
import os

import numpy as np
import pytest

from helpers import use_oracle_engine
from simple_mip_solver_b200 import (BaseNode, BranchAndBound, CutGeneratingLP, CyLPArray,
                                    DisjunctiveCutBoundNode, DisjunctiveCutBoundPseudoCostBranchNode,
                                    MILPInstance)

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))


def model_from(rec):
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def integer_points(rec, cap=6):
    """All integer points of the (small) model that satisfy A x >= b within the bounds."""
    A, b = np.array(rec['A']), np.array(rec['b'])
    ranges = [range(int(lo), int(min(hi, lo + cap)) + 1) for lo, hi in zip(rec['l'], rec['u'])]
    pts = [np.array(p, dtype=float) for p in itertools.product(*ranges)]
    return [p for p in pts if (A @ p >= b - 1e-9).all()]


def partial_tree(rec, node_limit=8):
    bb = BranchAndBound(model_from(rec), BaseNode, node_limit=node_limit, gomory_cuts=False)
    bb.solve()
    return bb


@pytest.mark.parametrize('name', ['cut1', 'cut2', 'lift_project', 'square', 'small_branch'])
def test_cglp_cut_is_valid_and_separates(monkeypatch, name):
    use_oracle_engine(monkeypatch)
    rec = EXAMPLES[name]
    bb = partial_tree(rec)
    root = bb.root_node
    cglp = CutGeneratingLP(bb, root.idx)
    assert cglp.lp.nVariables > len(rec['c']) + 1 and cglp.lp.nConstraints >= 2
    pi, pi0 = cglp.solve()
    assert pi is not None and not cglp.cylp_failure
    if not root.mip_feasible and np.linalg.norm(pi) > 1e-9:
        assert float(np.dot(pi, root.solution)) < pi0 - 1e-9          # cuts off the root LP vertex
    for p in integer_points(rec):
        assert float(np.dot(pi, p)) >= pi0 - 1e-7, (name, p)             # valid for every MIP point
    pi2, pi02 = cglp.solve(x_star=CyLPArray(np.asarray(root.solution) * 0.9))
    for p in integer_points(rec):
        assert float(np.dot(pi2, p)) >= pi02 - 1e-7


def test_cglp_argument_checks(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = partial_tree(EXAMPLES['cut1'])
    with pytest.raises(AssertionError, match='bb must be a BranchAndBound instance'):
        CutGeneratingLP('bb', 0)
    with pytest.raises(AssertionError, match='root node of the disjunction must be present'):
        CutGeneratingLP(bb, 999)
    with pytest.raises(AssertionError, match='A and b must both have values'):
        CutGeneratingLP(bb, 0, A=np.matrix(np.eye(2)))
    cglp = CutGeneratingLP(bb, 0)
    with pytest.raises(AssertionError, match='x_star must be a CyLPArray'):
        cglp.solve(x_star=[1, 2])
    with pytest.raises(AssertionError, match='starting basis must be an iterable with two elements'):
        cglp.solve(starting_basis=(np.zeros(1),))
    m = model_from(EXAMPLES['cut1'])
    with pytest.raises(AssertionError, match='cglp must be CutGeneratingLP instance'):
        DisjunctiveCutBoundNode(cglp='x', lp=m.lp, integer_indices=m.integerIndices)
    with pytest.raises(AssertionError, match='cannot force'):
        DisjunctiveCutBoundNode(force_create_cglp=True, lp=m.lp, integer_indices=m.integerIndices)
    with pytest.raises(AssertionError, match='is bool'):
        DisjunctiveCutBoundNode(force_create_cglp=1, lp=m.lp, integer_indices=m.integerIndices)


KW = [dict(), dict(gomory_cuts=False), dict(max_cglp_calls=1, gomory_cuts=False),
      dict(cglp_cumulative_constraints=True, cglp_cumulative_bounds=True, gomory_cuts=False),
      dict(warm_start_cglp=False, cglp_cumulative_bounds=True, gomory_cuts=False)]


@pytest.mark.parametrize('kw', KW)
def test_branch_and_bound_with_disjunctive_cuts(monkeypatch, kw):
    """helpers.py:75-126 of the reference: first a short B&B gives the disjunction, then the
    model is solved with disjunctive cut nodes; the optimum must be the MIP optimum."""
    use_oracle_engine(monkeypatch)
    recs = list(SCALE1.items())[::7] + [(k, EXAMPLES[k]) for k in ('cut1', 'cut2', 'lift_project', 'small_branch')]
    made = added = 0
    for name, rec in recs:
        tree = partial_tree(rec)
        cglp = CutGeneratingLP(tree, tree.root_node.idx)
        for Node, extra in ((DisjunctiveCutBoundNode, {}), (DisjunctiveCutBoundPseudoCostBranchNode, dict(pseudo_costs={}))):
            bb = BranchAndBound(model_from(rec), Node, cglp=cglp, **dict(kw), **extra)
            bb.solve()
            want = rec.get('mip_optimum', rec['reference']['BaseNode']['objective'])
            assert bb.status == 'optimal', name
            assert bb.objective_value == pytest.approx(want, abs=1e-6), (name, kw)
            made += bb._kwargs['total_number_cglp_created']
            added += bb._kwargs['total_number_cglp_added']
            assert bb._kwargs['total_number_cglp_removed'] >= 0
    assert made > 0 and added > 0


def test_children_inherit_cglp_only_after_a_useful_cut(monkeypatch):
    use_oracle_engine(monkeypatch)
    rec = EXAMPLES['cut1']
    tree = partial_tree(rec)
    cglp = CutGeneratingLP(tree, tree.root_node.idx)
    m = model_from(rec)
    node = DisjunctiveCutBoundNode(lp=m.lp, integer_indices=m.integerIndices, idx=0, cglp=cglp)
    assert node.previous_cglp_added and not node.current_node_added_cglp and node.prev_cglp_basis is None
    rtn = node.bound(gomory_cuts=False, max_cut_generation_iterations=3)
    assert rtn['total_number_cglp_created'] >= 1
    if not node.mip_feasible and node.lp_feasible:
        kids = node.branch(next_node_idx=1)
        for d in ('left', 'right'):
            assert isinstance(kids[d], DisjunctiveCutBoundNode)
            assert (kids[d].cglp is cglp) == node.current_node_added_cglp
