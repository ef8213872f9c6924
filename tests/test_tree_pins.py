"""CPU: the reference's known answers about the branch-and-bound TREE, on
exact LP answers (tests/helpers.OracleBatchLP): leaf queries, subtree dual bounds, the disjunction of
a subtree and the prune rule. Sources: test_simple_mip_solver/test_algorithms/test_branch_and_bound.py
:48-134 (get_leaves), :136-150 (get_disjunction), :168-188 (subtree_dual_bound), :420-436 (prune).
"""
import json
import os
from unittest.mock import patch

import numpy as np
import pytest

from helpers import use_oracle_engine
from simple_mip_solver_b200 import BaseNode, BranchAndBound, CyLPArray, MILPInstance

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))


def model(name):
    rec = EXAMPLES[name]
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def test_get_leaves_small_branch(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model('small_branch'), gomory_cuts=False, node_limit=1)
    bb.solve()
    assert len(bb.tree.get_leaves(0, keep='not infeasible')) == 2         # children not bounded yet
    assert not bb.tree.get_leaves(0, keep='feasible')
    bb.node_limit = float('inf')
    bb.solve()
    ids = lambda nodes: {n.idx for n in nodes}
    leaves = ids(bb.tree.get_leaves(0))
    for node_id in bb.tree.nodes:
        assert len(bb.tree.get_children(node_id)) == (0 if node_id in leaves else 2)
    feasible = ids(bb.tree.get_leaves(0, keep='feasible'))
    for node_id, v in bb.tree.nodes.items():
        if node_id in feasible:
            assert not bb.tree.get_children(node_id) and v.attr['node'].lp_feasible
        else:
            assert len(bb.tree.get_children(node_id)) == 2 or not v.attr['node'].lp_feasible
    assert ids(bb.tree.get_leaves(2, depth=0)) == {2}
    assert not bb.tree.get_leaves(2, depth=0, keep='feasible')
    assert ids(bb.tree.get_leaves(0, depth=1)) == {1, 2}
    assert ids(bb.tree.get_leaves(0, depth=1, keep='feasible')) == {1}
    d2 = bb.tree.get_leaves(1, depth=2)
    assert ids(d2) == {5, 6, 7, 8}
    assert all(bb.tree.get_parent(bb.tree.get_parent(n.idx)) == 1 for n in d2)
    assert ids(bb.tree.get_leaves(1, depth=2, keep='feasible')) == {5, 7}
    d3 = bb.tree.get_leaves(1, depth=3)
    assert ids(d3) == {5, 6, 8, 9, 10}
    for n in d3:
        up = bb.tree.get_parent(bb.tree.get_parent(n.idx))
        assert (up if n.idx <= 8 else bb.tree.get_parent(up)) == 1
    assert ids(bb.tree.get_leaves(1, depth=3, keep='feasible')) == {5, 9}
    with pytest.raises(AssertionError, match='subtree_root_id must belong to the tree'):
        bb.tree.get_leaves(20)
    with pytest.raises(AssertionError, match='depth is a nonnegative integer'):
        bb.tree.get_leaves(subtree_root_id=0, depth=1.5)
    with pytest.raises(AssertionError, match="keep is one of 'all', 'feasible', or 'not infeasible'"):
        bb.tree.get_leaves(subtree_root_id=0, keep=False)


def test_get_disjunction_and_node_instances(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model('small_branch'), gomory_cuts=False)
    bb.solve()
    dis = bb.tree.get_disjunction(0)
    assert list(dis[5][0]) == [0, 0, 0] and list(dis[5][1]) == [0, 1, 1]
    assert list(dis[11][0]) == [1, 0, 0] and list(dis[11][1]) == [1, 1, 0]
    with pytest.raises(AssertionError, match='subtree_root_id must belong to the tree'):
        bb.tree.get_disjunction(20)
    n1, n2 = bb.tree.get_node_instances([1, 2])
    assert (n1.idx, n2.idx) == (1, 2) and isinstance(n1, BaseNode) and isinstance(n2, BaseNode)
    assert bb.tree.get_node_instances(1).idx == 1
    with pytest.raises(AssertionError, match='must be an integer or iterable'):
        bb.tree.get_node_instances('1')
    with pytest.raises(AssertionError, match='node_ids are not in the tree'):
        bb.tree.get_node_instances([20])
    del bb.tree.nodes[0].attr['node']
    with pytest.raises(AssertionError, match='must have an attribute for a node instance'):
        bb.tree.get_node_instances([0])


def test_subtree_dual_bound(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model('small_branch'), gomory_cuts=False, node_limit=1)
    with pytest.raises(AssertionError, match='subtree_root_id must belong to the tree'):
        bb.tree.subtree_dual_bound(subtree_root_id=1)
    assert bb.tree.subtree_dual_bound(0) == -float('inf')          # nothing bounded yet
    bb.solve()
    assert bb.tree.subtree_dual_bound(0) == -2.75                  # the root LP
    bb.node_limit = 2
    bb.solve()
    assert bb.tree.subtree_dual_bound(0) == -2.75
    bb.node_limit = float('inf')
    bb.solve()
    assert bb.tree.subtree_dual_bound(0) == -2
    assert bb.tree.subtree_dual_bound(2) == float('inf')           # the infeasible right child
    assert bb.tree.subtree_dual_bound(0, depth=1) == -2.75
    assert bb.dual_bound == -2


def test_evaluate_node_prunes_by_dual_bound(monkeypatch):
    """A node whose dual bound cannot beat the incumbent is neither bounded nor counted."""
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model('no_branch'), initial_primal_bound=-2)
    called = BaseNode(bb.model.lp, bb.model.integerIndices, dual_bound=-4)
    pruned = BaseNode(bb.model.lp, bb.model.integerIndices, dual_bound=0)
    with patch.object(called, 'bound') as cnb, patch.object(pruned, 'bound') as pnb:
        cnb.return_value = {}
        pnb.return_value = {}
        bb._node_queue.put(called)
        bb._node_queue.put(pruned)
        bb._evaluate_node(bb._node_queue.get())
        bb._evaluate_node(bb._node_queue.get())
        assert cnb.call_count == 1 and pnb.call_count == 0
        assert bb._node_queue.empty() and bb.evaluated_nodes == 1


def test_unbounded_root_stops_the_search(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model('unbounded'))
    bb._evaluate_node(bb.root_node)
    assert bb._unbounded and bb.evaluated_nodes == 1
    bb2 = BranchAndBound(model('unbounded'))
    bb2.solve()
    assert bb2.status == 'unbounded'
