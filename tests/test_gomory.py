"""Gomory mixed integer cuts from the active-set basis: the reference's pinned example and
validity on every integer point, at the root and below branching bounds (CPU, exact LP answers)."""
import itertools
import json
import os

import numpy as np
import pytest

from helpers import use_oracle_engine
from simple_mip_solver_b200 import BaseNode, BranchAndBound, CyLPArray, MILPInstance

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))


def model_from(rec):
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def integer_points(rec, lo, hi, cap=12):
    A, b = np.array(rec['A']), np.array(rec['b'])
    ranges = [range(int(np.ceil(a)), int(min(c, a + cap)) + 1) for a, c in zip(lo, hi)]
    for p in itertools.product(*ranges):
        p = np.array(p, dtype=float)
        if (A @ p >= b - 1e-9).all():
            yield p


def test_reference_pins_cut3(monkeypatch):
    use_oracle_engine(monkeypatch)
    m = model_from(EXAMPLES['cut3'])
    node = BaseNode(m.lp, m.integerIndices)
    node._bound_lp()
    assert list(node.basic_variable_indices) == [0, 2, 3]                      # test_base_node.py:681-684
    want = np.array([[1, 2, 0, 0, 1], [0, -2, 1, 0, -3], [0, 0, 0, 1, -5]])    # :669-679
    assert np.max(np.abs(node.tableau - want)) < 1e-4
    cuts = node._find_gomory_cuts()                                             # :654-667
    assert len(cuts) == 1
    assert np.max(np.abs(cuts[0][0] - np.array([-5, -10]))) < 1e-4 and abs(cuts[0][1] + 5) < .01


def test_reference_pins_cut2_rounds(monkeypatch):
    use_oracle_engine(monkeypatch)
    m = model_from(EXAMPLES['cut2'])
    node = BaseNode(m.lp, m.integerIndices, idx=0)
    node._bound_lp()
    assert node.objective_value == pytest.approx(-38.0)                         # test_base_node.py:479-489
    node._cut_generation_iteration()
    assert node.objective_value == pytest.approx(-36.48, abs=0.01)
    m = model_from(EXAMPLES['cut2'])
    node = BaseNode(m.lp, m.integerIndices, idx=0)
    node._base_bound()
    assert node.mip_feasible and node.objective_value == pytest.approx(-36.0, abs=0.01)


@pytest.mark.parametrize('name', ['cut1', 'cut2', 'cut3', 'square', 'small_branch', 'random'])
def test_gmi_cuts_are_valid_below_branching_bounds(monkeypatch, name):
    """Run a few B&B nodes; at every fractional node the GMI cuts must keep every integer point of
    the NODE (its bounds), including when nonbasic variables sit at shifted / upper bounds."""
    use_oracle_engine(monkeypatch)
    rec = EXAMPLES[name]
    bb = BranchAndBound(model_from(rec), BaseNode, node_limit=12, gomory_cuts=False)
    bb.solve()
    checked = 0
    for vert in bb.tree.nodes.values():
        node = vert.attr['node']
        if not node.lp_feasible or node.mip_feasible or node.solution is None:
            continue
        cuts = node._find_gomory_cuts()
        lo = np.asarray(node.lp.variablesLower)
        hi = np.minimum(np.asarray(node.lp.variablesUpper), lo + 12)
        pts = list(integer_points(rec, lo, hi)) if len(rec['c']) <= 3 else []
        for pi, pi0 in cuts.values():
            assert float(np.dot(pi, node.solution)) < pi0 - 1e-9                # separates the LP point
            for p in pts:
                assert float(np.dot(pi, p)) >= pi0 - 1e-7, (name, node.idx, p)
            checked += 1
    assert checked > 0 or name in ('small_branch',)


def test_default_bound_loop_reaches_the_reference_optimum(monkeypatch):
    use_oracle_engine(monkeypatch)
    for name, rec in list(SCALE1.items()) + [(k, EXAMPLES[k]) for k in ('cut1', 'cut2', 'cut3', 'square', 'random')]:
        bb = BranchAndBound(model_from(rec), BaseNode)
        bb.solve()
        gold = rec['reference']['BaseNode_gomory']
        assert bb.status == 'optimal' and bb.objective_value == pytest.approx(gold['objective'], abs=1e-6), name
