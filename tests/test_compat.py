"""The CyLP / cuppy / gimpy look-alikes: the calls the reference makes on them behave the same."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import use_oracle_engine
from simple_mip_solver_b200.compat import (COIN_INFINITY, BinaryTree, CyClpSimplex, CyLPArray, MILPInstance,
                                           SharedLP, read_mps, solve_lps)


def _lp():
    A = sp.csr_matrix(np.array([[-1., 0, -1], [0, -1, 0]]))
    return CyClpSimplex(SharedLP(A, [-1.5, -1.25], [-1, -1, -1]), [0, 0, 0], [10, 10, 10])


def test_protocol_attributes():
    lp = _lp()
    assert lp.nVariables == lp.nCols == 3 and lp.nConstraints == 2
    assert len(lp.variables) == 1 and lp.variables[0].name == 'x' and lp.getVarByName('x').dim == 3
    assert lp.getCoinInfinity() >= 1e300
    assert np.array_equal(lp.constraintsLower, [-1.5, -1.25]) and (lp.constraintsUpper >= 1e300).all()
    assert lp.coefMatrix.shape == (2, 3) and sp.issparse(lp.coefMatrix)
    assert np.array_equal(lp.objective, [-1, -1, -1])
    c0 = lp.constraints[0]
    assert c0.nRows == 2 and c0.varCoefs[lp.getVarByName('x')].shape == (2, 3)
    lp.variablesUpper = CyLPArray([5, 5, 5])            # test_base_node.py:126-128 sets bounds
    assert list(lp.variablesUpper) == [5, 5, 5]
    cols, rows = lp.getBasisStatus()
    assert cols.dtype == np.int32 and len(cols) == 3 and len(rows) == 2


def test_modelling_expressions_and_cut_rows(monkeypatch):
    use_oracle_engine(monkeypatch)
    lp = _lp()
    x = lp.getVarByName('x')
    lp.addConstraint(CyLPArray([-1, -1, -1]) * x >= -2.5, 'cut_a')        # base_node.py:459-460
    assert lp.nConstraints == 3 and lp.coefMatrix.shape == (3, 3)
    assert [c.name for c in lp.constraints] == ['R_base', 'cut_a']
    lp.dual()
    assert lp.getStatusCode() == 0 and lp.objectiveValue == pytest.approx(-2.5)
    duals = lp.dualConstraintSolution
    assert set(duals) == {'R_base', 'cut_a'} and duals['cut_a'][0] > 0
    assert lp.primalVariableSolution['x'].shape == (3,)
    lp.removeConstraint('cut_a')                                          # base_node.py:337-338
    lp.dual()
    assert lp.objectiveValue == pytest.approx(-2.75)
    with pytest.raises(KeyError):
        lp.removeConstraint('cut_a')
    # bounds statement and building a model from scratch (base_node.py:592-608 style)
    new = CyClpSimplex()
    y = new.addVariable('x', 3)
    new += CyLPArray([0, 0, 0]) <= y <= CyLPArray([10, 10, 1])
    new.addConstraint(CyLPArray([-1.5, -1.25]) <= np.array([[-1., 0, -1], [0, -1, 0]]) * y <= CyLPArray([COIN_INFINITY] * 2), 'R')
    new.objective = CyLPArray([-1, -1, -1]) * y
    new.dual()
    assert new.getStatusCode() == 0 and new.objectiveValue == pytest.approx(-2.75)
    assert list(new.variablesUpper) == [10, 10, 1]


def test_solve_lps_batches_and_caches(monkeypatch):
    eng = use_oracle_engine(monkeypatch)
    lp = _lp()
    kids = [lp.copy_for_child() for _ in range(3)]
    kids[0].variablesUpper[2] = 1
    kids[1].variablesLower[2] = 2
    assert solve_lps([lp] + kids) == 4 and eng.calls == 1 and eng.batch_sizes == [4]
    assert [k.getStatusCode() for k in kids] == [0, 1, 0]
    assert kids[1].objectiveValue == float('inf') and kids[1].primalVariableSolution['x'] is None
    assert solve_lps([lp] + kids) == 0                    # unchanged LPs are cache hits
    lp.dual()
    assert eng.calls == 1
    kids[2].variablesUpper[0] = 0.5                       # an in-place bound change invalidates the cache
    kids[2].dual()
    assert eng.calls == 2
    kids[0].maxNumIteration = 5                           # a different budget goes in its own call
    kids[1].variablesLower[2] = 0
    assert solve_lps(kids) == 2 and eng.calls == 4


def test_milp_instance_forms():
    A = np.array([[1, 0, 1], [0, 1, 0]])
    m = MILPInstance(A=A, b=CyLPArray([1.5, 1.25]), c=CyLPArray([1, 1, 1]), l=CyLPArray([0, 0, 0]),
                     u=CyLPArray([10, 10, 10]), sense=['Max', '<='], integerIndices=[0, 1, 2], numVars=3)
    assert m.sense == '<=' and list(m.lp.objective) == [-1, -1, -1]          # maximisation is negated
    g = MILPInstance(A=-A, b=-CyLPArray([1.5, 1.25]), c=-CyLPArray([1, 1, 1]), sense=['Min', '>='],
                     integerIndices=[0], numVars=3)
    assert g.sense == '>=' and (g.u >= 1e300).all() and (g.l == 0).all()
    assert isinstance(g.lp, CyClpSimplex) and g.lp.nConstraints == 2


def test_mps_reader(tmp_path):
    text = """NAME          BLANK
ROWS
 N  OBJROW
 L  R_1_0
 L  R_1_1
COLUMNS
    x_0  OBJROW  -3.  R_1_0  7.
    x_0  R_1_1  2.
    x_1  OBJROW  -5.
    x_2  OBJROW  -1.  R_1_1  4.
RHS
    RHS  R_1_0  9.  R_1_1  6.
BOUNDS
 UI BOUND  x_0  10.
 UI BOUND  x_1  10.
 UP BOUND  x_2  3.5
ENDATA
"""
    p = tmp_path / 'm.mps'
    p.write_text(text)
    mdl = read_mps(str(p))
    assert mdl.A.shape == (2, 3) and mdl.row_senses == ['L', 'L']
    assert np.array_equal(mdl.A.toarray(), [[7, 0, 0], [2, 0, 4]])           # x_1: an empty column
    assert list(mdl.rhs) == [9, 6] and list(mdl.c) == [-3, -5, -1]
    assert mdl.integer_indices == [0, 1] and list(mdl.u) == [10, 10, 3.5] and list(mdl.l) == [0, 0, 0]
    inst = MILPInstance(file_name=str(p))
    assert inst.sense == '<=' and inst.integerIndices == [0, 1] and inst.numVars == 3


def test_binary_tree():
    t = BinaryTree()
    t.add_root(0, node='r')
    t.add_left_child(1, 0, node='l')
    t.add_right_child(2, 0, node='rr')
    assert 1 in t and 3 not in t and t.get_children(0) == [1, 2] and t.get_parent(2) == 0
    assert t.nodes[1].attr['node'] == 'l' and t.get_node_attr(2, 'node') == 'rr'
    with pytest.raises(AssertionError):
        t.add_left_child(3, 0)
    with pytest.raises(AssertionError):
        t.add_left_child(2, 1)


def test_cut_pool_is_garbage_collected_and_cut_names_are_content_keyed(monkeypatch):
    """Cut rows of finished nodes do not pile up in the device matrix for ever, and a cut NAME that is
    re-used with other coefficients (a second run on the same model) never resolves to the old row."""
    from helpers import use_oracle_engine
    from simple_mip_solver_b200 import CyLPArray, MILPInstance
    from simple_mip_solver_b200.compat import solve_lps
    use_oracle_engine(monkeypatch)
    m = MILPInstance(A=np.array([[-1.0, -1.0]]), b=CyLPArray([-3.5]), c=CyLPArray([-1.0, -1.0]),
                     l=CyLPArray([0, 0]), u=CyLPArray([10, 10]), sense=['Min', '>='], integerIndices=[0, 1], numVars=2)
    sh = m.lp._need_shared()
    sh.pool_cap = 4
    x = m.lp.getVarByName('x')
    vals = []
    for k in range(12):
        lp = m.lp.copy_for_child()
        lp.addConstraint(CyLPArray([-1.0, 0.0]) * x >= -float(k % 3), 'cut_same_name')      # same name, new content
        lp.addConstraint(CyLPArray([0.0, -1.0]) * x >= -(0.5 + k), f'cut_{k}')
        solve_lps([lp])
        vals.append(lp.objectiveValue)
        assert lp.objectiveValue == -min(3.5, float(k % 3) + min(10.0, 0.5 + k))
        assert len(sh.cut_names) <= 6
    assert sh.pool_rebuilds >= 2


def test_cut_generation_dual_bound_dict_messages(monkeypatch):
    """Message strings of the kwargs-bus validation (the reference's tests match on them,
    test_base_node.py:176-313)."""
    from helpers import use_oracle_engine
    from simple_mip_solver_b200 import BaseNode, CyLPArray, MILPInstance
    use_oracle_engine(monkeypatch)
    m = MILPInstance(A=np.array([[-1.0, -1.0]]), b=CyLPArray([-3.5]), c=CyLPArray([-1.0, -1.0]),
                     l=CyLPArray([0, 0]), u=CyLPArray([10, 10]), sense=['Min', '>='], integerIndices=[0, 1], numVars=2)
    node = BaseNode(m.lp, m.integerIndices, idx=4)
    good = node._good_cut_generation_dual_bound_dict
    assert good({1: {0: -3.0, 1: -2.5}, 2: {0: 1}}) == (True, None)
    assert good([]) == (False, 'cut_generation_dual_bound_dict should be a dictionary')
    assert good({'a': {0: 1.0}}) == (False, 'index a should be integer')
    assert good({4: {0: 1.0}}) == (False, 'index 4 has already been processed')
    assert good({1: [1.0]}) == (False, 'index 1 should have dictionary value')
    assert good({1: {'x': 1.0}}) == (False, 'cut index x for node 1 should be integer')
    assert good({1: {0: 'v'}}) == (False, 'dual bound for node 1 cut index 0 should be a number')
    assert good({1: {0: 1.0, 2: 2.0}}) == (False, 'index 1 should have dictionary keyed by range of ints')
    with pytest.raises(AssertionError, match='index 4 has already been processed'):
        node._base_bound(cut_generation_dual_bound_dict={4: {0: 1.0}})


def test_models_the_device_cannot_hold_are_kept_for_inspection_and_rejected(monkeypatch):
    """CyLP models may hold rows with an upper bound, further variable vectors and '<=' rows; the look-alike
    keeps them so that the Node layer rejects them with the reference's own messages
    (base_node.py:111-112, 683-710; test_base_node.py:123-125, 383-386, 867-905), and refuses to solve them."""
    from helpers import use_oracle_engine
    from simple_mip_solver_b200 import BaseNode, CyLPArray, MILPInstance
    from simple_mip_solver_b200.compat import solve_lps
    use_oracle_engine(monkeypatch)

    def small(sense='>='):
        A = np.array([[1.0, 0.0, 1.0], [0.0, 1.0, 0.0]])
        sgn = -1.0 if sense == '>=' else 1.0
        return MILPInstance(A=sgn * A, b=sgn * CyLPArray([1.5, 1.25]), c=-CyLPArray([1, 1, 1]), l=CyLPArray([0, 0, 0]),
                            u=CyLPArray([10, 10, 10]), sense=['Min', sense], integerIndices=[0, 1, 2], numVars=3)

    # a '<=' model: a real LP object with rows bounded from above; the node refuses it by its sense
    m = small('<=')
    assert m.lp.nConstraints == 2 and (m.lp.constraintsUpper == [1.5, 1.25]).all() and not m.lp.solvable
    assert (np.asarray(m.lp.constraintsLower) <= -1e300).all()
    with pytest.raises(AssertionError, match='must have Ax >= b'):
        BaseNode(lp=m.lp, integer_indices=m.integerIndices)
    with pytest.raises(AssertionError, match='the device solves A x >= b'):
        solve_lps([m.lp])

    # a row with an upper bound on a '>=' model: mixed senses
    m = small()
    node = BaseNode(m.lp, m.integerIndices, 0)
    x = node.lp.getVarByName('x')
    node.lp.addConstraint(CyLPArray([1, 0, 0]) * x <= 1, 'upper')
    assert node.lp.nConstraints == 3
    with pytest.raises(AssertionError, match='all constraints should be bounded same way'):
        node._sense
    node.lp.removeConstraint('upper')
    assert node._sense == '>=' and node.lp.solvable

    # all rows replaced by '<=' rows (test_base_node.py:873-884)
    A, b = -node.lp.constraints[0].varCoefs[x], -node.lp.constraints[0].lower
    for constr in node.lp.constraints:
        node.lp.removeConstraint(constr.name)
    assert node.lp.nConstraints == 0
    node.lp.addConstraint(A * x <= b)
    assert node._sense == '<=' and node.lp.nConstraints == 2
    with pytest.raises(Exception, match='Constraint "nope" does not exist'):
        node.lp.removeConstraint('nope')

    # a second variable vector
    m = small()
    s = m.lp.addVariable('s', 1)
    m.lp += s >= CyLPArray([0])
    node = BaseNode(m.lp, m.integerIndices, 0)
    assert not node._x_only_variable and node.lp.nVariables == 4 and len(node.lp.variablesLower) == 4
    with pytest.raises(AssertionError, match='x must be our only variable'):
        node._bound_lp()
    with pytest.raises(AssertionError, match='x must be our only variable'):
        node._base_branch(branch_idx=1)


def test_basis_round_trip_and_strong_branch_override(monkeypatch):
    """A child right after _base_branch reports its parent's basis (base_node.py:608, test_base_node.py:748-751);
    an overridden _strong_branch is the one _update_pseudo_costs calls (test_pseudo_cost.py:72-87)."""
    from unittest.mock import patch
    from helpers import use_oracle_engine
    from simple_mip_solver_b200 import BaseNode, CyLPArray, MILPInstance, PseudoCostBranchNode
    use_oracle_engine(monkeypatch)
    A = np.array([[1.0, 0.0, 1.0], [0.0, 1.0, 0.0]])
    mk = lambda: MILPInstance(A=-A, b=-CyLPArray([1.5, 1.25]), c=-CyLPArray([1, 1, 1]), l=CyLPArray([0, 0, 0]),
                              u=CyLPArray([10, 10, 10]), sense=['Min', '>='], integerIndices=[0, 1, 2], numVars=3)
    m = mk()
    node = BaseNode(m.lp, m.integerIndices, 0)
    node.bound(gomory_cuts=False)
    kids = node._base_branch(2, 1)
    for d in ('left', 'right'):
        for i in (0, 1):
            assert (node.lp.getBasisStatus()[i] == kids[d].lp.getBasisStatus()[i]).all()
    m = mk()
    node = PseudoCostBranchNode(m.lp, m.integerIndices)
    node.pseudo_costs = {}
    node._base_bound(gomory_cuts=False)
    with patch.object(node, '_strong_branch') as sb, patch.object(node, '_calculate_costs') as cc:
        sb.return_value = {'right': PseudoCostBranchNode(mk().lp, [0, 1, 2]), 'left': PseudoCostBranchNode(mk().lp, [0, 1, 2])}
        node._update_pseudo_costs()
        assert sb.call_count == 2 and cc.call_count == 4


def test_integer_index_checks_run_once_per_model_but_still_catch_a_new_list(monkeypatch):
    """Children carry the list their parent was checked with, so the O(|I|) checks are skipped for them;
    any other list is checked as before (test_base_node.py:80-92)."""
    from helpers import use_oracle_engine
    from simple_mip_solver_b200 import BaseNode, CyLPArray, MILPInstance
    use_oracle_engine(monkeypatch)
    A = np.array([[1.0, 0.0, 1.0], [0.0, 1.0, 0.0]])
    m = MILPInstance(A=-A, b=-CyLPArray([1.5, 1.25]), c=-CyLPArray([1, 1, 1]), l=CyLPArray([0, 0, 0]),
                     u=CyLPArray([10, 10, 10]), sense=['Min', '>='], integerIndices=[0, 1, 2], numVars=3)
    node = BaseNode(m.lp, m.integerIndices, 0)
    assert m.lp.integer_index_set == frozenset({0, 1, 2})
    node.bound(gomory_cuts=False)
    kids = node._base_branch(2, 1)
    assert kids['left'].lp.integer_index_set is m.lp.integer_index_set
    assert kids['left']._integer_indices is m.integerIndices
    with pytest.raises(AssertionError, match='indices must match variables'):
        BaseNode(kids['left'].lp, [0, 1, 5])
    with pytest.raises(AssertionError, match='indices must be distinct'):
        BaseNode(kids['left'].lp, [0, 1, 1])
    with pytest.raises(AssertionError, match='branch index corresponds to integer variable'):
        BaseNode(kids['right'].lp, m.integerIndices, b_idx=4, b_dir='right', b_val=.5)
    assert node.max_term == 1.0 and node.max_term.shape == ()
