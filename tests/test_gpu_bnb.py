"""GPU: BranchAndBound / Node API on the CUDA bound step against the reference's answers.

Config 1/2 of BASELINE.json: the hand-written example models and the 64 ``scale_1_models`` fixtures.
Golden answers come from the unmodified reference run on an exact simplex
(tests/golden/make_goldens.py). Bar: same status, MIP optimum within 1e-6 relative, an integral
optimal solution; per-node LP values within 1e-6 of the reference's wherever the trees coincide.
"""
import json
import os

import numpy as np
import pytest

from simple_mip_solver_b200 import (BaseNode, BranchAndBound, CyLPArray, DepthFirstSearchNode,
                                    MILPInstance, PseudoCostBranchDepthFirstSearchNode,
                                    PseudoCostBranchNode)
from simple_mip_solver_b200.compat import solve_lps

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))
CASES = {
    'BaseNode': (BaseNode, dict(gomory_cuts=False)),
    'DepthFirstSearchNode': (DepthFirstSearchNode, dict(gomory_cuts=False)),
    'PseudoCostBranchNode': (PseudoCostBranchNode, dict(pseudo_costs={}, gomory_cuts=False)),
    'PseudoCostBranchDepthFirstSearchNode': (PseudoCostBranchDepthFirstSearchNode,
                                             dict(pseudo_costs={}, gomory_cuts=False)),
}


def unfl(v):
    return {'inf': float('inf'), '-inf': -float('inf')}.get(v, v) if isinstance(v, str) else v


def model_from(rec):
    return MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']),
                        l=CyLPArray(rec['l']), u=CyLPArray(rec['u']), sense=['Min', '>='],
                        integerIndices=list(rec['integer_indices']), numVars=len(rec['c']))


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def solve(rec, label):
    Node, kw = CASES[label]
    kwargs = {k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}
    bb = BranchAndBound(model_from(rec), Node, **kwargs)
    bb.solve()
    return bb


def same_tree(bb, gold):
    if set(map(str, bb.tree.nodes)) != set(gold['tree']):
        return False
    for idx, (parent, b_idx, b_dir, obj, feas, mipf) in gold['tree'].items():
        n = bb.tree.get_node_instances(int(idx))
        if bb.tree.get_parent(int(idx)) != parent or (n._b_idx, n._b_dir) != (b_idx, b_dir):
            return False
    return True


def test_root_lps_of_all_scale_1_models_in_one_batch(blp_lib):
    """64 different MILPs: each root LP through BaseNode._bound_lp, objective vs reference."""
    for name, rec in SCALE1.items():
        m = model_from(rec)
        node = BaseNode(m.lp, m.integerIndices, idx=0)
        node._bound_lp()
        assert node.lp_feasible and rel(node.objective_value, rec['root_lp']['objective']) <= 1e-6, name
        assert node.mip_feasible == rec['root_lp']['mip_feasible'] or not rec['root_lp']['mip_feasible'], name
        m.lp._shared.close()


def check_tree_values(bb, gold, name):
    for idx, (_, _, _, obj, feas, mipf) in gold['tree'].items():
        n = bb.tree.get_node_instances(int(idx))
        assert n.lp_feasible == feas and n.mip_feasible == mipf, (name, idx)
        if feas and n.objective_value is not None:
            assert rel(n.objective_value, unfl(obj)) <= 1e-9, (name, idx)


@pytest.mark.parametrize('label', list(CASES))
def test_scale_1_models_optimum(blp_lib, label):
    """Config 2. The node LPs of these fixtures go through the dual simplex kernels, so the search must
    be the reference's search: the SAME TREE (parent, branching variable, direction, LP value of every
    node) on 64/64 fixtures as the unmodified reference run on the same textbook dual simplex
    (``reference_ds``), and the same optimum as the reference run on HiGHS (``reference``), where the
    tree may differ only through alternative optimal vertices / last-ulp ties of the most fractional
    rule."""
    matches = highs_matches = 0
    for name, rec in SCALE1.items():
        gold, gold_h = rec['reference_ds'][label], rec['reference'][label]
        bb = solve(rec, label)
        assert bb.status == gold['status'] == gold_h['status'] == 'optimal', name
        assert rel(bb.objective_value, unfl(gold_h['objective'])) <= 1e-9, (name, bb.objective_value)
        assert rel(bb.objective_value, rec['mip_optimum']) <= 1e-9, name
        assert bb.evaluated_nodes == gold['evaluated_nodes'], name
        ints = rec['integer_indices']
        sol = np.asarray(bb.solution)
        assert np.max(np.abs(sol[ints] - np.round(sol[ints]))) <= 1e-9
        assert np.allclose(sol, gold['solution'], atol=1e-9), name
        assert (np.array(rec['A']) @ sol >= np.array(rec['b']) - 1e-9).all()
        assert same_tree(bb, gold), name
        matches += 1
        check_tree_values(bb, gold, name)
        if same_tree(bb, gold_h):
            highs_matches += 1
            check_tree_values(bb, gold_h, name)
        bb.model.lp._shared.close()
    print(f'{label}: identical trees on {matches}/{len(SCALE1)} instances (reference on the textbook dual '
          f'simplex), {highs_matches}/{len(SCALE1)} (reference on HiGHS)')
    assert matches == len(SCALE1)
    assert highs_matches >= 60


@pytest.mark.parametrize('label', list(CASES))
def test_scale_1_models_optimum_pdhg_path(blp_lib, label, monkeypatch):
    """The same fixtures forced through the PDHG kernels (the path of LPs too large for the simplex):
    status, optimum and an integral solution; trees may differ on degenerate faces."""
    from simple_mip_solver_b200.compat.cylp_like import SharedLP
    monkeypatch.setattr(SharedLP, 'default_method', 'pdhg')
    matches = 0
    for name, rec in list(SCALE1.items())[::4]:
        gold = rec['reference'][label]
        bb = solve(rec, label)
        assert bb.status == gold['status'] == 'optimal', name
        assert rel(bb.objective_value, unfl(gold['objective'])) <= 1e-6, (name, bb.objective_value)
        ints = rec['integer_indices']
        sol = np.asarray(bb.solution)
        assert np.max(np.abs(sol[ints] - np.round(sol[ints]))) <= 1e-4
        assert (np.array(rec['A']) @ sol >= np.array(rec['b']) - 1e-5).all()
        matches += same_tree(bb, gold)
        bb.model.lp._shared.close()
    print(f'{label} (PDHG path): identical trees on {matches}/16 instances')
    assert matches >= 12


@pytest.mark.parametrize('name', ['no_branch', 'small_branch', 'infeasible', 'infeasible2', 'random', 'cut1',
                                  'cut2', 'cut3', 'square', 'h3p1', 'h3p1_0', 'h3p1_1', 'h3p1_2', 'h3p1_3',
                                  'h3p1_4', 'h3p1_5', 'lift_project'])
@pytest.mark.parametrize('label', ['BaseNode', 'PseudoCostBranchNode'])
def test_example_models(blp_lib, name, label):
    rec = EXAMPLES[name]
    gold = rec['reference_ds'][label]
    bb = solve(rec, label)
    assert bb.status == gold['status'] == rec['reference'][label]['status'], (bb.status, gold['status'])
    assert same_tree(bb, gold)
    assert bb.evaluated_nodes == gold['evaluated_nodes']
    if gold['status'] == 'optimal':
        assert rel(bb.objective_value, unfl(rec['reference'][label]['objective'])) <= 1e-9
        ints = rec['integer_indices']
        sol = np.asarray(bb.solution)
        assert np.max(np.abs(sol[ints] - np.round(sol[ints]))) <= 1e-4
    else:
        assert bb.objective_value == float('inf')
    bb.model.lp._shared.close()


def test_unbounded_root(blp_lib):
    rec = EXAMPLES['unbounded']
    m = model_from(rec)
    node = BaseNode(m.lp, m.integerIndices, idx=0)
    node._bound_lp()
    assert node.lp_feasible and node.unbounded          # test_base_node.py:430-437
    bb = BranchAndBound(model_from(rec), BaseNode, gomory_cuts=False)
    bb.solve()
    assert bb.status == 'unbounded'                     # test_branch_and_bound.py:306-322


def test_small_branch_node_api_pins(blp_lib):
    m = model_from(EXAMPLES['small_branch'])
    node = BaseNode(m.lp, m.integerIndices, idx=0)
    node._bound_lp()
    assert rel(node.objective_value, -2.75) <= 1e-6 and node.lp_feasible and not node.mip_feasible
    assert list(node.solution) == [0.0, 1.25, 1.5]                          # test_base_node.py:406-416
    assert node._most_fractional_index == 2                                 # :824-826
    assert node.tableau is not None
    kids = node._strong_branch_batch([1], iterations=5)[1]
    assert kids['left'].lp.getStatusCode() == 0 and rel(kids['left'].lp.objectiveValue, -2.5) <= 1e-6
    assert kids['right'].lp.getStatusCode() == 1
    sh = m.lp._shared
    assert sh.solve_calls == 2 and sh.lps_solved == 3       # root alone, then both children in one call


def test_frontier_batches_reach_gpu(blp_lib):
    rec = EXAMPLES['random']
    bb = BranchAndBound(model_from(rec), BaseNode, frontier_batch=16, gomory_cuts=False)
    bb.solve()
    gold = rec['reference']['BaseNode']
    assert bb.status == 'optimal' and rel(bb.objective_value, unfl(gold['objective'])) <= 1e-6
    sh = bb.model.lp._shared
    assert sh.solve_calls < sh.lps_solved and sh.kernel_launches > 0
    print('random 20x10:', bb.evaluated_nodes, 'nodes (reference', gold['evaluated_nodes'], '),',
          sh.lps_solved, 'LPs in', sh.solve_calls, 'GPU calls')


def test_cut_rows_through_node_api(blp_lib):
    """Append cut rows through lp.addConstraint on some nodes only and re-solve as one batch."""
    rec = EXAMPLES['random']
    m = model_from(rec)
    root = BaseNode(m.lp, m.integerIndices, idx=0)
    root._bound_lp()
    base = root.objective_value
    kids = root._base_branch(root._most_fractional_index, next_node_idx=1)
    x = m.lp.getVarByName('x')
    pi = CyLPArray(-np.ones(len(rec['c'])))
    rhs = -float(np.floor(np.sum(root.solution) - 0.5))
    kids['left'].lp.addConstraint(pi * x >= rhs, 'cut_test_0')
    assert solve_lps([kids['left'].lp, kids['right'].lp]) == 2
    for k in kids.values():
        if not isinstance(k, int):
            k._bound_lp()
            if k.lp_feasible:
                assert k.objective_value >= base - 1e-6 * abs(base)
    if kids['left'].lp_feasible:
        assert float(np.dot(pi, kids['left'].solution)) >= rhs - 1e-5 * max(1.0, abs(rhs))   # solver tolerance is relative
        assert 'cut_test_0' in kids['left'].lp.dualConstraintSolution
    assert 'cut_test_0' not in kids['right'].lp.dualConstraintSolution


def test_default_bound_with_gomory_rounds(blp_lib):
    """BaseNode.bound() with the reference's defaults (gomory_cuts=True, up to 10 cut rounds,
    base_node.py:137-230, 365): cut rows are generated from the active-set basis of the GPU
    solution, appended through lp.addConstraint and re-solved on the GPU. The MIP optimum must be
    the reference's; how many rounds help depends on the vertex and is not pinned."""
    used_cuts = 0
    for name, rec in list(SCALE1.items()) + [(k, EXAMPLES[k]) for k in ('small_branch', 'cut1', 'cut2', 'cut3', 'square')]:
        gold = rec['reference']['BaseNode_gomory']
        bb = BranchAndBound(model_from(rec), BaseNode)
        bb.solve()
        assert bb.status == gold['status'] == 'optimal', name
        assert rel(bb.objective_value, unfl(gold['objective'])) <= 1e-6, (name, bb.objective_value, gold['objective'])
        used_cuts += bb._kwargs.get('total_number_gmic_added', 0)
        bb.model.lp._shared.close()
    print('GMI cuts appended over all instances:', used_cuts)
    assert used_cuts > 0


def test_disjunctive_cut_nodes(blp_lib):
    """DisjunctiveCutBoundNode as a caller of the GPU bound step; its CGLP is solved by the same
    engine (reference usage: test_simple_mip_solver/helpers.py:75-126)."""
    from simple_mip_solver_b200 import (CutGeneratingLP, DisjunctiveCutBoundNode,
                                        DisjunctiveCutBoundPseudoCostBranchNode)
    recs = list(SCALE1.items())[::16] + [(k, EXAMPLES[k]) for k in ('cut1', 'cut2', 'lift_project')]
    created = 0
    for name, rec in recs:
        tree = BranchAndBound(model_from(rec), BaseNode, node_limit=8, gomory_cuts=False)
        tree.solve()
        cglp = CutGeneratingLP(tree, tree.root_node.idx)
        pi, pi0 = cglp.solve()
        assert pi is not None
        # validity on the integer points of the model (small boxes only)
        A, b = np.array(rec['A']), np.array(rec['b'])
        if len(rec['c']) <= 4:
            import itertools
            for p in itertools.product(*[range(int(lo), int(min(hi, lo + 6)) + 1) for lo, hi in zip(rec['l'], rec['u'])]):
                p = np.array(p, dtype=float)
                if (A @ p >= b - 1e-9).all():
                    assert float(np.dot(pi, p)) >= pi0 - 1e-5 * max(1.0, abs(pi0)), (name, p)
        for Node, extra in ((DisjunctiveCutBoundNode, dict(gomory_cuts=False)),
                            (DisjunctiveCutBoundPseudoCostBranchNode, dict(pseudo_costs={}, gomory_cuts=False))):
            bb = BranchAndBound(model_from(rec), Node, cglp=cglp, **extra)
            bb.solve()
            want = rec.get('mip_optimum', rec['reference']['BaseNode']['objective'])
            assert bb.status == 'optimal' and rel(bb.objective_value, unfl(want)) <= 1e-6, (name, bb.objective_value, want)
            created += bb._kwargs['total_number_cglp_created']
            bb.model.lp._shared.close()
    assert created > 0
