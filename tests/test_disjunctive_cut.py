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


# ---------------------------------------------------------------- batched CGLP
def cold(cglp):
    return (np.full(cglp.lp.nVariables, 3, dtype=np.int32), np.full(cglp.lp.nConstraints, 1, dtype=np.int32))


def test_cglp_known_answers_of_the_reference(monkeypatch):
    """test_cut_generating_lp.py:371-415 of the reference: the cuts its CGLP finds on `square` and on
    small_branch, and the one that cuts off a point from above."""
    use_oracle_engine(monkeypatch)
    bb = BranchAndBound(model_from(EXAMPLES['square']), BaseNode, gomory_cuts=False)
    bb.solve()
    pi, pi0 = CutGeneratingLP(bb, bb.root_node.idx).solve()
    want = [0, 1] if abs(pi[1]) > abs(pi[0]) else [1, 0]
    np.testing.assert_allclose(pi / pi0, want, atol=.01)
    assert (pi - .01 < 0).all() and pi0 - .01 < 0

    bb = partial_tree(EXAMPLES['small_branch'], node_limit=10)
    pi, pi0 = CutGeneratingLP(bb, bb.root_node.idx).solve()
    np.testing.assert_allclose(pi / pi0, [0, 0, 1], atol=.01)

    bb = partial_tree(EXAMPLES['square'], node_limit=1)
    cglp = CutGeneratingLP(bb, bb.root_node.idx)
    pi, pi0 = cglp.solve(x_star=CyLPArray([1.5, 2]))
    assert pi0 == pytest.approx(-.75, abs=.01)
    np.testing.assert_allclose(pi, [0, -.5], atol=.01)
    assert cglp.lp.objectiveValue == pytest.approx(float(np.dot(pi, [1.5, 2])) - pi0, abs=1e-9)
    pi, pi0 = cglp.solve(x_star=CyLPArray([.5, .5]))        # a point that cannot be separated: still an answer
    assert pi is not None and pi0 is not None and cglp.lp.objectiveValue >= -1e-9


@pytest.mark.parametrize('name', ['cut1', 'cut2', 'lift_project', 'square', 'small_branch', 'random'])
def test_cglp_batch_is_the_single_solves(monkeypatch, name):
    """K points through ONE device call give, point for point, what K single solves give, and the
    optimum of every LP is the optimum of the reference's primal model (HiGHS on `cglp._M`)."""
    from oracle.highs_lp import HIGHS_INF, HighsLP
    oracle = use_oracle_engine(monkeypatch)
    bb = partial_tree(EXAMPLES[name])
    root = bb.root_node
    cglp = CutGeneratingLP(bb, root.idx)
    x = np.asarray(root.solution, dtype=float)
    rng = np.random.default_rng(5)
    points = [CyLPArray(x)] + [CyLPArray(np.maximum(x * rng.uniform(.7, 1.2, len(x)), 0)) for _ in range(6)]
    before = list(oracle.batch_sizes)
    cuts = cglp.solve_batch(points, starting_bases=[cold(cglp)] * len(points))
    assert oracle.batch_sizes[len(before):] == [len(points)] and cglp.batch_calls == 1
    fin = lambda v, big: np.where(np.isinf(v), big, v)
    for p, (pi, pi0) in zip(points, cuts):
        one = CutGeneratingLP(bb, root.idx)
        pi1, pi01 = one.solve(x_star=p, starting_basis=cold(one))
        assert np.array_equal(pi, pi1) and pi0 == pi01
        c = np.zeros(cglp.lp.nVariables)
        c[:cglp.n], c[cglp.n] = p, -1.0
        r = HighsLP(cglp._M, c, cglp._r, np.full(cglp._M.shape[0], HIGHS_INF),
                    fin(cglp._lo, -HIGHS_INF), fin(cglp._hi, HIGHS_INF)).solve()
        assert r.status == 0
        assert float(np.dot(pi, p)) - pi0 == pytest.approx(r.objective, abs=1e-8)
        # valid for every integer point (small boxes) / for the LP optimum of every disjunctive term
        for q in integer_points(EXAMPLES[name]) if len(x) <= 4 else []:
            assert float(np.dot(pi, q)) >= pi0 - 1e-7
        for leaf in bb.tree.get_leaves(root.idx):
            if leaf.lp_feasible and leaf.solution is not None:
                lo = np.maximum(np.asarray(leaf.lp.variablesLower), 0)
                assert float(np.dot(pi, np.maximum(leaf.solution, lo))) >= pi0 - 1e-6


def test_cglp_basis_round_trip(monkeypatch):
    """test_cut_generating_lp.py:417-428: a solve started from the basis of an earlier one needs no
    iteration; the arrays have the reference model's shapes and survive a plain-array copy."""
    use_oracle_engine(monkeypatch)
    bb = partial_tree(EXAMPLES['small_branch'], node_limit=10)
    cglp = CutGeneratingLP(bb, bb.root_node.idx)
    pi, pi0 = cglp.solve()
    assert cglp.lp.iteration > 0
    cols, rows = cglp.lp.getBasisStatus()
    assert cols.shape == (cglp.lp.nVariables,) and rows.shape == (cglp.lp.nConstraints,)
    assert set(np.unique(cols)) <= {1, 3} and set(np.unique(rows)) <= {1, 3}
    again = CutGeneratingLP(bb, bb.root_node.idx)
    pi2, pi02 = again.solve(starting_basis=(cols, rows))
    assert again.lp.iteration == 0 and np.allclose(pi, pi2) and pi0 == pytest.approx(pi02)
    plain = CutGeneratingLP(bb, bb.root_node.idx)            # CLP-coded arrays without the device basis
    pi3, pi03 = plain.solve(starting_basis=(np.array(cols), np.array(rows)))
    assert plain.lp.iteration < cglp.lp.iteration and np.allclose(pi, pi3)
    with pytest.raises(AssertionError, match='first starting_basis element'):
        cglp.solve(starting_basis=(np.append(cols, 1), rows))
    with pytest.raises(AssertionError, match='second starting_basis element'):
        cglp.solve(starting_basis=(cols, np.append(rows, 1)))


def test_frontier_prefetch_batches_the_first_disjunctive_cut(monkeypatch):
    """With a frontier batch the nodes that share a CGLP get their first cut from one batched call
    (DisjunctiveCutBoundNode.prefetch); the optimum is the one-node-at-a-time optimum."""
    use_oracle_engine(monkeypatch)
    hits = batched = 0
    for name, rec in list(SCALE1.items())[::9] + [('cut2', EXAMPLES['cut2']), ('random', EXAMPLES['random'])]:
        tree = partial_tree(rec)
        want = rec.get('mip_optimum', rec['reference']['BaseNode']['objective'])
        for fb in (1, 8):
            cglp = CutGeneratingLP(tree, tree.root_node.idx)
            bb = BranchAndBound(model_from(rec), DisjunctiveCutBoundNode, cglp=cglp,
                                gomory_cuts=False, frontier_batch=fb)
            bb.solve()
            assert bb.status == 'optimal' and bb.objective_value == pytest.approx(want, abs=1e-6), (name, fb)
            assert cglp.points_solved >= cglp.batch_calls
            if fb == 1:
                assert cglp.prefetch_hits == 0
            else:
                hits += cglp.prefetch_hits
                batched += cglp.points_solved - cglp.batch_calls
    assert hits > 0 and batched > 0


def test_cglp_first_order_path_gives_the_same_optimum(monkeypatch):
    """CGLPs too large for the simplex kernel go to the engine's first-order path (here: HiGHS behind the
    same call); the cut is read from the row duals in the same way."""
    use_oracle_engine(monkeypatch)
    bb = partial_tree(EXAMPLES['random'])
    exact = CutGeneratingLP(bb, bb.root_node.idx)
    pi, pi0 = exact.solve()
    first_order = CutGeneratingLP(bb, bb.root_node.idx)
    first_order.method = 'pdhg'
    x = np.asarray(bb.root_node.solution, dtype=float)
    (pi1, pi01), (pi2, pi02) = first_order.solve_batch([CyLPArray(x), CyLPArray(x * .9)])
    assert first_order.batch_calls == 1 and first_order.lp._basis is None
    assert float(np.dot(pi1, x)) - pi01 == pytest.approx(exact.lp.objectiveValue, abs=1e-8)
    assert float(np.dot(pi1, x)) - pi01 == pytest.approx(float(np.dot(pi, x)) - pi0, abs=1e-8)
    for leaf in bb.tree.get_leaves(bb.root_node.idx):
        if leaf.lp_feasible and leaf.solution is not None:
            for p, p0 in ((pi1, pi01), (pi2, pi02)):
                assert float(np.dot(p, np.maximum(leaf.solution, 0))) >= p0 - 1e-6


def test_cglp_points_with_round_off_below_zero_and_outside_the_orthant(monkeypatch):
    use_oracle_engine(monkeypatch)
    bb = partial_tree(EXAMPLES['small_branch'], node_limit=10)
    cglp = CutGeneratingLP(bb, bb.root_node.idx)
    x = np.asarray(bb.root_node.solution, dtype=float)
    pi, pi0 = cglp.solve()
    noisy = x.copy()
    noisy[0] = -3e-13                                   # x_0 = 0 at the root, read back with round-off
    pi1, pi01 = cglp.solve(x_star=CyLPArray(noisy), starting_basis=cold(cglp))
    assert pi1 is not None and np.allclose(pi1, pi) and pi01 == pytest.approx(pi0)
    # a point outside x >= 0 makes the CGLP unbounded (reference: CLP status 2): reported as a failure
    far = x.copy()
    far[0] = -1.0
    assert cglp.solve(x_star=CyLPArray(far)) == (None, None)
    assert cglp.cylp_failure and cglp.lp.getStatusCode() == 2


def test_cglp_dual_form_optimum_on_every_fixture(monkeypatch):
    """All 64 scale_1 fixtures: the cut read from the device LP's multipliers attains the optimum of the
    reference's primal CGLP model (HiGHS), is valid for every integer point, and never fails."""
    from oracle.highs_lp import HIGHS_INF, HighsLP
    use_oracle_engine(monkeypatch)
    fin = lambda v, big: np.where(np.isinf(v), big, v)
    solved = separated = 0
    for name, rec in SCALE1.items():
        bb = partial_tree(rec, node_limit=6)
        root = bb.root_node
        if root.solution is None:
            continue
        cglp = CutGeneratingLP(bb, root.idx)
        if not cglp.term_ids:
            continue
        x = np.asarray(root.solution, dtype=float)
        pts = [CyLPArray(x), CyLPArray(np.maximum(x * 0.8, 0)), CyLPArray(np.maximum(x + 0.3, 0))]
        pool = integer_points(rec) if len(x) <= 4 else []
        for p, (pi, pi0) in zip(pts, cglp.solve_batch(pts)):
            assert pi is not None, name
            c = np.zeros(cglp.lp.nVariables)
            c[:cglp.n], c[cglp.n] = p, -1.0
            M, r, lo, hi = cglp.primal_model()
            h = HighsLP(M, c, r, np.full(M.shape[0], HIGHS_INF), fin(lo, -HIGHS_INF), fin(hi, HIGHS_INF)).solve()
            assert h.status == 0 and float(np.dot(pi, p)) - pi0 == pytest.approx(h.objective, abs=1e-7), name
            for q in pool:
                assert float(np.dot(pi, q)) >= pi0 - 1e-6 * max(1.0, abs(pi0)), (name, q)
            solved += 1
            separated += float(np.dot(pi, p)) - pi0 < -1e-9
    assert solved >= 150 and separated >= 40


# ---------------------------------------------------------------- goldens made by the unmodified reference
CGLP_GOLD = json.load(open(os.path.join(GOLD, 'cglp.json')))


def check_against_reference_cglp(name, gold, rec):
    """The product's disjunction, LP size and optima against what the UNMODIFIED reference's
    CutGeneratingLP produced on the same model (tests/golden/make_cglp_goldens.py)."""
    bb = partial_tree(rec, node_limit=gold['node_limit'])
    root = bb.root_node
    terms = {str(n.idx): n for n in bb.tree.get_leaves(root.idx) if n.lp_feasible is not False}
    assert set(terms) == set(gold['terms']), name                              # the same disjunction ...
    for idx, n in terms.items():
        assert np.array_equal(np.asarray(n.lp.variablesLower), gold['terms'][idx]['lower']), (name, idx)
        assert np.array_equal(np.minimum(np.asarray(n.lp.variablesUpper), 1e308), gold['terms'][idx]['upper'])
    cglp = CutGeneratingLP(bb, root.idx)
    assert cglp.lp.nVariables == gold['n_variables'] and cglp.lp.nConstraints == gold['n_constraints'], name
    points = [CyLPArray(q['x_star']) for q in gold['points']]
    assert np.allclose(points[0], np.asarray(root.solution), rtol=1e-12, atol=1e-12)    # ... and the same root vertex
    cuts = cglp.solve_batch(points)
    for q, p, (pi, pi0) in zip(gold['points'], points, cuts):
        assert q['status'] == 0 and pi is not None, name
        mine = float(np.dot(pi, p)) - pi0
        assert mine == pytest.approx(q['optimum'], abs=1e-8 * (1 + abs(q['optimum']))), (name, mine, q['optimum'])
        # the reference's own cut and ours are both valid for every term's LP optimum
        for n in terms.values():
            if n.lp_feasible and n.solution is not None:
                z = np.maximum(np.asarray(n.solution, dtype=float), 0)
                assert float(np.dot(pi, z)) >= pi0 - 1e-6 and float(np.dot(q['pi'], z)) >= q['pi0'] - 1e-6
    return cglp, bb


@pytest.mark.parametrize('name', sorted(CGLP_GOLD))
def test_cglp_against_the_unmodified_reference(monkeypatch, name):
    use_oracle_engine(monkeypatch)
    rec = EXAMPLES[name] if name in EXAMPLES else SCALE1[name]
    gold = CGLP_GOLD[name]
    _, bb = check_against_reference_cglp(name, gold, rec)
    # the constructor's options (reference :13-50), as the generator applied them
    root = bb.root_node
    x = np.asarray(root.solution, dtype=float)
    options = dict(
        depth_1=dict(depth=1),
        root_rows=dict(A=root.lp.coefMatrix.copy(), b=CyLPArray(np.asarray(root.lp.constraintsLower).copy())),
        unit_box=dict(var_lb=CyLPArray(np.floor(x)), var_ub=CyLPArray(np.floor(x) + 1)),
    )
    for key, kw in options.items():
        want = gold['variants'][key]
        other = CutGeneratingLP(bb, root.idx, **kw)
        if 'error' in want:
            # no term survives (the reference trips over its empty term list there): nothing to separate with
            assert not other.term_ids and other.solve(x_star=CyLPArray(x)) == (None, None) and other.cylp_failure
            continue
        assert (other.lp.nVariables, other.lp.nConstraints) == (want['n_variables'], want['n_constraints']), (name, key)
        pi, pi0 = other.solve(x_star=CyLPArray(x))
        assert want['status'] == 0 and pi is not None
        assert float(np.dot(pi, x)) - pi0 == pytest.approx(want['optimum'], abs=1e-8 * (1 + abs(want['optimum']))), \
            (name, key)
