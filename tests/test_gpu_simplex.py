"""GPU: the batched dual simplex kernels (blp_simplex_*, csrc/blp_simplex.cuh) against the oracles.

Two checkers: HiGHS (oracle/highs_lp.py) for optimal values and statuses, and the numpy restatement
of the device algorithm (oracle/dual_simplex.py) for everything a simplex is asked beyond the value —
the vertex, the basis, the pivot count, the iteration-limited objective. The restatement runs the
same pivoting rules with the same single-rounded operations in the same order, so the comparison is
EXACT (np.array_equal), not a tolerance.
"""
import json
import os

import numpy as np
import pytest

from oracle.dual_simplex import dual_simplex
from oracle.highs_lp import HIGHS_INF, HighsLP
from simple_mip_solver_b200 import engine
from simple_mip_solver_b200.instances import grumpy_random_mip

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SCALE1 = json.load(open(os.path.join(GOLD, 'scale_1_models.json')))
EXAMPLES = json.load(open(os.path.join(GOLD, 'example_models.json')))


def arrays(rec):
    return (np.array(rec['A'], float), np.array(rec['b'], float), np.array(rec['c'], float),
            np.array(rec['l'], float), np.array(rec['u'], float))


def same(res, k, ref, what):
    assert res.status[k] == ref.status, (what, res.status[k], ref.status)
    assert res.pivots[k] == ref.pivots, (what, res.pivots[k], ref.pivots)
    assert np.array_equal(res.col_status[k], ref.col_status), what
    assert np.array_equal(res.row_status[k], ref.row_status), what
    if ref.status != 1:
        assert np.array_equal(res.x[k], ref.x), (what, res.x[k] - ref.x)
        assert res.objective[k] == ref.objective, (what, res.objective[k], ref.objective)
        assert np.array_equal(res.y[k], ref.y), what
        assert np.array_equal(res.reduced_costs[k], ref.rc), what


def test_fixture_roots_bit_exact(blp_lib):
    """Root LP of every scale_1 fixture and example model: cold start from the slack basis."""
    for name, rec in list(SCALE1.items()) + [(k, v) for k, v in EXAMPLES.items() if 'root_lp' in v]:
        A, b, c, l, u = arrays(rec)
        lp = engine.BatchLP(A, b, c)
        res = lp.simplex_batch(l[None], u[None])
        ref = dual_simplex(A, b, c, l, u)
        same(res, 0, ref, name)
        if ref.status == 0:
            h = HighsLP(A, c, b, np.full(len(b), HIGHS_INF), l, u).solve()
            assert h.status == 0 and abs(res.objective[0] - h.objective) <= 1e-9 * max(1, abs(h.objective)), name
        lp.close()


def test_reference_pins(blp_lib):
    """test_base_node.py:394-437 on the device: small_branch root x == [0, 1.25, 1.5] exactly."""
    A, b, c, l, u = arrays(EXAMPLES['small_branch'])
    lp = engine.BatchLP(A, b, c)
    r = lp.simplex_batch(l[None], u[None])
    assert r.status[0] == 0 and r.objective[0] == -2.75 and r.x[0].tolist() == [0.0, 1.25, 1.5]
    # children on x2 (test_branch_and_bound.py:187): right infeasible, left -2.75, from the root's basis
    kids = lp.simplex_children(l, u, [[(2, 0.0, 1.0)], [(2, 2.0, 10.0)]], col_status=r.col_status[0],
                               row_status=r.row_status[0])
    assert kids.status.tolist() == [0, 1] and kids.objective[0] == -2.75
    lp.close()
    A, b, c, l, u = arrays(EXAMPLES['no_branch'])
    lp = engine.BatchLP(A, b, c)
    r = lp.simplex_batch(l[None], u[None])
    assert r.objective[0] == -2.0 and r.x[0].tolist() == [1.0, 1.0, 0.0]
    lp.close()
    for name, code in (('infeasible', 1), ('unbounded', 2)):
        A, b, c, l, u = arrays(EXAMPLES[name])
        lp = engine.BatchLP(A, b, c)
        assert lp.simplex_batch(l[None], u[None]).status[0] == code, name
        lp.close()


def _children(d, x, k):
    ints = np.asarray(d.integer_indices)
    frac = np.minimum(x[ints] - np.floor(x[ints]), np.ceil(x[ints]) - x[ints])
    cand = ints[np.argsort(-frac, kind='stable')][:k]
    deltas = []
    for j in cand:
        deltas.append([(int(j), float(d.l[j]), float(np.floor(x[j])))])
        deltas.append([(int(j), float(np.ceil(x[j])), float(d.u[j]))])
    return deltas


@pytest.mark.parametrize('shape', [(40, 20, 0.2), (120, 60, 0.1)])
def test_children_from_status_and_from_stored_factor(blp_lib, shape):
    """Strong-branching children (pseudo_cost.py:57-62): from the parent's basis status (what
    setBasisStatus hands over) and from the parent's stored factor; pivot limit 5 and unlimited."""
    nv, nc, dens = shape
    d = grumpy_random_mip(nv, nc, density=dens, rand_seed=2)
    A = d.A.toarray()
    lp = engine.BatchLP(d.A, d.b, d.c)
    root = lp.simplex_batch(d.l[None], d.u[None])
    ref_root = dual_simplex(A, d.b, d.c, d.l, d.u)
    same(root, 0, ref_root, 'root')
    deltas = _children(d, ref_root.x, 8)
    for limit in (5, 2147483647):
        for cached in (False, True):
            if cached:          # the store holds the last call: solve the root again first
                lp.simplex_batch(d.l[None], d.u[None])
            res = lp.simplex_children(d.l, d.u, deltas, col_status=root.col_status[0],
                                      row_status=root.row_status[0], parent_slot=0 if cached else -1,
                                      max_pivots=limit)
            for k, dl in enumerate(deltas):
                l, u = d.l.copy(), d.u.copy()
                for j, lo, hi in dl:
                    l[j], u[j] = lo, hi
                ref = dual_simplex(A, d.b, d.c, l, u, col_status=ref_root.col_status, row_status=ref_root.row_status,
                                   max_pivots=limit, start=ref_root if cached else None)
                same(res, k, ref, (limit, cached, k))
                if ref.status in (0, 3):
                    assert res.objective[k] >= root.objective[0] - 1e-9      # test_base_node.py:774-785
                    assert res.pivots[k] <= limit
    lp.close()


def test_config3_root_and_children(blp_lib):
    """Config 3 (500 x 300, 10 %): root cold, 128 children; values against HiGHS, pivoting against the
    numpy restatement."""
    d = grumpy_random_mip(500, 300, density=0.1, rand_seed=2)
    A = d.A.toarray()
    lp = engine.BatchLP(d.A, d.b, d.c)
    root = lp.simplex_batch(d.l[None], d.u[None])
    ref_root = dual_simplex(A, d.b, d.c, d.l, d.u)
    same(root, 0, ref_root, 'root')
    h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u)
    hr = h.solve()
    assert abs(root.objective[0] - hr.objective) <= 1e-9 * abs(hr.objective)
    deltas = _children(d, ref_root.x, 64)
    res = lp.simplex_children(d.l, d.u, deltas, col_status=root.col_status[0], row_status=root.row_status[0])
    for k in range(0, len(deltas), 9):
        l, u = d.l.copy(), d.u.copy()
        for j, lo, hi in deltas[k]:
            l[j], u[j] = lo, hi
        ref = dual_simplex(A, d.b, d.c, l, u, col_status=ref_root.col_status, row_status=ref_root.row_status)
        same(res, k, ref, k)
    for k, dl in enumerate(deltas):
        h.set_col_bounds(d.l, d.u)
        for j, lo, hi in dl:
            h.set_one_col_bound(j, lo, hi)
        h.set_basis(hr.col_basis, hr.row_basis)
        r = h.solve()
        assert res.status[k] == r.status, k
        if r.status == 0:
            assert abs(res.objective[k] - r.objective) <= 1e-9 * abs(r.objective), k
    lp.close()


def test_cut_rows_masks_and_tableau(blp_lib):
    """Appended rows switched per node, and tableau rows of a stored factor against inv(B) [A, -I]."""
    d = grumpy_random_mip(30, 15, density=0.3, rand_seed=5)
    A = d.A.toarray()
    lp = engine.BatchLP(d.A, d.b, d.c)
    rng = np.random.default_rng(3)
    cuts = -rng.integers(0, 4, size=(3, d.n)).astype(float)
    rhs = -rng.integers(20, 60, size=3).astype(float)
    lp.append_rows(cuts, rhs)
    masks = np.array([[1, 1, 1], [1, 0, 0], [0, 0, 0], [0, 1, 1]], dtype=np.uint8)
    lb = np.tile(d.l, (4, 1))
    ub = np.tile(d.u, (4, 1))
    res = lp.simplex_batch(lb, ub, row_mask=masks)
    Afull = np.vstack([A, cuts])
    bfull = np.concatenate([d.b, rhs])
    for k in range(4):
        on = np.concatenate([np.ones(d.m, bool), masks[k].astype(bool)])
        ref = dual_simplex(Afull, bfull, d.c, d.l, d.u, row_on=on)
        same(res, k, ref, k)
        sel = np.flatnonzero(on)
        h = HighsLP(Afull[sel], d.c, bfull[sel], np.full(len(sel), HIGHS_INF), d.l, d.u).solve()
        assert abs(res.objective[k] - h.objective) <= 1e-9 * max(1, abs(h.objective))
        basic = np.flatnonzero(np.concatenate([res.col_status[k], res.row_status[k]]) == 1)
        rows = lp.simplex_tableau_rows(k, basic)
        full = np.concatenate([Afull, -np.eye(Afull.shape[0])], axis=1)
        want = np.linalg.solve(full[:, basic], full)
        assert np.allclose(rows, want, atol=1e-9), k
    lp.close()


def test_wide_kernel_is_bit_identical_too(blp_lib):
    """Above 1024 rows the whole GPU works on one node (k_simplex_wide, cooperative grid): same code
    path per element, order-independent reductions, so the numpy restatement is matched exactly again —
    root cold, children from the basis status and from the stored factor."""
    d = grumpy_random_mip(500, 1100, density=0.02, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=7)
    A = d.A.toarray()
    lp = engine.BatchLP(d.A, d.b, d.c)
    assert lp.simplex_capable and not lp.simplex_batched
    root = lp.simplex_batch(d.l[None], d.u[None])
    ref_root = dual_simplex(A, d.b, d.c, d.l, d.u)
    same(root, 0, ref_root, 'wide root')
    h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    assert abs(root.objective[0] - h.objective) <= 1e-9 * max(1.0, abs(h.objective))
    deltas = _children(d, ref_root.x, 2)
    for cached in (False, True):
        if cached:
            lp.simplex_batch(d.l[None], d.u[None])
        res = lp.simplex_children(d.l, d.u, deltas, col_status=root.col_status[0], row_status=root.row_status[0],
                                  parent_slot=0 if cached else -1, max_pivots=5)
        for k, dl in enumerate(deltas):
            l, u = d.l.copy(), d.u.copy()
            for j, lo, hi in dl:
                l[j], u[j] = lo, hi
            ref = dual_simplex(A, d.b, d.c, l, u, col_status=ref_root.col_status, row_status=ref_root.row_status,
                               max_pivots=5, start=ref_root if cached else None)
            same(res, k, ref, ('wide child', cached, k))
    basic = np.flatnonzero(np.concatenate([res.col_status[0], res.row_status[0]]) == 1)[:5]
    rows = lp.simplex_tableau_rows(0, basic)
    assert rows.shape == (5, d.n + d.m) and np.isfinite(rows).all()
    print('wide simplex 500 x 1100: root', int(root.pivots[0]), 'pivots', root.stats['total_ms'], 'ms')
    lp.close()


def test_too_many_rows_is_refused(blp_lib):
    d = grumpy_random_mip(40, 20, density=0.2, rand_seed=2)
    big = np.vstack([d.A.toarray()] * 420)          # 8400 rows
    lp = engine.BatchLP(big, np.tile(d.b, 420), d.c)
    assert not lp.simplex_capable
    with pytest.raises(engine.BlpError):
        lp.simplex_batch(d.l[None], d.u[None])
    lp.close()


def test_random_lps_bit_exact_and_against_highs(blp_lib):
    """400 random small LPs (infeasible, unbounded through infinite bounds, degenerate, fixed variables):
    the device's status, basis, pivot count and vertex equal the numpy restatement's exactly; status and
    value equal HiGHS's."""
    from test_oracle_pins import _random_lp
    rng = np.random.default_rng(2024)
    seen = {0: 0, 1: 0, 2: 0}
    for t in range(400):
        A, b, c, l, u = _random_lp(rng)
        if not A.any():
            continue
        lp = engine.BatchLP(A, b, c)
        res = lp.simplex_batch(l[None], np.minimum(u, 1e300)[None])
        ref = dual_simplex(A, b, c, l, u)
        same(res, 0, ref, t)
        seen[ref.status] = seen.get(ref.status, 0) + 1
        h = HighsLP(A, c, b, np.full(len(b), HIGHS_INF), l, u).solve()
        assert ref.status == h.status or (h.status in (2, -1) and ref.status in (1, 2)), t
        if ref.status == 0:
            assert abs(res.objective[0] - h.objective) <= 1e-7 * max(1, abs(h.objective)), t
        lp.close()
    assert min(seen.values()) > 20
