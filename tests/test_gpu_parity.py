"""GPU parity tests: the CUDA bound step (through the C ABI) against the exact CPU oracle.

Bar: LP objective within 1e-6 relative of the simplex optimum (north_star), identical
feasible / infeasible / unbounded status; batched SpMV within 1e-12 relative of scipy.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.highs_lp import HIGHS_INF, HighsLP, solve_node_lps
from simple_mip_solver_b200.instances import grumpy_random_mip, numpy_random_mip

pytestmark = pytest.mark.gpu

REL = 1e-6


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def _engine(blp_lib):
    from simple_mip_solver_b200 import engine
    return engine


@pytest.mark.parametrize('B', [1, 3, 8, 32, 45, 64, 130, 300])
@pytest.mark.parametrize('shape', [(7, 5, 0.5), (300, 500, 0.1), (2000, 5000, 0.004)])
def test_spmv_matches_scipy(blp_lib, B, shape):
    import torch
    eng = _engine(blp_lib)
    m, n, dens = shape
    rng = np.random.default_rng(5)
    A = sp.random(m, n, density=dens, random_state=rng, format='csr')
    lp = eng.BatchLP(A, np.zeros(m), np.zeros(n))
    ld = eng.leading_dim(B)
    X = rng.standard_normal((n, ld))
    Y = lp.spmv_device(torch.from_numpy(X).cuda(), transpose=False, B=B).cpu().numpy()
    ref = A @ X
    assert np.allclose(Y[:, :B], ref[:, :B], rtol=1e-12, atol=1e-12)
    Z = rng.standard_normal((m, ld))
    G = lp.spmv_device(torch.from_numpy(Z).cuda(), transpose=True, B=B).cpu().numpy()
    assert np.allclose(G[:, :B], (A.T @ Z)[:, :B], rtol=1e-12, atol=1e-12)
    lp.close()


def _small_branch():
    # test_simple_mip_solver/example_models.py:101-110 in canonical >= form
    A = sp.csr_matrix(np.array([[-1., 0, -1], [0, -1, 0]]))
    return A, np.array([-1.5, -1.25]), -np.ones(3)


def test_small_branch_root_and_children(blp_lib):
    eng = _engine(blp_lib)
    A, b, c = _small_branch()
    lp = eng.BatchLP(A, b, c)
    lb = np.array([[0, 0, 0], [0, 0, 2], [0, 0, 0], [1, 0, 1], [0, 2, 0]], float)
    ub = np.array([[10, 10, 10], [10, 10, 10], [10, 10, 1], [10, 10, 10], [10, 10, 10]], float)
    r = lp.solve_batch(lb, ub, integer_indices=[0, 1, 2])
    assert list(r.status) == [0, 1, 0, 1, 1]
    # root LP: test_base_node.py:406-416 pins objective == -2.75
    assert _rel(r.objective[0], -2.75) <= REL
    assert _rel(r.lower_bound[0], -2.75) <= REL
    assert np.isinf(r.objective[1]) and r.objective[1] > 0
    assert _rel(r.objective[2], -2.75) <= REL
    assert abs(r.x[0, 1] - 1.25) <= 1e-6 and abs(r.x[0, 0] + r.x[0, 2] - 1.5) <= 1e-6
    assert r.frac_idx[0] in (0, 1, 2) and r.frac_idx[1] == -1
    lp.close()


def test_no_branch_and_unbounded(blp_lib):
    eng = _engine(blp_lib)
    # no_branch (example_models.py:90-98): min -x0-x1 s.t. x0+x2<=1, x1<=1 ... integral root
    A = sp.csr_matrix(np.array([[-1., 0, -1], [0, -1, 0]]))
    lp = eng.BatchLP(A, np.array([-1., -1.]), np.array([-1., -1., 0.]))
    r = lp.solve_batch(np.zeros((1, 3)), np.full((1, 3), 10.), integer_indices=[0, 1, 2])
    assert r.status[0] == 0 and _rel(r.objective[0], -2.0) <= REL
    lp.close()
    # unbounded: min -x0 s.t. x0 - x1 >= 0, x >= 0 without upper bounds
    lp = eng.BatchLP(sp.csr_matrix(np.array([[1., -1.]])), np.array([0.]), np.array([-1., 0.]))
    r = lp.solve_batch(np.zeros((1, 2)), np.full((1, 2), np.inf), opts=eng.default_opts(max_iters=20000))
    assert r.status[0] == 2
    lp.close()


@pytest.mark.parametrize('seed', [2, 3, 4, 5])
@pytest.mark.parametrize('shape', [(2, 2), (4, 2), (2, 4), (4, 4), (20, 10)])
def test_random_small_roots(blp_lib, seed, shape):
    eng = _engine(blp_lib)
    nv, nc = shape
    d = grumpy_random_mip(nv, nc, density=0.8, maxObjCoeff=100, maxConsCoeff=100, rand_seed=seed)
    ref = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(d.l[None], d.u[None], integer_indices=d.integer_indices)
    assert r.status[0] == ref.status == 0
    assert _rel(r.objective[0], ref.objective) <= REL
    assert _rel(r.lower_bound[0], ref.objective) <= REL
    lp.close()


def _children(d, x_root, k):
    ints = np.asarray(d.integer_indices)
    frac = np.minimum(x_root[ints] - np.floor(x_root[ints]), np.ceil(x_root[ints]) - x_root[ints])
    cand = ints[np.argsort(-frac, kind='stable')][:k]
    deltas, lbs, ubs = [], [], []
    for j in cand:
        for direction in ('left', 'right'):
            l, u = d.l.copy(), d.u.copy()
            if direction == 'left':
                u[j] = np.floor(x_root[j])
            else:
                l[j] = np.ceil(x_root[j])
            deltas.append([(int(j), l[j], u[j])])
            lbs.append(l)
            ubs.append(u)
    return deltas, np.array(lbs), np.array(ubs)


def test_strong_branch_children_c3(blp_lib):
    """Config 3: 500 vars x 300 rows, density 0.1, 64 candidates x 2 children in one batch."""
    eng = _engine(blp_lib)
    d = grumpy_random_mip(500, 300, density=0.1, maxObjCoeff=10, maxConsCoeff=10, tightness=2, rand_seed=2)
    root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    deltas, lbs, ubs = _children(d, root.x, 64)
    refs = solve_node_lps(d.A, d.b, d.c, lbs, ubs, root_l=d.l, root_u=d.u)
    lp = eng.BatchLP(d.A, d.b, d.c)
    rr = lp.solve_batch(d.l[None], d.u[None], integer_indices=d.integer_indices)
    assert rr.status[0] == 0 and _rel(rr.objective[0], root.objective) <= REL
    r = lp.solve_children(d.l, d.u, deltas, x0=rr.x[0], y0=rr.y[0], integer_indices=d.integer_indices)
    r2 = lp.solve_batch(lbs, ubs, integer_indices=d.integer_indices)
    for k, ref in enumerate(refs):
        assert r.status[k] == ref.status, (k, r.status[k], ref.status)
        assert r2.status[k] == ref.status
        if ref.status == 0:
            assert _rel(r.objective[k], ref.objective) <= REL, (k, r.objective[k], ref.objective)
            assert _rel(r2.objective[k], ref.objective) <= REL, (k, r2.objective[k], ref.objective)
            assert r.lower_bound[k] <= ref.objective + REL * max(1, abs(ref.objective))
    lp.close()


def test_cut_rows_and_masks(blp_lib):
    eng = _engine(blp_lib)
    d = grumpy_random_mip(30, 15, density=0.4, rand_seed=7)
    lp = eng.BatchLP(d.A, d.b, d.c)
    root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    # two valid-looking cuts: sum of x over a subset <= floor-ish value, written as >= rows
    cut1 = np.zeros(d.n); cut1[:10] = -1.0
    cut2 = np.zeros(d.n); cut2[10:25] = -1.0
    rhs = np.array([-np.floor(root.x[:10].sum() - 0.5), -np.floor(root.x[10:25].sum() - 0.5)])
    first = lp.append_rows(np.vstack([cut1, cut2]), rhs)
    assert first == d.m and lp.m == d.m + 2
    masks = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.uint8)
    r = lp.solve_batch(np.tile(d.l, (4, 1)), np.tile(d.u, (4, 1)), row_mask=masks)
    for k in range(4):
        h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u)
        if masks[k, 0]:
            h.add_row(cut1, rhs[0])
        if masks[k, 1]:
            h.add_row(cut2, rhs[1])
        ref = h.solve()
        assert r.status[k] == ref.status == 0
        assert _rel(r.objective[k], ref.objective) <= REL, (k, r.objective[k], ref.objective)
        if not masks[k, 0]:
            assert r.y[k, d.m] == 0.0
    lp.truncate_rows(d.m)
    assert lp.m == d.m
    r = lp.solve_batch(d.l[None], d.u[None])
    assert _rel(r.objective[0], root.objective) <= REL
    lp.close()


def test_sparse_dive_nodes_c4_shape(blp_lib):
    """A reduced C4-like instance (2000 x 1000, ~20 nnz/row): 48 dive nodes vs the oracle."""
    from simple_mip_solver_b200.instances import random_dive_bounds
    eng = _engine(blp_lib)
    d = numpy_random_mip(2000, 1000, density=0.01, seed=2)
    root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    lbs, ubs, _ = random_dive_bounds(d, root.x, 48, 8, seed=1)
    refs = solve_node_lps(d.A, d.b, d.c, lbs, ubs, root_l=d.l, root_u=d.u)
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, x0=np.tile(root.x, (48, 1)), y0=np.tile(np.maximum(root.row_dual, 0), (48, 1)))
    for k, ref in enumerate(refs):
        assert r.status[k] == ref.status, (k, r.status[k], ref.status)
        if ref.status == 0:
            assert _rel(r.objective[k], ref.objective) <= REL, (k, r.objective[k], ref.objective)
    lp.close()


def test_device_most_fractional_index_matches_host_rule(blp_lib):
    """frac_idx (index work, must be exact): the device's most fractional integer column equals
    BaseNode._most_fractional_index (base_node.py:544-562: distance > 1e-4, first wins ties)
    evaluated on the x the call returned — with a ragged integer set."""
    eng = _engine(blp_lib)
    d = numpy_random_mip(800, 400, density=0.02, seed=9)
    rng = np.random.default_rng(1)
    ints = sorted(rng.choice(d.n, size=333, replace=False).tolist())
    from simple_mip_solver_b200.instances import frontier_nodes
    lbs, ubs, _ = frontier_nodes(d, np.full(d.n, 0.5), 0, 70, 5, seed=3, p_down=0.9)
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, integer_indices=ints)
    ii = np.asarray(ints)
    for k in range(70):
        if r.status[k] != 0:
            assert r.frac_idx[k] == -1
            continue
        v = r.x[k, ii]
        dist = np.minimum(v - np.floor(v), np.ceil(v) - v)
        j = int(np.argmax(dist))
        want = int(ii[j]) if dist[j] > 1e-4 else -1
        assert r.frac_idx[k] == want, (k, r.frac_idx[k], want)
    none = lp.solve_batch(lbs[:3], ubs[:3])                      # no integer set: all -1
    assert (none.frac_idx == -1).all()
    lp.close()
