"""GPU: infeasible / unbounded / degenerate inputs — statuses must match the exact simplex and must
be reached by certificate (not by running into the iteration limit)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.highs_lp import HIGHS_INF, HighsLP
from simple_mip_solver_b200.instances import grumpy_random_mip

pytestmark = pytest.mark.gpu


def _ref(A, b, c, l, u):
    return HighsLP(sp.csr_matrix(A), c, b, np.full(len(b), HIGHS_INF), l, u).solve()


def test_jointly_infeasible_rows_not_caught_by_the_bound_screen(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    # x0 + x1 >= 3 and x0 + x1 <= 2: each row is satisfiable inside the box, together they are not
    A = np.array([[1., 1, 0], [-1, -1, 0], [0, 0, 1]])
    b = np.array([3., -2, 0])
    lp = eng.BatchLP(sp.csr_matrix(A), b, np.array([1., 2, 3]))
    r = lp.solve_batch(np.zeros((1, 3)), np.full((1, 3), 10.))
    assert r.status[0] == 1 and r.iterations[0] < 20000
    assert np.isinf(r.objective[0])
    lp.close()


@pytest.mark.parametrize('seed', range(6))
def test_random_infeasible_and_feasible_mix(blp_lib, seed):
    """A batch where a contradicting pair of rows is switched on (by row mask) for some nodes."""
    from simple_mip_solver_b200 import engine as eng
    d = grumpy_random_mip(30, 15, density=0.4, rand_seed=seed + 10)
    root = _ref(d.A, d.b, d.c, d.l, d.u)
    assert root.status == 0
    rng = np.random.default_rng(seed)
    a = np.zeros(d.n)
    S = rng.choice(d.n, size=8, replace=False)
    a[S] = rng.integers(1, 5, size=8)
    v = float(a @ root.x)
    # rows: a.x >= v + 3  and  -a.x >= -(v + 1)  -> together infeasible; each alone feasible or not?
    hi = float(a @ d.u)
    cut_hi = min(v + 3.0, hi - 1.0)
    rows = np.vstack([a, -a])
    rhs = np.array([cut_hi, -(cut_hi - 2.0)])
    lp = eng.BatchLP(d.A, d.b, d.c)
    lp.append_rows(rows, rhs)
    masks = np.array([[0, 0], [1, 1], [1, 0], [0, 1], [1, 1]], dtype=np.uint8)
    B = len(masks)
    r = lp.solve_batch(np.tile(d.l, (B, 1)), np.tile(d.u, (B, 1)), row_mask=masks)
    for k in range(B):
        h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u)
        for t in np.flatnonzero(masks[k]):
            h.add_row(rows[t], rhs[t])
        ref = h.solve()
        assert r.status[k] == ref.status, (k, r.status[k], ref.status, r.iterations[k])
        if ref.status == 0:
            assert abs(r.objective[k] - ref.objective) <= 1e-6 * max(1, abs(ref.objective))
        else:
            assert r.iterations[k] < 100000
    lp.close()


def test_unbounded_ray_and_free_columns(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    # the reference's `unbounded` example (example_models.py:154-163) in canonical form
    A = np.array([[-1., 1], [1, -1]])
    lp = eng.BatchLP(sp.csr_matrix(A), np.array([-0.5, -0.5]), np.array([-1., -1]))
    r = lp.solve_batch(np.zeros((1, 2)), np.full((1, 2), np.inf))
    assert r.status[0] == 2
    lp.close()
    # free variables and an empty row / empty column are accepted; bounded optimum
    A = np.array([[1., 1, 0, 0], [0, 0, 0, 0], [1, -1, 0, 0]])
    b = np.array([1., -5, -2])
    c = np.array([1., 1, 0, 2])
    l = np.array([-np.inf, -np.inf, 0, 1])
    u = np.array([np.inf, np.inf, 4, 3])
    ref = _ref(A, b, c, np.where(np.isinf(l), -HIGHS_INF, l), np.where(np.isinf(u), HIGHS_INF, u))
    lp = eng.BatchLP(sp.csr_matrix(A), b, c)
    r = lp.solve_batch(l[None], u[None])
    assert r.status[0] == ref.status == 0 and abs(r.objective[0] - ref.objective) <= 1e-6 * max(1, abs(ref.objective))
    lp.close()


def test_argument_errors_are_reported_not_crashed(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    with pytest.raises(eng.BlpError, match='row_lb'):
        eng.BatchLP(sp.eye(2, format='csr'), np.array([0., np.inf]), np.ones(2))
    lp = eng.BatchLP(sp.eye(2, format='csr'), np.zeros(2), np.ones(2))
    with pytest.raises(ValueError):
        lp.solve_batch(np.zeros((1, 3)), np.ones((1, 3)))
    with pytest.raises(eng.BlpError):
        lp.truncate_rows(1)
    with pytest.raises(TypeError):
        eng.default_opts(nonsense=1)
    with pytest.raises(eng.BlpError, match='eps_rel'):
        lp.solve_batch(np.zeros((1, 2)), np.ones((1, 2)), opts=eng.default_opts(eps_rel=0.0))
    lp.close()
