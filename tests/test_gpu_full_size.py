"""GPU: BASELINE.json's full shapes (config 4: 10 000 x 5 000, config 5: 50 000 x 20 000).

At these sizes the simplex oracle needs minutes per LP, so parity is checked two ways:
  * against the committed golden root solutions ``bench_data/c{4,5}_root.npz`` (HiGHS dual simplex
    run once offline by tests/tools/make_bench_fixture.py): the root LP solved cold on the GPU must
    reproduce the golden optimum within 1e-6 relative;
  * through properties that do not depend on the size: every child of the root is at least as
    expensive as the root, weak duality and the reported duality gap, primal feasibility of the
    returned x in the unscaled problem, sign of the row duals, the returned objective and Lagrangian
    bound recomputed on the host from the returned x and y, and the most-fractional index rule
    (index work: exact).
The C5 case runs through continuous batching (more nodes than resident slots).
"""
import os

import numpy as np
import pytest

from simple_mip_solver_b200.instances import frontier_nodes, numpy_random_mip

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-6           # north_star: per-node LP objective within 1e-6 relative
EPS = 1e-7           # blp_opts.eps_rel default


def _golden(name):
    z = np.load(os.path.join(ROOT, 'bench_data', f'{name}_root.npz'))
    return z['x'], np.maximum(z['y'], 0.0), float(z['objective'])


def _check_properties(d, lbs, ubs, r, root_obj, ints):
    B = lbs.shape[0]
    assert (r.status == 0).all(), np.unique(r.status, return_counts=True)
    bnorm, scale = np.linalg.norm(d.b), 1.0 + np.abs(r.objective) + np.abs(r.lower_bound)
    # children only tighten bounds: never cheaper than the root
    assert (r.objective >= root_obj - REL * abs(root_obj)).all()
    # weak duality and the gap criterion of the solve
    assert (r.lower_bound <= r.objective + 1.01 * EPS * scale).all()
    assert (np.abs(r.objective - r.lower_bound) <= 1.01 * EPS * scale).all()
    for k in range(B):
        x, y = r.x[k], r.y[k]
        assert (x >= lbs[k] - 1e-12).all() and (x <= ubs[k] + 1e-12).all()          # the projection is exact
        viol = np.maximum(d.b - d.A @ x, 0.0)
        assert np.linalg.norm(viol) <= 1.01 * EPS * (1.0 + bnorm), (k, np.linalg.norm(viol))
        assert (y >= 0.0).all()
        assert abs(float(d.c @ x) - r.objective[k]) <= 1e-10 * scale[k]
        rc = d.c - d.A.T @ y
        lag = float(d.b @ y) + float(np.minimum(rc * lbs[k], rc * ubs[k]).sum())
        assert abs(lag - r.lower_bound[k]) <= 1e-9 * scale[k], (k, lag, r.lower_bound[k])
        xi = x[ints]
        dist = np.minimum(xi - np.floor(xi), np.ceil(xi) - xi)
        best = int(np.argmax(dist))                                                   # first maximum wins ties
        want = ints[best] if dist[best] > 1e-4 else -1
        assert r.frac_idx[k] == want, (k, r.frac_idx[k], want)


def test_config5_full_size_root_golden_and_properties(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    d = numpy_random_mip(50000, 20000, density=2e-4, seed=2)
    gx, gy, gobj = _golden('c5')
    B = 72
    lbs, ubs, _ = frontier_nodes(d, gx, 0, B, 32, seed=0)
    lbs[0], ubs[0] = d.l, d.u                        # node 0: the root LP itself, solved cold
    x0, y0 = np.tile(gx, (B, 1)), np.tile(gy, (B, 1))
    x0[0], y0[0] = 0.0, 0.0
    ints = np.arange(0, d.n, 3)                       # a ragged integer set
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, integer_indices=ints.tolist(),
                       opts=eng.default_opts(max_active=64))
    lp.close()
    assert r.stats['refills'] == B - 64
    assert abs(r.objective[0] - gobj) <= REL * abs(gobj), (r.objective[0], gobj)
    assert r.iterations[0] > r.iterations[1:].max() / 4       # the cold root is no shortcut
    _check_properties(d, lbs, ubs, r, gobj, ints)


def test_config4_full_size_root_golden_and_properties(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    d = numpy_random_mip(10000, 5000, density=2e-3, seed=2)
    gx, gy, gobj = _golden('c4')
    B = 48
    lbs, ubs, _ = frontier_nodes(d, gx, 0, B, 16, seed=0)
    lbs[0], ubs[0] = d.l, d.u
    x0, y0 = np.tile(gx, (B, 1)), np.tile(gy, (B, 1))
    x0[0], y0[0] = 0.0, 0.0
    ints = np.arange(d.n)
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, x0=x0, y0=y0, integer_indices=ints.tolist())
    lp.close()
    assert abs(r.objective[0] - gobj) <= REL * abs(gobj), (r.objective[0], gobj)
    _check_properties(d, lbs, ubs, r, gobj, ints)


@pytest.mark.parametrize('name,shape', [('c4', (10000, 5000, 2e-3, 16)), ('c5', (50000, 20000, 2e-4, 32))])
def test_children_against_committed_highs_goldens(blp_lib, name, shape):
    """64 frontier children of the C4 / C5 root against HiGHS answers made offline
    (tests/tools/make_child_goldens.py -> bench_data/<name>_children.npz): equal status and objective
    within 1e-6 relative, through the children form of the C ABI (bounds as deltas against the root,
    the root's primal/dual pair as warm start, continuous batching with 48 resident slots)."""
    from simple_mip_solver_b200 import engine as eng
    from simple_mip_solver_b200.instances import GOLD_COUNT, GOLD_FIRST
    n, m, dens, depth = shape
    d = numpy_random_mip(n, m, density=dens, seed=2)
    gx, gy, gobj = _golden(name)
    z = np.load(os.path.join(ROOT, 'bench_data', f'{name}_children.npz'))
    assert z['node_ids'][0] == GOLD_FIRST and len(z['node_ids']) == GOLD_COUNT
    _, _, deltas = frontier_nodes(d, gx, GOLD_FIRST, GOLD_COUNT, depth, seed=0, dense=False)
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_children(d.l, d.u, deltas, x0=gx, y0=gy, opts=eng.default_opts(max_active=48), want_x=False,
                          want_y=False)
    lp.close()
    assert np.array_equal(r.status, z['status']), (r.status, z['status'])
    ok = z['status'] == 0
    err = np.abs(r.objective[ok] - z['objective'][ok]) / np.maximum(1.0, np.abs(z['objective'][ok]))
    print(f'{name}: 64 children, max rel. objective error vs HiGHS {err.max():.2e}, '
          f'mean {int(r.iterations.mean())} PDHG iterations')
    assert err.max() <= REL
