"""The CPU laboratory (tests/tools/cpu_pdhg_lab.py) must stay the same algorithm as the batched numpy
model of the device iteration (oracle/pdhg_numpy.py) — otherwise what it screens says nothing about
the kernels. Same instance, same constants: identical iteration counts and objectives."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools'))
import cpu_pdhg_lab as lab                                       # noqa: E402
from oracle.pdhg_numpy import BatchPDHG                          # noqa: E402
from simple_mip_solver_b200.instances import frontier_nodes, numpy_random_mip   # noqa: E402


@pytest.mark.parametrize('balance,dead', [(0.0, 0.0), (0.3, 0.0), (0.3, 0.25)])
def test_lab_matches_the_batched_model(balance, dead):
    d = numpy_random_mip(300, 120, 0.03, seed=5)
    P = BatchPDHG(d.A, d.b, d.c)
    root = P.solve(d.l[:, None], d.u[:, None], eps=1e-8, theta=0.05)
    assert root['status'][0] == 0
    x0, y0 = root['x'][:, 0], root['y'][:, 0]
    lbs, ubs, _ = frontier_nodes(d, x0, 0, 4, 6, seed=3)
    ref = P.solve(lbs.T, ubs.T, eps=1e-7, theta=0.05, balance=balance, balance_dead=dead,
                  x0=np.tile(x0[:, None], (1, 4)), y0=np.tile(y0[:, None], (1, 4)))
    for k in range(4):
        # long_after beyond the run: the batched model evaluates every 64 iterations throughout
        r = lab.solve1(P, lbs[k], ubs[k], x0=x0, y0=y0, eps=1e-7, theta=0.05, balance=balance,
                       bal_dead=dead, long_after=10 ** 9)
        assert ref['status'][k] == 0
        assert r['iters'] == ref['iters'][k], (k, r['iters'], ref['iters'][k])
        assert abs(r['obj'] - ref['obj'][k]) <= 1e-9 * max(1.0, abs(ref['obj'][k]))


def test_variant_table():
    class P:
        omega0 = 2.0
    assert lab.variant_kwargs('base', P) == {}
    assert lab.variant_kwargs('nobal', P) == dict(balance=0.0, bal_dead=0.0)
    assert lab.variant_kwargs('dz0.5_0.3', P) == dict(bal_dead=0.5, balance=0.3)
    assert lab.variant_kwargs('om0.5', P)['omega_init'] == 1.0
    with pytest.raises(SystemExit):
        lab.variant_kwargs('nonsense', P)


def test_freeze_lab_without_freezing_is_the_lab_and_freezing_keeps_the_answers():
    """tests/tools/cpu_freeze_lab.py screens the freezing rule of the wide-batch kernels (k_freeze_cols / k_freeze_rows):
    with the rule off it is the single-node lab run in lock step; with it on, a good share of the coordinate updates
    is skipped, every node ends on the same objective within the solve tolerance and about as fast."""
    import cpu_freeze_lab as flab
    d = numpy_random_mip(300, 120, 0.03, seed=5)
    P = BatchPDHG(d.A, d.b, d.c)
    root = P.solve(d.l[:, None], d.u[:, None], eps=1e-8, theta=0.05)
    x0, y0 = root['x'][:, 0], root['y'][:, 0]
    lbs, ubs, _ = frontier_nodes(d, x0, 0, 4, 6, seed=3)
    off = flab.solve_tile(P, lbs, ubs, x0, y0)
    assert off['done'].all() and off['skipped'] == 0.0
    for k in range(4):
        r = lab.solve1(P, lbs[k], ubs[k], x0=x0, y0=y0)
        assert r['iters'] == off['iters'][k]
        assert abs(r['obj'] - off['obj'][k]) <= 1e-9 * max(1.0, abs(r['obj']))
    on = flab.solve_tile(P, lbs, ubs, x0, y0, m_lo=0.03, m_hi=0.1)
    assert on['done'].all() and on['skipped'] > 0.2
    assert np.allclose(on['obj'], off['obj'], rtol=5e-7, atol=0)
    assert on['iters'].mean() <= 1.15 * off['iters'].mean()
    # the step size of the submatrix that is not frozen (larger than 1 / ||A||) converges to the same answers
    sub = flab.solve_tile(P, lbs, ubs, x0, y0, m_lo=0.03, m_hi=0.1, sub_step=0.9)
    assert sub['done'].all() and np.allclose(sub['obj'], off['obj'], rtol=5e-7, atol=0)
