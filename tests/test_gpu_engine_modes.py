"""GPU: execution modes of blp_solve_batch must not change the answer.

Every node's iteration sequence is independent of how the batch is laid out or launched, so the
same LPs solved with / without compaction of finished nodes, with / without CUDA graphs, with the
one- or two-nodes-per-lane step kernels, alone or inside a larger batch, must give the same
status and iteration count and objectives equal to round-off.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from simple_mip_solver_b200.instances import frontier_nodes, numpy_random_mip

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _instance():
    d = numpy_random_mip(1500, 700, density=0.01, seed=4)
    x = np.full(d.n, 0.5)
    lbs, ubs, _ = frontier_nodes(d, x, 0, 150, 6, seed=2, p_down=0.9)
    return d, lbs, ubs


def _solve(eng, d, lbs, ubs, **kw):
    lp = eng.BatchLP(d.A, d.b, d.c)
    r = lp.solve_batch(lbs, ubs, integer_indices=d.integer_indices, opts=eng.default_opts(**kw))
    lp.close()
    return r


def _same(a, b, tol=1e-11):
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.iterations, b.iterations)
    ok = a.status == 0
    assert np.allclose(a.objective[ok], b.objective[ok], rtol=tol, atol=tol)
    assert np.allclose(a.lower_bound[ok], b.lower_bound[ok], rtol=tol, atol=tol)
    assert np.allclose(a.x[ok], b.x[ok], rtol=1e-9, atol=1e-9)
    assert np.array_equal(a.frac_idx[ok], b.frac_idx[ok])


def test_compaction_and_graphs_do_not_change_results(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs = _instance()
    base = _solve(eng, d, lbs, ubs)
    assert base.stats['compactions'] > 0 and (base.status == 0).sum() > 100
    assert len(np.unique(base.iterations)) > 5            # nodes really finish at different times
    _same(base, _solve(eng, d, lbs, ubs, compact=0))
    _same(base, _solve(eng, d, lbs, ubs, use_graph=0))
    _same(base, _solve(eng, d, lbs, ubs, profile=1))


def test_batch_composition_does_not_change_a_node(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs = _instance()
    full = _solve(eng, d, lbs, ubs)
    part = _solve(eng, d, lbs[40:75], ubs[40:75])        # 35 nodes: one-node-per-lane kernels
    assert np.array_equal(full.status[40:75], part.status)
    assert np.array_equal(full.iterations[40:75], part.iterations)
    ok = part.status == 0
    assert np.allclose(full.objective[40:75][ok], part.objective[ok], rtol=1e-11, atol=1e-11)
    one = _solve(eng, d, lbs[7:8], ubs[7:8])
    assert one.status[0] == full.status[7] and one.iterations[0] == full.iterations[7]


def test_one_and_two_nodes_per_lane_kernels_agree(blp_lib):
    """BLP_V2=0 forces the one-node-per-lane step kernels for wide batches (env read per call)."""
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
        "from test_gpu_engine_modes import _instance, _solve\n"
        "from simple_mip_solver_b200 import engine as eng\n"
        "d, lbs, ubs = _instance(); r = _solve(eng, d, lbs, ubs)\n"
        "print(json.dumps(dict(status=r.status.tolist(), iters=r.iterations.tolist(), obj=np.where(r.status == 0, r.objective, 0).tolist())))\n"
    ) % (ROOT, ROOT)
    outs = []
    for v2 in ('1', '0'):
        env = dict(os.environ, BLP_V2=v2)
        res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        import json
        outs.append(json.loads(res.stdout.strip().splitlines()[-1]))
    assert outs[0]['status'] == outs[1]['status'] and outs[0]['iters'] == outs[1]['iters']
    assert np.allclose(outs[0]['obj'], outs[1]['obj'], rtol=1e-11, atol=1e-11)


def test_iteration_budget_returns_lagrangian_bound(blp_lib):
    """status 3 (budget, the analogue of lp.maxNumIteration): lower_bound is a valid bound."""
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs = _instance()
    full = _solve(eng, d, lbs[:64], ubs[:64])
    short = _solve(eng, d, lbs[:64], ubs[:64], max_iters=256)
    lim = short.status == 3
    assert lim.sum() > 10
    both = lim & (full.status == 0)
    assert (short.lower_bound[both] <= full.objective[both] + 1e-6 * np.abs(full.objective[both])).all()
    assert (short.iterations[lim] == 256).all()
