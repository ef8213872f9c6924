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


def test_one_and_two_nodes_per_lane_kernels_agree(blp_lib, tmp_path):
    """BLP_V2=0 forces the one-node-per-lane step kernels for wide batches. The production library
    ignores BLP_* variables; a tuning build (-DBLP_TUNING) of the same sources reads them."""
    from simple_mip_solver_b200 import _build
    tuning_lib = _build.build_extension(tuning=True, out=tmp_path / 'libblp_tuning.so')
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
        "from test_gpu_engine_modes import _instance, _solve\n"
        "from simple_mip_solver_b200 import engine as eng\n"
        "d, lbs, ubs = _instance(); r = _solve(eng, d, lbs, ubs)\n"
        "print(json.dumps(dict(status=r.status.tolist(), iters=r.iterations.tolist(), obj=np.where(r.status == 0, r.objective, 0).tolist())))\n"
    ) % (ROOT, ROOT)
    outs = []
    for v2 in ('1', '0'):
        env = dict(os.environ, BLP_V2=v2, BLP_LIB=str(tuning_lib))
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


def test_continuous_batching_matches_resident_batch(blp_lib):
    """blp_opts.max_active < B: pending nodes take over the slots of finished ones. Every node
    must end with the same status, an objective equal within the solve tolerance and its own
    iteration count (counted from the moment it was loaded)."""
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs = _instance()
    B = lbs.shape[0]
    lbs, ubs = lbs.copy(), ubs.copy()
    lbs[97, 5], ubs[97, 5] = 3.0, 2.0          # empty boxes among the pending nodes: status 1
    lbs[149, 0], ubs[149, 0] = 1.0, 0.0        # without an iteration, also as the very last node
    base = _solve(eng, d, lbs, ubs)
    assert base.status[97] == 1 and base.status[149] == 1
    for slots in (64, 40, 7):
        r = _solve(eng, d, lbs, ubs, max_active=slots)
        assert r.stats['refills'] == B - slots
        assert r.stats['compactions'] > 0                     # the tail still compacts
        assert np.array_equal(r.status, base.status)
        ok = base.status == 0
        assert np.allclose(r.objective[ok], base.objective[ok], rtol=5e-7, atol=1e-9)
        assert np.allclose(r.lower_bound[ok], base.lower_bound[ok], rtol=5e-7, atol=1e-9)
        assert (r.iterations[ok] > 0).all() and r.iterations[97] == 0 and r.iterations[149] == 0
        # per-node iteration counts stay in the range of the resident solve (own clock, not the batch's)
        assert r.iterations[ok].max() <= 2 * base.iterations[ok].max()
        assert np.isinf(r.objective[97]) and np.isinf(r.objective[149])
        # x is the point the objective was computed at
        assert np.allclose((r.x[ok] * d.c).sum(1), r.objective[ok], rtol=1e-9, atol=1e-9)
    # max_active >= B is the resident batch itself
    _same(base, _solve(eng, d, lbs, ubs, max_active=B))
    _same(base, _solve(eng, d, lbs, ubs, max_active=10 * B))


def test_continuous_batching_iteration_budget_is_per_node(blp_lib):
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs = _instance()
    r = _solve(eng, d, lbs[:96], ubs[:96], max_active=32, max_iters=256)
    lim = r.status == 3
    assert lim.sum() > 10
    assert (r.iterations[lim] >= 256).all() and (r.iterations[lim] < 256 + 256).all()
    full = _solve(eng, d, lbs[:96], ubs[:96])
    both = lim & (full.status == 0)
    assert (r.lower_bound[both] <= full.objective[both] + 1e-6 * np.abs(full.objective[both])).all()


def test_continuous_batching_carries_row_masks_and_warm_starts(blp_lib):
    """Refilled slots read their own row mask, x0 and y0 columns from the caller's arrays."""
    from oracle.highs_lp import HIGHS_INF, HighsLP
    from simple_mip_solver_b200 import engine as eng
    d = numpy_random_mip(600, 300, density=0.02, seed=7)
    root = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), d.l, d.u).solve()
    B = 80
    lbs, ubs, _ = frontier_nodes(d, root.x, 0, B, 6, seed=1)
    rng = np.random.default_rng(3)
    rows, rhs = [], []
    for k in range(6):
        S = rng.choice(d.n, size=60, replace=False)
        row = np.zeros(d.n)
        row[S] = -1.0
        rows.append(row)
        rhs.append(-np.floor(root.x[S].sum()))
    lp = eng.BatchLP(d.A, d.b, d.c)
    lp.append_rows(np.array(rows), np.array(rhs))
    masks = (rng.random((B, 6)) < 0.5).astype(np.uint8)
    x0 = np.tile(root.x, (B, 1))
    y0 = np.hstack([np.tile(np.maximum(root.row_dual, 0), (B, 1)), np.zeros((B, 6))])
    res = lp.solve_batch(lbs, ubs, row_mask=masks, x0=x0, y0=y0, opts=eng.default_opts(max_active=24))
    lp.close()
    assert res.stats['refills'] == B - 24
    for k in range(0, B, 3):
        h = HighsLP(d.A, d.c, d.b, np.full(d.m, HIGHS_INF), lbs[k], ubs[k])
        for t in np.flatnonzero(masks[k]):
            h.add_row(rows[t], rhs[t])
        ref = h.solve()
        assert res.status[k] == ref.status, k
        if ref.status == 0:
            assert abs(res.objective[k] - ref.objective) <= 1e-6 * max(1.0, abs(ref.objective)), k
    ok = res.status == 0
    assert (res.y[ok][:, d.m:][masks[ok] == 0] == 0).all()


def test_objective_cutoff_retires_nodes_early_with_a_valid_bound(blp_lib):
    """blp_opts.obj_cutoff: a node whose dual bound reaches the limit stops with status 5; its
    lower_bound is >= the cutoff and <= its true LP value; nodes below the limit are not affected."""
    from simple_mip_solver_b200 import engine
    d = numpy_random_mip(2000, 1000, density=5e-3, seed=3)
    lp = engine.BatchLP(d.A, d.b, d.c)
    root = lp.solve_batch(d.l[None], d.u[None])
    _, _, deltas = frontier_nodes(d, root.x[0], 0, 64, 8, seed=1, dense=False)
    full = lp.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0])
    assert (full.status == 0).all()
    cutoff = float(np.median(full.objective))
    cut = lp.solve_children(d.l, d.u, deltas, x0=root.x[0], y0=root.y[0], opts=engine.default_opts(obj_cutoff=cutoff))
    above = full.objective > cutoff + 1e-6 * abs(cutoff)
    below = full.objective < cutoff - 1e-6 * abs(cutoff)
    assert above.sum() >= 16 and below.sum() >= 16
    assert (cut.status[above] == 5).all() and (cut.status[below] == 0).all()
    assert (cut.lower_bound[above] >= cutoff).all()
    assert (cut.lower_bound[above] <= full.objective[above] + 1e-7 * np.abs(full.objective[above])).all()
    assert np.allclose(cut.objective[below], full.objective[below], rtol=1e-7)
    assert cut.iterations[above].mean() < 0.8 * full.iterations[above].mean()
    print(f'cutoff at the median: {int(above.sum())} nodes retired after {int(cut.iterations[above].mean())} '
          f'instead of {int(full.iterations[above].mean())} iterations')
    lp.close()


def test_branch_and_bound_with_lp_cutoff_builds_the_same_tree(blp_lib, monkeypatch):
    import json, os
    from simple_mip_solver_b200 import BaseNode, BranchAndBound, CyLPArray, MILPInstance
    from simple_mip_solver_b200.compat.cylp_like import SharedLP
    monkeypatch.setattr(SharedLP, 'default_method', 'pdhg')
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'example_models.json')))
    rec = gold['random']

    def run(cut):
        m = MILPInstance(A=np.array(rec['A']), b=CyLPArray(rec['b']), c=CyLPArray(rec['c']), l=CyLPArray(rec['l']),
                         u=CyLPArray(rec['u']), sense=['Min', '>='], integerIndices=list(rec['integer_indices']),
                         numVars=len(rec['c']))
        bb = BranchAndBound(m, BaseNode, frontier_batch=8, gomory_cuts=False, lp_cutoff=cut)
        bb.solve()
        tree = {i: (bb.tree.get_parent(i), v.attr['node']._b_idx, v.attr['node']._b_dir) for i, v in bb.tree.nodes.items()}
        cut_nodes = sum(1 for v in bb.tree.nodes.values() if getattr(v.attr['node'], 'lp_cut_off', False))
        m.lp._shared.close()
        return bb.status, bb.objective_value, tree, cut_nodes
    a, b = run(False), run(True)
    assert a[0] == b[0] == 'optimal' and abs(a[1] - b[1]) <= 1e-6 * abs(a[1]) and a[2] == b[2]
    print('nodes stopped by the objective limit:', b[3], 'of', len(b[2]))


def _c4_frontier(B):
    z = np.load(os.path.join(ROOT, 'bench_data', 'c4_root.npz'))
    d = numpy_random_mip(int(z['n']), int(z['m']), density=float(z['density']), seed=int(z['seed']))
    lbs, ubs, _ = frontier_nodes(d, z['x'], 0, B, 16, seed=0)
    return d, lbs, ubs, z['x'], np.maximum(z['y'], 0.0)


def test_frozen_coordinates_do_not_change_answers(blp_lib):
    """blp_opts.freeze: the wide-batch step kernels skip coordinates that rest for a whole 64-node tile (columns on a
    bound with a safe reduced cost, rows with zero multiplier and safe slack). The evaluation is always the full
    problem's, so statuses are equal and objectives agree within the solve tolerance; the iteration counts stay
    within a few percent (the frozen iteration IS the full iteration while the margins hold), and the statistics
    say how much stream was saved. Resident batch and continuous batching (refills re-derive the flags)."""
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs, x0, y0 = _c4_frontier(160)
    B = lbs.shape[0]
    lbs, ubs = lbs.copy(), ubs.copy()
    lbs[150, 3], ubs[150, 3] = 1.0, 0.0                 # an empty box among the pending nodes
    X0, Y0 = np.tile(x0, (B, 1)), np.tile(y0, (B, 1))
    lp = eng.BatchLP(d.A, d.b, d.c)
    res = {}
    for name, kw in (('off', dict(freeze=0)), ('on', dict(step_safety=0.0)), ('on_refill', dict(max_active=128, step_safety=0.0)),
                     ('off_refill', dict(freeze=0, max_active=128)), ('loose', dict(freeze_margin=0.005, step_safety=0.0)),
                     ('step', {}), ('step_refill', dict(max_active=128))):
        res[name] = lp.solve_batch(lbs, ubs, x0=X0, y0=Y0, integer_indices=d.integer_indices, opts=eng.default_opts(**kw))
    lp.close()
    off = res['off']
    ok = off.status == 0
    assert ok.sum() == B - 1 and off.status[150] == 1
    assert off.stats['skipped_col_updates'] == 0 and off.stats['skipped_row_updates'] == 0
    total_c, total_r = off.stats['node_iterations'] * d.n, off.stats['node_iterations'] * d.m
    for name in ('on', 'on_refill', 'loose'):
        r = res[name]
        assert np.array_equal(r.status, off.status), name
        assert np.allclose(r.objective[ok], off.objective[ok], rtol=2e-7, atol=0), name
        # the most fractional index is read from x, which on these degenerate LPs is not unique to more than ~1e-4
        # between two runs that agree on the objective to 1e-9: nearly always the same variable, never a non-fractional one
        assert (r.frac_idx[ok] == off.frac_idx[ok]).mean() > 0.8, name
        assert ((r.frac_idx[ok] >= 0) == (off.frac_idx[ok] >= 0)).all(), name
        assert abs(r.iterations[ok].mean() / off.iterations[ok].mean() - 1.0) < 0.05, name
        assert r.stats['skipped_col_updates'] > 0.3 * total_c, (name, r.stats['skipped_col_updates'] / total_c)
        assert r.stats['skipped_row_updates'] > 0.1 * total_r, (name, r.stats['skipped_row_updates'] / total_r)
        # x is the point the objective was computed at, and it is feasible to the solve tolerance
        assert np.allclose((r.x[ok] * d.c).sum(1), r.objective[ok], rtol=1e-9, atol=1e-9)
        viol = np.maximum(d.b[None, :] - (d.A @ r.x[ok].T).T, 0.0)
        assert (np.linalg.norm(viol, axis=1) <= 1.01e-7 * (1.0 + np.linalg.norm(d.b))).all(), name
    assert np.array_equal(res['off_refill'].status, off.status)
    assert res['loose'].stats['skipped_col_updates'] >= res['on'].stats['skipped_col_updates']
    # blp_opts.step_safety (default on): the step of a tile follows the norm of the rows and columns that still move
    # (0.47 ||A|| on this matrix), which nearly halves the iterations; same statuses, same objectives
    for name in ('step', 'step_refill'):
        r = res[name]
        assert np.array_equal(r.status, off.status), name
        assert np.allclose(r.objective[ok], off.objective[ok], rtol=2e-7, atol=0), name
        assert r.iterations[ok].mean() < 0.75 * off.iterations[ok].mean(), (name, r.iterations[ok].mean())
        assert r.stats['step_resets'] <= 0.1 * B, (name, r.stats['step_resets'])
        viol = np.maximum(d.b[None, :] - (d.A @ r.x[ok].T).T, 0.0)
        assert (np.linalg.norm(viol, axis=1) <= 1.01e-7 * (1.0 + np.linalg.norm(d.b))).all(), name
    assert off.stats['step_resets'] == 0 and res['on'].stats['step_resets'] == 0


def test_frozen_coordinates_with_cut_rows_masked_per_node(blp_lib):
    """A pool row that is switched off for every node of a tile counts as frozen (its multiplier is held at zero);
    rows that are on for some nodes keep iterating. Same answers as without freezing."""
    from simple_mip_solver_b200 import engine as eng
    d, lbs, ubs, x0, y0 = _c4_frontier(128)
    B = lbs.shape[0]
    rng = np.random.default_rng(5)
    k = 6
    rows = np.zeros((k, d.n))
    for i in range(k):
        idx = rng.choice(d.n, 300, replace=False)
        rows[i, idx] = rng.integers(1, 10, 300)
    rhs = rows @ x0 + np.array([0.5, 1.0, 0.25, 2.0, 0.1, 0.75])      # violated by the root optimum
    lp = eng.BatchLP(d.A, d.b, d.c)
    lp.append_rows(rows, rhs)
    mask = np.zeros((B, k), dtype=np.uint8)
    mask[:64, 0] = 1                      # on for the first tile only
    mask[::2, 1] = 1                      # on for every other node
    mask[:, 2] = 1                        # on everywhere; rows 3..5 off everywhere
    X0, Y0 = np.tile(x0, (B, 1)), np.tile(np.concatenate([y0, np.zeros(k)]), (B, 1))
    a = lp.solve_batch(lbs, ubs, row_mask=mask, x0=X0, y0=Y0, opts=eng.default_opts(freeze=0))
    b = lp.solve_batch(lbs, ubs, row_mask=mask, x0=X0, y0=Y0, opts=eng.default_opts())
    lp.close()
    assert np.array_equal(a.status, b.status) and (a.status == 0).all()
    assert np.allclose(a.objective, b.objective, rtol=2e-7, atol=0)
    assert b.stats['skipped_row_updates'] > 0
    # the cuts bind: nodes that carry row 2 (all) are more expensive than the root
    assert (b.objective > float(np.load(os.path.join(ROOT, 'bench_data', 'c4_root.npz'))['objective'])).all()


def test_wide_batch_with_rows_too_heavy_for_the_shared_memory_slab(blp_lib):
    """Rows of ~50 entries: a CTA's chunk of A (128 rows) does not fit the shared-memory slab, so the dual step reads
    the entries of its rows from global memory — with frozen coordinates, from the tile's folded matrix with the row
    ends — and the slab is at its 44 KB cap next to the kernels' static shared memory (a launch that needs the
    opt-in above 48 KB; the C5 stress variant of SURVEY 8d failed to launch before it was requested)."""
    from simple_mip_solver_b200 import engine as eng
    d = numpy_random_mip(20000, 2000, density=0.0025, seed=9)
    lp = eng.BatchLP(d.A, d.b, d.c)
    root = lp.solve_batch(d.l[None], d.u[None], want_y=True)
    assert root.status[0] == 0
    lbs, ubs, _ = frontier_nodes(d, root.x[0], 0, 64, 12, seed=1)
    X0, Y0 = np.tile(root.x[0], (64, 1)), np.tile(root.y[0], (64, 1))
    off = lp.solve_batch(lbs, ubs, x0=X0, y0=Y0, opts=eng.default_opts(freeze=0))
    on = lp.solve_batch(lbs, ubs, x0=X0, y0=Y0, opts=eng.default_opts())
    lp.close()
    assert np.array_equal(on.status, off.status) and (off.status == 0).sum() >= 60
    ok = off.status == 0
    assert np.allclose(on.objective[ok], off.objective[ok], rtol=2e-7, atol=0)
    assert on.stats['skipped_col_updates'] > 0 and off.stats['skipped_col_updates'] == 0


def test_infeasible_nodes_inside_a_frozen_wide_batch(blp_lib):
    """Nodes of a wide batch that are infeasible only through TWO rows together (sum_S x >= K and sum_S x <= K - 1, pool
    rows switched on per node; each row alone passes the row-activity screen) must get status 1 from the Farkas test
    with freezing and the larger step on, exactly as without, while their neighbours in the same tiles solve to the
    same objectives. The certificate needs multipliers on rows that start frozen at zero: they have to be released."""
    from simple_mip_solver_b200 import engine as eng
    d = numpy_random_mip(20000, 2000, density=0.0025, seed=9)
    lp = eng.BatchLP(d.A, d.b, d.c)
    root = lp.solve_batch(d.l[None], d.u[None], want_y=True)
    B, k = 64, 2
    lbs, ubs, _ = frontier_nodes(d, root.x[0], 0, B, 12, seed=1)
    S = np.arange(0, 400)
    K = float(np.floor(root.x[0][S].sum()))
    rows = np.zeros((k, d.n))
    rows[0, S], rows[1, S] = 1.0, -1.0
    lp.append_rows(rows, np.array([K, -(K - 1.0)]))
    mask = np.zeros((B, k), dtype=np.uint8)
    bad = np.arange(B) % 5 == 2
    mask[bad] = 1                       # both rows on: infeasible
    mask[np.arange(B) % 5 == 3, 0] = 1  # only the first row: feasible
    X0, Y0 = np.tile(root.x[0], (B, 1)), np.tile(np.concatenate([root.y[0], np.zeros(k)]), (B, 1))
    res = [lp.solve_batch(lbs, ubs, row_mask=mask, x0=X0, y0=Y0, opts=eng.default_opts(**kw))
           for kw in (dict(freeze=0), dict(step_safety=0.0), dict())]
    lp.close()
    off = res[0]
    assert (off.status[bad] == 1).all() and (off.status[~bad] == 0).all(), off.status
    for r in res[1:]:
        assert np.array_equal(r.status, off.status)
        assert np.allclose(r.objective[~bad], off.objective[~bad], rtol=2e-7, atol=0)
        assert r.stats['skipped_col_updates'] > 0
