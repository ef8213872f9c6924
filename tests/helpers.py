"""Test helpers: a CPU stand-in for engine.BatchLP backed by the HiGHS oracle.

TEST INFRASTRUCTURE ONLY. It lets the host-side logic (Node classes, BranchAndBound, the batched
LP plumbing of compat/cylp_like.py) run in the CPU test suite, where no GPU exists, with exact LP
answers. The product never imports this file; on a GPU box the same host code talks to libblp.so.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from oracle.dual_simplex import dual_simplex
from oracle.highs_lp import HIGHS_INF, HighsLP


class OracleBatchLP:
    """Same call surface as simple_mip_solver_b200.engine.BatchLP (the part compat uses)."""
    calls = 0
    lps = 0
    batch_sizes = []

    def __init__(self, A, b, c, device=0):
        self.A = sp.csr_matrix(A, dtype=float)
        self.m_base, self.n = self.A.shape
        self.b = np.asarray(b, float)
        self.c = np.asarray(c, float)
        self.cut_rows = []
        self.cut_rhs = []

    @property
    def m(self):
        return self.m_base + len(self.cut_rows)

    def append_rows(self, rows, rhs):
        first = self.m
        rows = np.atleast_2d(np.asarray(rows.todense()) if sp.issparse(rows) else rows)
        for r, v in zip(rows, np.atleast_1d(rhs)):
            self.cut_rows.append(np.asarray(r, float))
            self.cut_rhs.append(float(v))
        return first

    def truncate_rows(self, m_keep):
        keep = m_keep - self.m_base
        del self.cut_rows[keep:]
        del self.cut_rhs[keep:]
        self._store = ()

    def close(self):
        pass

    # -- dual simplex path: the numpy restatement of the device kernel (oracle/dual_simplex.py), with
    #    the engine's two-call factor store emulated so that parent_slot means the same thing
    simplex_capable = True
    simplex_batched = True
    _store = ()

    def _full(self):
        A = self.A.toarray()
        if self.cut_rows:
            A = np.vstack([A] + [r[None, :] for r in self.cut_rows])
        return A, np.concatenate([self.b, self.cut_rhs])

    def _simplex(self, lbs, ubs, masks, css, rss, parents, max_pivots):
        from simple_mip_solver_b200.engine import SimplexBatchResult
        A, b = self._full()
        m, B = A.shape[0], len(lbs)
        type(self).calls += 1
        type(self).lps += B
        type(self).batch_sizes.append(B)
        out, pivots = [], 0
        for k in range(B):
            on = np.ones(m, bool)
            if masks is not None:
                on[self.m_base:] = np.asarray(masks[k], bool)
            start = None
            if parents is not None and parents[k] >= 0:
                start = self._store[parents[k]]
                assert start.Binv.shape[0] == m, 'the factor store was dropped when rows changed'
            r = dual_simplex(A, b, self.c, lbs[k], ubs[k], row_on=on,
                             col_status=None if css is None else css[k],
                             row_status=None if rss is None else rss[k], max_pivots=max_pivots, start=start)
            out.append(r)
            pivots += r.pivots
        self._store = tuple(out)
        return SimplexBatchResult(
            objective=np.array([r.objective for r in out]), status=np.array([r.status for r in out], np.int32),
            pivots=np.array([r.pivots for r in out], np.int32), x=np.array([r.x for r in out]),
            y=np.array([r.y for r in out]), reduced_costs=np.array([r.rc for r in out]),
            col_status=np.array([r.col_status for r in out], np.int8),
            row_status=np.array([r.row_status for r in out], np.int8),
            stats=dict(kernel_launches=0, iterations=max(r.pivots for r in out), node_iterations=pivots))

    def simplex_batch(self, lb, ub, row_mask=None, col_status=None, row_status=None, parent_slot=None,
                      max_pivots=2147483647):
        lb, ub = np.atleast_2d(lb), np.atleast_2d(ub)
        return self._simplex(lb, ub, row_mask, col_status, row_status, parent_slot, max_pivots)

    def simplex_children(self, parent_lb, parent_ub, deltas, row_mask=None, col_status=None, row_status=None,
                         parent_slot=-1, max_pivots=2147483647):
        B = len(deltas)
        lb = np.tile(np.asarray(parent_lb, float), (B, 1))
        ub = np.tile(np.asarray(parent_ub, float), (B, 1))
        for k, d in enumerate(deltas):
            for j, lo, hi in d:
                lb[k, j], ub[k, j] = lo, hi
        rep = lambda a: None if a is None else np.tile(np.asarray(a), (B, 1))
        return self._simplex(lb, ub, rep(row_mask), rep(col_status), rep(row_status),
                             None if parent_slot < 0 else np.full(B, parent_slot), max_pivots)

    def simplex_tableau_rows(self, slot, variables):
        r = self._store[slot]
        A, _ = self._full()
        m = A.shape[0]
        full = np.concatenate([A, -np.eye(m)], axis=1)
        pos = {int(v): i for i, v in enumerate(r.head)}
        out = np.zeros((len(variables), full.shape[1]))
        for t, v in enumerate(np.asarray(variables).ravel()):
            if int(v) in pos:
                out[t] = r.Binv[pos[int(v)]] @ full
        return out

    def solve_batch(self, lb, ub, row_mask=None, x0=None, y0=None, integer_indices=None, opts=None,
                    want_x=True, want_y=True):
        from simple_mip_solver_b200.engine import BatchResult
        lb, ub = np.atleast_2d(lb), np.atleast_2d(ub)
        B = lb.shape[0]
        type(self).calls += 1
        type(self).lps += B
        type(self).batch_sizes.append(B)
        m = self.m
        obj = np.full(B, np.inf); lower = np.full(B, np.inf)
        status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); frac = np.full(B, -1, np.int32)
        x = np.zeros((B, self.n)); y = np.zeros((B, m))
        for k in range(B):
            h = HighsLP(self.A, self.c, self.b, np.full(self.m_base, HIGHS_INF), lb[k], ub[k])
            on = [j for j in range(len(self.cut_rows)) if row_mask is None or row_mask[k, j]]
            for j in on:
                h.add_row(self.cut_rows[j], self.cut_rhs[j])
            r = h.solve()
            status[k] = r.status
            iters[k] = r.iterations
            if r.status == 0:
                obj[k] = lower[k] = r.objective
                x[k] = r.x
                y[k, :self.m_base] = r.row_dual[:self.m_base]
                for t, j in enumerate(on):
                    y[k, self.m_base + j] = r.row_dual[self.m_base + t]
                if integer_indices is not None and len(integer_indices):
                    ii = np.asarray(sorted(integer_indices))
                    d = np.minimum(x[k, ii] - np.floor(x[k, ii]), np.ceil(x[k, ii]) - x[k, ii])
                    if d.max() > 1e-4:
                        frac[k] = ii[int(np.argmax(d))]
            elif r.status == 2:
                obj[k] = lower[k] = -np.inf
                x[k] = r.x if r.x is not None else 0
        return BatchResult(objective=obj, lower_bound=lower, status=status, iterations=iters,
                           frac_idx=frac, x=x, y=y, stats=dict(kernel_launches=0))


def _solve_children(self, parent_lb, parent_ub, deltas, row_mask=None, x0=None, y0=None, integer_indices=None,
                    opts=None, want_x=True, want_y=True):
    B = len(deltas)
    lb = np.tile(np.asarray(parent_lb, float), (B, 1))
    ub = np.tile(np.asarray(parent_ub, float), (B, 1))
    for k, d in enumerate(deltas):
        for j, lo, hi in d:
            lb[k, j], ub[k, j] = lo, hi
    type(self).children_calls = getattr(type(self), 'children_calls', 0) + 1
    return self.solve_batch(lb, ub, row_mask=row_mask, integer_indices=integer_indices, opts=opts)


OracleBatchLP.solve_children = _solve_children


def use_oracle_engine(monkeypatch, method='auto'):
    """Route SharedLP's engine to the CPU stand-in (tests of host logic): the numpy dual simplex for
    the simplex path, HiGHS for the path that goes to the PDHG kernels (``method='pdhg'``)."""
    import simple_mip_solver_b200.engine as engine
    from simple_mip_solver_b200.compat.cylp_like import SharedLP
    monkeypatch.setattr(SharedLP, 'default_method', method)
    OracleBatchLP.calls = 0
    OracleBatchLP.lps = 0
    OracleBatchLP.batch_sizes = []
    OracleBatchLP.children_calls = 0
    monkeypatch.setattr(engine, 'BatchLP', OracleBatchLP)
    monkeypatch.setattr(engine, 'default_opts', lambda **kw: kw)
    return OracleBatchLP
