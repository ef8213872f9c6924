"""Test helpers: a CPU stand-in for engine.BatchLP backed by the HiGHS oracle.

TEST INFRASTRUCTURE ONLY. It lets the host-side logic (Node classes, BranchAndBound, the batched
LP plumbing of compat/cylp_like.py) run in the CPU test suite, where no GPU exists, with exact LP
answers. The product never imports this file; on a GPU box the same host code talks to libblp.so.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

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

    def close(self):
        pass

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


def use_oracle_engine(monkeypatch):
    """Route SharedLP's engine to the HiGHS stand-in (CPU tests of host logic)."""
    import simple_mip_solver_b200.engine as engine
    OracleBatchLP.calls = 0
    OracleBatchLP.lps = 0
    OracleBatchLP.batch_sizes = []
    monkeypatch.setattr(engine, 'BatchLP', OracleBatchLP)
    monkeypatch.setattr(engine, 'default_opts', lambda **kw: kw)
    return OracleBatchLP
