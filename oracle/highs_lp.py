"""TEST INFRASTRUCTURE — exact CPU solve of the reference's node LP (stand-in for CLP).

The reference solves every node LP with COIN-OR CLP's dual simplex through CyLP
(``simple_mip_solver/nodes/base_node.py:273`` ``self.lp.dual()``, and ``:645-646`` with
``maxNumIteration`` for strong branching). CLP/CyLP are third-party, unpinned
(environment.yml:7,15-16) and absent from this image, so the exact LP arithmetic is delegated to
HiGHS 1.12 dual simplex (bundled with scipy 1.18, reached through its private binding), used the
way the reference uses CLP: presolve off, warm start from the parent's basis, optional pivot limit.
LP optimal values are solver independent, which is what the 1e-6 parity bar is measured against.
The choice among alternative optimal vertices is not; see DESIGN.md ("vertex parity").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (simple_mip_solver_b200/) never does.

CLP status codes are kept (base_node.py:274-275, pseudo_cost.py:86):
  0 optimal, 1 primal infeasible, 2 dual infeasible (unbounded), 3 iteration limit, -1 unknown.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp
from scipy.optimize._highspy import _core as _hc

HIGHS_INF = _hc.kHighsInf
_BASIS_TO_CLP = {  # HiGHS basis status -> CLP getBasisStatus code (1 basic, 2 at upper, 3 at lower, 0 free)
    _hc.HighsBasisStatus.kLower: 3, _hc.HighsBasisStatus.kBasic: 1, _hc.HighsBasisStatus.kUpper: 2,
    _hc.HighsBasisStatus.kZero: 0, _hc.HighsBasisStatus.kNonbasic: 3,
}
_CLP_TO_BASIS = {1: _hc.HighsBasisStatus.kBasic, 2: _hc.HighsBasisStatus.kUpper,
                 3: _hc.HighsBasisStatus.kLower, 0: _hc.HighsBasisStatus.kZero,
                 4: _hc.HighsBasisStatus.kLower, 5: _hc.HighsBasisStatus.kLower}


@dataclass
class LpSolution:
    status: int
    objective: float
    x: Optional[np.ndarray]
    row_dual: Optional[np.ndarray]
    reduced_cost: Optional[np.ndarray]
    col_basis: Optional[np.ndarray]      # CLP codes, int32
    row_basis: Optional[np.ndarray]
    iterations: int


def _inf(v):
    v = np.asarray(v, dtype=float).copy()
    v[v >= 1e300] = HIGHS_INF
    v[v <= -1e300] = -HIGHS_INF
    return v


class HighsLP:
    """One LP ``min c.x, row_lb <= A x <= row_ub, l <= x <= u`` held in a HiGHS instance."""

    def __init__(self, A, c, row_lb, row_ub, l, u, threads: int = 1, tol: float = 1e-9):
        A = sp.csc_matrix(A, dtype=float)
        m, n = A.shape
        self.m, self.n = m, n
        h = _hc._Highs()
        h.setOptionValue('output_flag', False)
        h.setOptionValue('presolve', 'off')
        h.setOptionValue('solver', 'simplex')
        h.setOptionValue('simplex_strategy', 1)          # serial dual simplex, as CLP's dual()
        h.setOptionValue('threads', threads)
        h.setOptionValue('primal_feasibility_tolerance', float(tol))
        h.setOptionValue('dual_feasibility_tolerance', float(tol))
        lp = _hc.HighsLp()
        lp.num_col_, lp.num_row_ = n, m
        lp.col_cost_ = np.asarray(c, dtype=float)
        lp.col_lower_, lp.col_upper_ = _inf(l), _inf(u)
        lp.row_lower_, lp.row_upper_ = _inf(row_lb), _inf(row_ub)
        lp.a_matrix_.format_ = _hc.MatrixFormat.kColwise
        lp.a_matrix_.start_ = A.indptr.astype(np.int32)
        lp.a_matrix_.index_ = A.indices.astype(np.int32)
        lp.a_matrix_.value_ = A.data
        st = h.passModel(lp)
        assert st != _hc.HighsStatus.kError, 'HiGHS rejected the model'
        self.h = h

    def set_col_bounds(self, l, u):
        l, u = _inf(l), _inf(u)
        idx = np.arange(self.n, dtype=np.int32)
        self.h.changeColsBounds(self.n, idx, l, u)

    def set_one_col_bound(self, j, lo, up):
        self.h.changeColBounds(int(j), float(_inf([lo])[0]), float(_inf([up])[0]))

    def add_row(self, coefs: np.ndarray, lo: float, up: float = HIGHS_INF):
        nz = np.flatnonzero(coefs)
        self.h.addRow(float(lo), float(up), len(nz), nz.astype(np.int32), np.asarray(coefs, float)[nz])
        self.m += 1

    def delete_rows_from(self, first: int):
        if first < self.m:
            self.h.deleteRows(self.m - first, np.arange(first, self.m, dtype=np.int32))
            self.m = first

    def set_basis(self, col_basis: Sequence[int], row_basis: Sequence[int]):
        b = _hc.HighsBasis()
        b.col_status = [_CLP_TO_BASIS[int(s)] for s in col_basis]
        b.row_status = [_CLP_TO_BASIS[int(s)] for s in row_basis]
        b.valid = True
        self.h.setBasis(b)

    def clear_basis(self):
        self.h.clearSolver()

    def solve(self, iteration_limit: Optional[int] = None) -> LpSolution:
        h = self.h
        h.setOptionValue('simplex_iteration_limit',
                         int(iteration_limit) if iteration_limit is not None else 2147483647)
        h.run()
        ms = h.getModelStatus()
        info = h.getInfo()
        S = _hc.HighsModelStatus
        if ms == S.kOptimal:
            code = 0
        elif ms == S.kInfeasible:
            code = 1
        elif ms in (S.kUnbounded,):
            code = 2
        elif ms == S.kUnboundedOrInfeasible:
            code = 2
        elif ms == S.kIterationLimit:
            code = 3
        else:
            code = -1
        sol = h.getSolution()
        x = np.asarray(sol.col_value, dtype=float) if len(sol.col_value) else None
        y = np.asarray(sol.row_dual, dtype=float) if len(sol.row_dual) else None
        rc = np.asarray(sol.col_dual, dtype=float) if len(sol.col_dual) else None
        basis = h.getBasis()
        cb = np.array([_BASIS_TO_CLP[s] for s in basis.col_status], dtype=np.int32) \
            if len(basis.col_status) else None
        rb = np.array([_BASIS_TO_CLP[s] for s in basis.row_status], dtype=np.int32) \
            if len(basis.row_status) else None
        return LpSolution(status=code, objective=float(info.objective_function_value), x=x,
                          row_dual=y, reduced_cost=rc, col_basis=cb, row_basis=rb,
                          iterations=int(info.simplex_iteration_count))


def solve_node_lps(A, b, c, lbs, ubs, warm_from_root=True, iteration_limit=None,
                   root_l=None, root_u=None, extra_rows=None, row_masks=None):
    """Solve a batch of node LPs ``min c.x, A x >= b, lbs[k] <= x <= ubs[k]`` one by one.

    lbs/ubs: [B, n]. With ``warm_from_root`` the root LP (root_l/root_u, default = elementwise
    hull of the batch) is solved first and every node starts from its optimal basis, mirroring
    ``lp.setBasisStatus(*basis)`` in ``_base_branch`` (base_node.py:589,608).
    Returns a list of LpSolution.
    """
    A = sp.csr_matrix(A)
    m, n = A.shape
    lbs = np.atleast_2d(lbs)
    ubs = np.atleast_2d(ubs)
    rl = lbs.min(axis=0) if root_l is None else root_l
    ru = ubs.max(axis=0) if root_u is None else root_u
    lp = HighsLP(A, c, b, np.full(m, HIGHS_INF), rl, ru)
    basis = None
    if warm_from_root:
        r = lp.solve()
        if r.status == 0:
            basis = (r.col_basis, r.row_basis)
    out = []
    for k in range(lbs.shape[0]):
        lp.set_col_bounds(lbs[k], ubs[k])
        if basis is not None:
            lp.set_basis(*basis)
        out.append(lp.solve(iteration_limit))
    return out


def _first_raw_run():
    """A raw HiGHS instance has to run once BEFORE scipy.optimize.milp is first used in a process: the other
    way round, every raw instance created afterwards returns kNotset from run() (seen with scipy 1.18's bundled
    HiGHS; found while generating tests/golden/fuzz_models.json). Importing this module settles the order."""
    one = HighsLP(np.eye(1), np.ones(1), np.ones(1), np.full(1, HIGHS_INF), np.zeros(1), np.full(1, 2.0)).solve()
    if one.status != 0 or abs(one.objective - 1.0) > 1e-12:
        raise RuntimeError('raw HiGHS instances do not run in this process (scipy.optimize.milp was used before '
                           'oracle.highs_lp was imported?): import the oracle first')


_first_raw_run()

