"""CutGeneratingLP — the disjunctive cut generating LP of a branch and bound subtree.

Contract of the reference's ``simple_mip_solver/utils/cut_generating_lp.py`` (constructor :13-50,
``_create_cglp`` :52-177, ``solve`` :179-221): the leaves of the subtree of ``bb`` rooted at
``root_id`` are the terms of a disjunction; the CGLP finds a cut ``pi.x >= pi0`` valid for the
convex hull of the terms' LP relaxations that is violated as much as possible by ``x_star``:

    min  x_star.pi - pi0
    s.t. pi  >= A_t' u_t + w_t - v_t                 for every term t      (n rows each)
         pi0 <= b_t.u_t + l_t.w_t - u_t.v_t                                 (1 row each)
         sum of all u, w, v = 1,   u, w, v >= 0   (w / v fixed to 0 where the bound is infinite)

The reference solves this tiny LP with CLP's primal simplex (:213) — SURVEY.md marks that solve
out of scope of the hot path. Here the model is assembled directly in the engine's canonical form
``min c.z, M z >= r, lo <= z <= hi`` and solved by the same batched GPU bound step (batch of one).
"""
from __future__ import annotations

from typing import Iterable, Tuple, Union

import numpy as np
import scipy.sparse as sp

from simple_mip_solver_b200.compat.cylp_like import CyLPArray

_INF = float('inf')


class _CglpLP:
    """What callers read from ``cglp.lp`` (the reference keeps a CyClpSimplex there)."""

    def __init__(self, n_vars: int, n_rows: int):
        self.nVariables = n_vars
        self.nConstraints = n_rows
        self._basis = None
        self._status = -1
        self.objectiveValue = 0.0
        self.logLevel = 0

    def getStatusCode(self):
        return self._status

    def getBasisStatus(self):
        if self._basis is None:
            return (np.full(self.nVariables, 3, dtype=np.int32), np.full(self.nConstraints, 1, dtype=np.int32))
        return self._basis[0].copy(), self._basis[1].copy()

    def setBasisStatus(self, cols, rows):
        self._basis = (np.asarray(cols, dtype=np.int32).copy(), np.asarray(rows, dtype=np.int32).copy())


class CutGeneratingLP:

    def __init__(self, bb, root_id: int, A=None, b: CyLPArray = None, var_lb: CyLPArray = None,
                 var_ub: CyLPArray = None, depth: int = None):
        from simple_mip_solver_b200.algorithms.branch_and_bound import BranchAndBound
        assert isinstance(bb, BranchAndBound), 'bb must be a BranchAndBound instance'
        assert root_id in bb.tree, 'root node of the disjunction must be present in B & B tree'
        if depth is not None:
            assert isinstance(depth, int) and depth > 0, 'depth is postive integer'
        self.bb = bb
        self.root_id = root_id
        self.depth = depth
        self.cylp_failure = False
        self._create_cglp(A, b, var_lb, var_ub)

    def _create_cglp(self, A, b, var_lb, var_ub) -> None:
        terms = {n.idx: n for n in self.bb.tree.get_leaves(self.root_id, depth=self.depth, keep='not infeasible')}
        root = self.bb.tree.get_node_instances(self.root_id)
        assert root.solution is not None, 'root must be solved to create CGLP'
        n = root.lp.nVariables
        assert all(t.lp.nVariables == n for t in terms.values()), \
            'Each disjunctive term should have the same variables. The feature allowing' \
            ' otherwise remains to be developed.'
        assert (A is None and b is None) or (A is not None and b is not None), \
            "A and b must both have values or must both be None"
        if A is not None:
            assert isinstance(A, np.matrix) or sp.issparse(A), "A must be a numpy or sparse csc matrix"
            assert A.shape[1] == n, "A must have same number of columns as each disjunctive term has variables"
            assert isinstance(b, CyLPArray), "b must be a CyLPArray"
            assert b.shape == (A.shape[0],), "A must have the same number of rows as b has entries"
        if var_lb is not None:
            assert isinstance(var_lb, CyLPArray), "var_lb must be a CyLPArray"
            assert var_lb.shape == (n,), "Must have same number of lower bounds as variables"
        else:
            var_lb = np.full(n, -_INF)
        if var_ub is not None:
            assert isinstance(var_ub, CyLPArray), "var_ub must be a CyLPArray"
            assert var_ub.shape == (n,), "Must have same number of upper bounds as variables"
        else:
            var_ub = np.full(n, _INF)

        inf = 1e30          # the engine's convention: |bound| >= 1e30 means no bound (CLP: DBL_MAX)
        blocks = []                       # per term: (A_t csr, b_t, l_t, u_t, has_l, has_u)
        self.term_ids = []
        for idx, node in terms.items():
            lo = np.maximum(np.asarray(node.lp.variablesLower, dtype=float), var_lb)
            hi = np.minimum(np.asarray(node.lp.variablesUpper, dtype=float), var_ub)
            if (lo > hi).any():
                continue                  # the extra bounds empty this term
            has_l, has_u = lo > -inf, hi < inf
            At = sp.csr_matrix(A if A is not None else node.lp.coefMatrix, dtype=float)
            bt = np.asarray(b if b is not None else node.lp.constraintsLower, dtype=float)
            blocks.append((At, bt, np.where(has_l, lo, 0.0), np.where(has_u, hi, 0.0), has_l, has_u))
            self.term_ids.append(idx)
        self.n = n
        self._x_root = np.asarray(root.solution, dtype=float).copy()

        # variable vector z = [pi (n) | pi0 | u_1 w_1 v_1 | u_2 w_2 v_2 | ...]
        ncol = n + 1 + sum(At.shape[0] + 2 * n for At, *_ in blocks)
        rows, rhs = [], []
        lo_z = np.full(ncol, 0.0)
        hi_z = np.full(ncol, _INF)
        lo_z[:n + 1] = -_INF                                   # pi, pi0 free
        ones = np.zeros(ncol)
        eye = sp.identity(n, format='csr')
        off = n + 1
        for At, bt, lt, ut, has_l, has_u in blocks:
            m = At.shape[0]
            iu, iw, iv = off, off + m, off + m + n
            # pi - A_t' u - w + v >= 0
            blk = sp.lil_matrix((n, ncol))
            blk[:, :n] = eye
            blk[:, iu:iu + m] = -At.T
            blk[:, iw:iw + n] = -eye
            blk[:, iv:iv + n] = eye
            rows.append(blk.tocsr())
            rhs.append(np.zeros(n))
            # -pi0 + b.u + l.w - u.v >= 0
            r = sp.lil_matrix((1, ncol))
            r[0, n] = -1.0
            r[0, iu:iu + m] = bt
            r[0, iw:iw + n] = lt
            r[0, iv:iv + n] = -ut
            rows.append(r.tocsr())
            rhs.append(np.zeros(1))
            hi_z[iw:iw + n] = np.where(has_l, _INF, 0.0)
            hi_z[iv:iv + n] = np.where(has_u, _INF, 0.0)
            ones[iu:iv + n] = 1.0
            off = iv + n
        # normalisation  sum(u, w, v) = 1  as two >= rows
        rows.append(sp.csr_matrix(ones[None, :]))
        rhs.append(np.ones(1))
        rows.append(sp.csr_matrix(-ones[None, :]))
        rhs.append(-np.ones(1))
        self._M = sp.vstack(rows, format='csr') if blocks else sp.csr_matrix((0, ncol))
        self._r = np.concatenate(rhs)
        self._lo, self._hi = lo_z, hi_z
        self.lp = _CglpLP(ncol, self._M.shape[0])

    def solve(self, x_star: CyLPArray = None, starting_basis: Tuple[np.ndarray, np.ndarray] = None) -> \
            Tuple[Union[CyLPArray, None], Union[float, None]]:
        """The valid inequality that separates ``x_star`` (default: the root node's LP solution) most
        from the convex hull of the disjunctive terms, or (None, None) if the solve fails."""
        if x_star is not None:
            assert isinstance(x_star, CyLPArray), 'x_star must be a CyLPArray'
            assert x_star.shape == (self.n,), \
                'x_star must have the same number of variables as the LP relaxations ' \
                'in the branch and bound tree this instance was created with'
        if starting_basis is not None:
            assert isinstance(starting_basis, Iterable) and not isinstance(starting_basis, str) \
                and len(starting_basis) == 2, 'starting basis must be an iterable with two elements'
            for status_array in starting_basis:
                assert isinstance(status_array, np.ndarray), 'elements of starting basis must be np.ndarrays'
            assert starting_basis[0].shape == (self.lp.nVariables,), \
                'first starting_basis element should give status for exactly each decision variable in CGLP'
            assert starting_basis[1].shape == (self.lp.nConstraints,), \
                'second starting_basis element should give status for exactly each slack variable in CGLP'
            self.lp.setBasisStatus(*starting_basis)
        if not self.term_ids:
            self.cylp_failure = True
            return None, None
        from simple_mip_solver_b200 import engine
        x = self._x_root if x_star is None else np.asarray(x_star, dtype=float)
        c = np.zeros(self.lp.nVariables)
        c[:self.n] = x
        c[self.n] = -1.0
        lp = engine.BatchLP(self._M, self._r, c, device=getattr(self.bb.model.lp._shared, 'device', 0))
        try:
            res = lp.solve_batch(self._lo[None], self._hi[None],
                                 opts=engine.default_opts(eps_rel=1e-9, max_iters=2_000_000))
        finally:
            lp.close()
        self.lp._status = int(res.status[0])
        self.lp.objectiveValue = float(res.objective[0])
        if self.lp._status in (0, 2):
            z = res.x[0]
            return CyLPArray(z[:self.n]), float(z[self.n])
        self.cylp_failure = True
        return None, None
