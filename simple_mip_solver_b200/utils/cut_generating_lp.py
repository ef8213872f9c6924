"""CutGeneratingLP — the disjunctive cut generating LP of a branch and bound subtree.

Contract of the reference's ``simple_mip_solver/utils/cut_generating_lp.py`` (constructor :13-50,
``_create_cglp`` :52-177, ``solve`` :179-221): the leaves of the subtree of ``bb`` rooted at
``root_id`` are the terms of a disjunction; the CGLP finds a cut ``pi.x >= pi0`` valid for the
convex hull of the terms' LP relaxations that is violated as much as possible by ``x_star``:

    min  x_star.pi - pi0
    s.t. pi  >= A_t' u_t + w_t - v_t                 for every term t      (n rows each)
         pi0 <= b_t.u_t + l_t.w_t - u_t.v_t                                 (1 row each)
         sum of all u, w, v = 1,   u, w, v >= 0   (w / v fixed to 0 where the bound is infinite)

The reference solves this LP with CLP's primal simplex (:213), one ``x_star`` at a time: only the
OBJECTIVE changes between the calls of one CGLP (:199). The engine batches LPs that differ in their
variable BOUNDS, so the LP handed to the GPU is the CGLP's dual, where ``x_star`` is a bound:

    max  gamma                                          (= x_star.pi - pi0 at the optimum)
    s.t. sum_t xi_t - s = 0,          s fixed to x_star by its bounds           multipliers -> pi
         sum_t lambda_t = 1                                                     multiplier  -> pi0
         A_t xi_t - b_t lambda_t >= gamma                                       multipliers = u_t
         xi_t - l_t lambda_t >= gamma,   h_t lambda_t - xi_t >= gamma           multipliers = w_t, v_t
         xi_t, lambda_t >= 0, gamma free

(a point of the convex hull of the terms, written as sum of scaled points ``xi_t`` of the terms, and
the largest margin ``gamma`` by which all of them satisfy their term). One device matrix per CGLP, any
number of points ``x_star`` per call (``solve_batch`` / ``prefetch``): the cut
``(pi, pi0)`` of every point is read from the row multipliers of its LP. Small CGLPs (up to
``blp_simplex_batch_rows()`` rows) go to the batched dual simplex kernel — exact vertices, and the
basis of one solve warm-starts the next, as ``starting_basis`` does in the reference —, larger ones to
the PDHG kernels. ``cglp.lp`` keeps the shape of the reference's primal model (``nVariables``,
``nConstraints``, CLP-coded basis arrays) for callers that size things by it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.sparse as sp

from simple_mip_solver_b200.compat.cylp_like import CyLPArray

_INF = float('inf')
SIMPLEX_CHUNK_BYTES = 4 << 30       # dense basis inverses one simplex call may hold (8 m^2 bytes per point)
PREFETCH_CACHE_POINTS = 4096        # answers kept for nodes that have not asked yet
ROUND_OFF_BELOW_ZERO = 1e-9


class BasisArray(np.ndarray):
    """CLP-coded status array (what ``getBasisStatus`` returns) that also carries the basis of the
    LP the device actually solved, so that handing it back through ``starting_basis`` /
    ``prev_cglp_basis`` restarts the dual simplex kernel exactly where it stopped."""

    def __new__(cls, data, device_basis=None):
        obj = np.asarray(data, dtype=np.int32).view(cls)
        obj.device_basis = device_basis
        return obj

    def __array_finalize__(self, obj):
        self.device_basis = getattr(obj, 'device_basis', None)


@dataclass
class CglpAnswer:
    """One point's cut, as ``solve`` reports it."""
    pi: Optional[np.ndarray]
    pi0: Optional[float]
    status: int                 # CLP code of the CGLP: 0 optimal, 2 unbounded (x_star outside x >= 0), 3 limit
    objective: float            # x_star.pi - pi0 (negative = x_star is cut off)
    iterations: int
    device_basis: Optional[Tuple[np.ndarray, np.ndarray]]


class _CglpLP:
    """What callers read from ``cglp.lp`` (the reference keeps a CyClpSimplex there)."""

    def __init__(self, n_vars: int, n_rows: int):
        self.nVariables = n_vars
        self.nConstraints = n_rows
        self._basis = None
        self._status = -1
        self.objectiveValue = 0.0
        self.iteration = 0
        self.logLevel = 0

    def getStatusCode(self):
        return self._status

    def getBasisStatus(self):
        if self._basis is None:
            return (np.full(self.nVariables, 3, dtype=np.int32), np.full(self.nConstraints, 1, dtype=np.int32))
        return self._basis[0].copy(), self._basis[1].copy()

    def setBasisStatus(self, cols, rows):
        keep = lambda a: a.copy() if isinstance(a, BasisArray) else np.asarray(a, dtype=np.int32).copy()
        self._basis = (keep(cols), keep(rows))


class CutGeneratingLP:
    # which device path solves the CGLP: 'auto' = the batched dual simplex while the device LP has at
    # most blp_simplex_batch_rows() rows, PDHG beyond; 'pdhg' forces the first-order path
    method = 'auto'

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

        # the reference's model: z = [pi (n) | pi0 | u_1 w_1 v_1 | u_2 w_2 v_2 | ...], per term n + 1 rows, then the
        # normalisation as two rows. Only its SIZES are needed to answer (cglp.lp); the model itself is
        # assembled on demand (`primal_model`, read by the parity tests)
        self._blocks = blocks
        ncol = n + 1 + sum(At.shape[0] + 2 * n for At, *_ in blocks)
        self.lp = _CglpLP(ncol, len(blocks) * (n + 1) + 1 if blocks else 0)     # one normalisation row (:169-171)
        self._primal = None
        self._create_dual_form(blocks)

    def primal_model(self):
        """``(M, r, lo, hi)`` of the reference's CGLP in the engine's canonical form ``min c.z, M z >= r,
        lo <= z <= hi`` with ``c = [x_star, -1, 0...]`` (reference :52-177); the normalisation, one equality
        row in the reference, is two ``>=`` rows here, so ``M`` has ``cglp.lp.nConstraints + 1`` rows. The device
        never sees it."""
        if self._primal is None:
            n, blocks = self.n, self._blocks
            ncol = self.lp.nVariables
            lo_z = np.zeros(ncol)
            hi_z = np.full(ncol, _INF)
            lo_z[:n + 1] = -_INF                                   # pi, pi0 free
            ones = np.zeros(ncol)
            eye = sp.identity(n, format='csr')
            rows, off = [], n + 1
            for At, bt, lt, ut, has_l, has_u in blocks:
                m = At.shape[0]
                pad_l, pad_r = off - (n + 1), ncol - (off + m + 2 * n)
                z = lambda k, w: sp.csr_matrix((k, w))
                row = lambda v: sp.csr_matrix(np.asarray(v, dtype=float)[None, :])
                # pi - A_t' u - w + v >= 0
                rows.append(sp.hstack([eye, z(n, 1 + pad_l), -At.T, -eye, eye, z(n, pad_r)], format='csr'))
                # -pi0 + b.u + l.w - u.v >= 0
                rows.append(sp.hstack([z(1, n), row([-1.0]), z(1, pad_l), row(bt), row(lt), row(-ut), z(1, pad_r)],
                                      format='csr'))
                hi_z[off + m:off + m + n] = np.where(has_l, _INF, 0.0)
                hi_z[off + m + n:off + m + 2 * n] = np.where(has_u, _INF, 0.0)
                ones[off:off + m + 2 * n] = 1.0
                off += m + 2 * n
            # normalisation  sum(u, w, v) = 1  as two >= rows
            rows += [sp.csr_matrix(ones[None, :]), sp.csr_matrix(-ones[None, :])]
            M = sp.vstack(rows, format='csr') if blocks else sp.csr_matrix((0, ncol))
            r = np.zeros(M.shape[0])
            if blocks:
                r[-2], r[-1] = 1.0, -1.0
            self._primal = (M, r, lo_z, hi_z)
        return self._primal

    _M = property(lambda self: self.primal_model()[0])
    _r = property(lambda self: self.primal_model()[1])
    _lo = property(lambda self: self.primal_model()[2])
    _hi = property(lambda self: self.primal_model()[3])

    def _create_dual_form(self, blocks) -> None:
        """The LP the device solves (module docstring): columns ``[xi_1 .. xi_T | lambda | gamma | s]``,
        rows ``[sum xi - s >= 0 | s - sum xi >= 0 | sum lambda >= 1 | -sum lambda >= -1 | per term: u, w, v rows]``,
        ``min -gamma``. Also records which column of the reference's primal model every u/w/v row
        is the multiplier of, for the CLP-coded basis arrays."""
        n, T = self.n, len(blocks)
        self._engine = None
        self._prefetched = {}
        self.batch_calls = 0            # device calls / points solved through them (monitoring, tests)
        self.points_solved = 0
        self.prefetch_hits = 0
        if not T:
            return
        c_lam, c_gam, c_s = T * n, T * n + T, T * n + T + 1
        ncol = c_s + n
        eye = sp.identity(n, format='csr')
        link = sp.hstack([eye] * T + [sp.csr_matrix((n, T + 1)), -eye], format='csr')
        lam = sp.csr_matrix((np.ones(T), (np.zeros(T, dtype=int), np.arange(c_lam, c_lam + T))), shape=(1, ncol))
        rows, primal_col = [link, -link, lam, -lam], []
        off = n + 1                                     # first column of term t in the primal model
        for t, (At, bt, lt, ut, has_l, has_u) in enumerate(blocks):
            m = At.shape[0]
            jl, ju = np.flatnonzero(has_l), np.flatnonzero(has_u)

            def block(xi_part, lam_coef):
                k = xi_part.shape[0]
                return sp.hstack([sp.csr_matrix((k, t * n)), xi_part, sp.csr_matrix((k, (T - 1 - t) * n + t)),
                                  sp.csr_matrix(np.asarray(lam_coef, dtype=float).reshape(k, 1)),
                                  sp.csr_matrix((k, T - 1 - t)), sp.csr_matrix(-np.ones((k, 1))),
                                  sp.csr_matrix((k, n))], format='csr')
            rows += [block(At, -bt), block(eye[jl], -lt[jl]), block(-eye[ju], ut[ju])]
            primal_col += [off + np.arange(m), off + m + jl, off + m + n + ju]
            off += m + 2 * n
        self._dM = sp.vstack(rows, format='csr')
        self._dr = np.zeros(self._dM.shape[0])
        self._dr[2 * n], self._dr[2 * n + 1] = 1.0, -1.0
        self._dc = np.zeros(ncol)
        self._dc[c_gam] = -1.0
        self._dlo = np.zeros(ncol)
        self._dhi = np.full(ncol, _INF)
        self._dlo[c_gam] = -_INF                  # replaced per point by the box of _gamma_box
        # |gamma| <= this + (1 + largest row sum) * max|x_star| at every optimum (see _gamma_box)
        finite = lambda v, on: float(np.max(np.abs(v[on]))) if on.any() else 0.0
        self._d_scale = 1.0 + max([0.0] + [max(finite(bt, np.ones(len(bt), bool)), finite(lt, has_l), finite(ut, has_u))
                                           for _, bt, lt, ut, has_l, has_u in blocks])
        self._d_rowsum = max([0.0] + [float(abs(At).sum(axis=1).max()) if At.shape[0] else 0.0
                                      for At, *_ in blocks])
        self._d_cols = (c_lam, c_gam, c_s)
        self._d_primal_col = np.concatenate(primal_col).astype(np.int64)
        # primal-model rows of term t: n rows (multipliers xi_t) then one row (multiplier lambda_t)
        self._d_primal_row_of_xi = (np.arange(T)[:, None] * (n + 1) + np.arange(n)[None, :]).ravel()
        self._d_primal_row_of_lam = np.arange(T) * (n + 1) + n

    # ------------------------------------------------------------------ basis coding
    def _primal_coded(self, cs: np.ndarray, rs: np.ndarray) -> Tuple[BasisArray, BasisArray]:
        """CLP-coded status of the reference's primal model that is complementary to the basis
        ``(cs, rs)`` of the device LP: a primal column is basic where the row it multiplies is
        tight (its slack nonbasic), a primal row's slack is basic where its multiplier is nonbasic."""
        n, T = self.n, len(self.term_ids)
        c_lam, c_gam, _ = self._d_cols
        cols = np.full(self.lp.nVariables, 3, dtype=np.int32)
        cols[:n] = np.where((rs[:n] != 1) | (rs[n:2 * n] != 1), 1, 3)
        cols[n] = 1 if (rs[2 * n] != 1 or rs[2 * n + 1] != 1) else 3
        cols[self._d_primal_col] = np.where(rs[2 * n + 2:] != 1, 1, 3)
        rows = np.full(self.lp.nConstraints, 1, dtype=np.int32)
        rows[self._d_primal_row_of_xi] = np.where(cs[:T * n] == 1, 3, 1)
        rows[self._d_primal_row_of_lam] = np.where(cs[c_lam:c_lam + T] == 1, 3, 1)
        rows[-1] = 3 if cs[c_gam] == 1 else 1
        basis = (np.array(cs, dtype=np.int8), np.array(rs, dtype=np.int8))
        return BasisArray(cols, basis), BasisArray(rows, basis)

    def _device_coded(self, start) -> Tuple[np.ndarray, np.ndarray]:
        """Starting basis of the device LP from what ``setBasisStatus`` was given: the device basis the
        arrays carry if they came from ``getBasisStatus`` of a CGLP of this shape, otherwise the
        complement of the CLP-coded arrays (all-nonbasic columns / all-basic rows = a cold start)."""
        n, T = self.n, len(self.term_ids)
        c_lam, c_gam, _ = self._d_cols
        ncol, nrow = self._dM.shape[1], self._dM.shape[0]
        if start is not None:
            carried = getattr(start[0], 'device_basis', None)
            if carried is not None and carried[0].shape == (ncol,) and carried[1].shape == (nrow,):
                return carried
        cs = np.full(ncol, 3, dtype=np.int8)
        rs = np.full(nrow, 1, dtype=np.int8)
        if start is None:
            return cs, rs
        cols, rows = np.asarray(start[0]), np.asarray(start[1])
        cs[:T * n] = np.where(rows[self._d_primal_row_of_xi] == 1, 3, 1)
        cs[c_lam:c_lam + T] = np.where(rows[self._d_primal_row_of_lam] == 1, 3, 1)
        cs[c_gam] = 1 if rows[-1] != 1 else 3
        rs[:n] = np.where(cols[:n] == 1, 3, 1)
        rs[2 * n] = 3 if cols[n] == 1 else 1
        rs[2 * n + 2:] = np.where(cols[self._d_primal_col] == 1, 3, 1)
        return cs, rs

    # ------------------------------------------------------------------ device solves
    def _device_lp(self):
        if self._engine is None:
            from simple_mip_solver_b200 import engine
            shared = getattr(getattr(self.bb.model, 'lp', None), '_shared', None)
            self._engine = engine.BatchLP(self._dM, self._dr, self._dc, device=getattr(shared, 'device', 0))
        return self._engine

    def close(self) -> None:
        """Release the device copy of the CGLP (it is re-created by the next solve)."""
        if getattr(self, '_engine', None) is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _solve_points(self, points: np.ndarray, starts: Sequence) -> List[CglpAnswer]:
        """The cuts of ``points`` ([K, n]) in ONE device call per chunk: K LPs that share the matrix and
        differ in the bounds that fix ``s`` to their point."""
        from simple_mip_solver_b200 import engine
        lp = self._device_lp()
        K = points.shape[0]
        # round-off below zero (an LP solution read back from the device) is zero: a negative component
        # makes the CGLP unbounded — its dual has no point sum xi_t = x_star with xi_t >= 0
        points = np.where((points < 0) & (points > -ROUND_OFF_BELOW_ZERO), 0.0, points)
        lo, hi = self._point_bounds(points)
        nrow = self._dM.shape[0]
        self.points_solved += K
        answers: List[CglpAnswer] = []
        if self.method != 'pdhg' and getattr(lp, 'simplex_batched', False):
            coded = [self._device_coded(st) for st in starts]
            chunk = max(1, min(K, SIMPLEX_CHUNK_BYTES // (8 * nrow * nrow)))
            for a in range(0, K, chunk):
                e = min(K, a + chunk)
                res = lp.simplex_batch(lo[a:e], hi[a:e], col_status=np.stack([c[0] for c in coded[a:e]]),
                                       row_status=np.stack([c[1] for c in coded[a:e]]))
                self.batch_calls += 1
                for k in range(e - a):
                    answers.append(self._answer(int(res.status[k]), float(res.objective[k]), res.y[k],
                                                int(res.pivots[k]), (res.col_status[k], res.row_status[k])))
        else:
            res = lp.solve_batch(lo, hi, opts=engine.default_opts(eps_rel=1e-9, max_iters=2_000_000))
            self.batch_calls += 1
            for k in range(K):
                answers.append(self._answer(int(res.status[k]), float(res.objective[k]), res.y[k],
                                            int(res.iterations[k]), None))
        return answers

    def _point_bounds(self, points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Column bounds [K, columns] of the device LPs of ``points``: the only thing they differ in."""
        K = points.shape[0]
        _, c_gam, c_s = self._d_cols
        lo = np.tile(self._dlo, (K, 1))
        hi = np.tile(self._dhi, (K, 1))
        lo[:, c_s:] = hi[:, c_s:] = points
        box = self._gamma_box(points)
        lo[:, c_gam], hi[:, c_gam] = -box, box
        return lo, hi

    def _gamma_box(self, points: np.ndarray) -> np.ndarray:
        """A box that holds the optimal ``gamma`` of every point, so that the dual simplex never works
        with a free column. Below: ``lambda_t = 1/T, xi_t = x_star/T`` is feasible with every row within
        ``(row sum of |A_t| * max|x_star| + |b|, |x_star| + |l|, |h| + |x_star|) / T`` of zero. Above: every row
        bounds gamma by its own left-hand side, and ``xi_t <= x_star``, ``lambda_t <= 1``."""
        return self._d_scale + (1.0 + self._d_rowsum) * np.max(np.abs(points), axis=1, initial=0.0)

    def _answer(self, status: int, objective: float, y: np.ndarray, iterations: int, basis) -> CglpAnswer:
        n = self.n
        # the device LP is the dual: its infeasibility (1) is the CGLP's unboundedness (2) and vice versa
        code = {0: 0, 1: 2, 2: 1}.get(status, status)
        if code != 0:
            return CglpAnswer(None, None, code, float('nan'), iterations, None)
        pi = np.asarray(y[n:2 * n] - y[:n], dtype=float)
        pi0 = float(y[2 * n] - y[2 * n + 1])
        return CglpAnswer(pi, pi0, 0, -objective, iterations, basis)

    def _check_point(self, x_star) -> None:
        assert isinstance(x_star, CyLPArray), 'x_star must be a CyLPArray'
        assert x_star.shape == (self.n,), \
            'x_star must have the same number of variables as the LP relaxations ' \
            'in the branch and bound tree this instance was created with'

    def solve_batch(self, x_stars: Sequence[CyLPArray], starting_bases: Sequence = None) -> \
            List[Tuple[Union[CyLPArray, None], Union[float, None]]]:
        """``solve`` for many points at once: the cuts that separate each of
        ``x_stars`` most from the convex hull of the disjunctive terms, from one batched device call.
        ``starting_bases[k]`` is what ``starting_basis`` is for ``solve``. ``cglp.lp`` afterwards
        describes the LAST point's solve."""
        for x in x_stars:
            self._check_point(x)
        K = len(x_stars)
        if starting_bases is None:
            starting_bases = [self.lp._basis] * K
        assert len(starting_bases) == K, 'one starting basis (or None) per point'
        if not self.term_ids or not K:
            self.cylp_failure = self.cylp_failure or bool(K)
            return [(None, None)] * K
        answers = self._solve_points(np.array([np.asarray(x, dtype=float) for x in x_stars]).reshape(K, self.n),
                                     list(starting_bases))
        self._record(answers[-1])
        return [(CyLPArray(a.pi), a.pi0) if a.status == 0 else (None, None) for a in answers]

    def prefetch(self, x_stars: Sequence[np.ndarray], starting_bases: Sequence = None) -> int:
        """Solve the CGLP for the points several nodes are ABOUT to ask for (the LP solutions of a
        prefetched frontier) in one device call; ``solve(x_star)`` then finds its answer ready. Returns the
        number of points sent to the device."""
        if not self.term_ids:
            return 0
        todo, bases = {}, {}
        for k, x in enumerate(x_stars):
            x = np.ascontiguousarray(x, dtype=float).reshape(self.n)
            key = x.tobytes()
            if key not in self._prefetched and key not in todo:
                todo[key] = x
                bases[key] = None if starting_bases is None else starting_bases[k]
        if not todo:
            return 0
        if len(self._prefetched) > PREFETCH_CACHE_POINTS:
            self._prefetched.clear()
        answers = self._solve_points(np.array(list(todo.values())), [bases[k] for k in todo])
        self._prefetched.update(zip(todo, answers))
        return len(todo)

    def _record(self, a: CglpAnswer) -> None:
        self.lp._status = a.status
        self.lp.iteration = a.iterations
        if a.status == 0:
            self.lp.objectiveValue = a.objective
            if a.device_basis is not None:
                self.lp._basis = self._primal_coded(*a.device_basis)
        else:
            self.cylp_failure = True


    def solve(self, x_star: CyLPArray = None, starting_basis: Tuple[np.ndarray, np.ndarray] = None) -> \
            Tuple[Union[CyLPArray, None], Union[float, None]]:
        """The valid inequality that separates ``x_star`` (default: the root node's LP solution) most
        from the convex hull of the disjunctive terms, or (None, None) if the solve fails
        (reference :179-221). Without ``starting_basis`` the solve continues from the basis of the
        previous one, as a CyClpSimplex does."""
        if x_star is not None:
            self._check_point(x_star)
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
        x = np.ascontiguousarray(self._x_root if x_star is None else x_star, dtype=float)
        answer = self._prefetched.pop(x.tobytes(), None)
        if answer is None:
            answer = self._solve_points(x[None, :], [self.lp._basis])[0]
        else:
            self.prefetch_hits += 1
        self._record(answer)
        if answer.status == 0:
            return CyLPArray(answer.pi), answer.pi0
        return None, None
