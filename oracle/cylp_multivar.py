"""TEST INFRASTRUCTURE — CyLP's multi-variable modelling algebra, as far as the reference's CGLP uses it.

``simple_mip_solver/utils/cut_generating_lp.py:52-221`` builds its LP from SEVERAL named variables
(``pi, pi0, u_t, w_t, v_t``) with expressions such as ``0 >= -pi + A.T * u + I * w - I * v``,
``sum(var.sum() ...) == 1`` and ``lp.objective = x_star * pi - pi0``, then calls ``lp.primal()``.
The single-variable look-alikes of oracle/ref_lookalikes.py (all the Node classes need) do not cover
that; this module does, with HiGHS answering ``primal()``, so that the UNMODIFIED reference CGLP runs
here and its optima become golden values (tests/golden/make_cglp_goldens.py). Independent of the
product's modelling layer; only tests/ and tests/golden/ import it.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import scipy.sparse as sp

from oracle.highs_lp import HIGHS_INF, HighsLP
from oracle.ref_lookalikes import COIN_INFINITY, CyLPArray


class MVar:
    __array_ufunc__ = None

    def __init__(self, name, dim):
        self.name, self.dim = name, int(dim)
        # CyLP variables are free until bounded (the reference bounds u, w, v and leaves pi, pi0 free, :153-161;
        # its test checks pi.lower == -inf, test_cut_generating_lp.py:103-108)
        self.lower = CyLPArray(np.full(self.dim, -COIN_INFINITY))
        self.upper = CyLPArray(np.full(self.dim, COIN_INFINITY))
        self.indices = np.arange(self.dim)
        self._half = None

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def _expr(self):
        return MExpr({self: sp.identity(self.dim, format='csr')})

    def __rmul__(self, coefs):
        M = sp.csr_matrix(coefs, dtype=float) if sp.issparse(coefs) else \
            sp.csr_matrix(np.atleast_2d(np.asarray(coefs, dtype=float)))
        assert M.shape[1] == self.dim, (M.shape, self.dim)
        return MExpr({self: M})

    __mul__ = __rmul__

    def __neg__(self):
        return MExpr({self: -sp.identity(self.dim, format='csr')})

    def __add__(self, other):
        return self._expr() + other

    __radd__ = __add__

    def __sub__(self, other):
        return self._expr() - other

    def __rsub__(self, other):
        return (-self) + other

    def sum(self):
        return MExpr({self: sp.csr_matrix(np.ones((1, self.dim)))})

    def __ge__(self, lower):            # "var >= l" and the first half of "l <= var <= u"
        self._half = np.array(np.broadcast_to(np.asarray(lower, float), (self.dim,)))
        return MBounds(self, self._half, None)

    def __le__(self, upper):
        lo, self._half = self._half, None
        return MBounds(self, lo, np.array(np.broadcast_to(np.asarray(upper, float), (self.dim,))))


class MBounds:
    def __init__(self, var, lower, upper):
        self.var, self.lower, self.upper = var, lower, upper

    def __bool__(self):
        return True


class MExpr:
    """Sum over variables of (matrix * variable); all blocks have the same number of rows."""
    __array_ufunc__ = None

    def __init__(self, terms: Dict[MVar, sp.csr_matrix]):
        self.terms = terms
        self.rows = next(iter(terms.values())).shape[0]
        assert all(M.shape[0] == self.rows for M in terms.values())
        self._half = None

    @staticmethod
    def _of(other):
        return other._expr() if isinstance(other, MVar) else other

    def __add__(self, other):
        if isinstance(other, (int, float)) and other == 0:          # the start value of sum()
            return self
        other = self._of(other)
        assert isinstance(other, MExpr) and other.rows == self.rows
        out = dict(self.terms)
        for v, M in other.terms.items():
            out[v] = out[v] + M if v in out else M
        return MExpr(out)

    __radd__ = __add__

    def __neg__(self):
        return MExpr({v: -M for v, M in self.terms.items()})

    def __sub__(self, other):
        return self + (-self._of(other))

    def __ge__(self, lower):
        self._half = np.array(np.broadcast_to(np.asarray(lower, float), (self.rows,)))
        return MConstraint(self, self._half, np.full(self.rows, COIN_INFINITY))

    def __le__(self, upper):
        lo, self._half = self._half, None
        lo = np.full(self.rows, -COIN_INFINITY) if lo is None else lo
        return MConstraint(self, lo, np.array(np.broadcast_to(np.asarray(upper, float), (self.rows,))))

    def __eq__(self, value):
        rhs = np.array(np.broadcast_to(np.asarray(value, float), (self.rows,)))
        return MConstraint(self, rhs, rhs.copy())

    __hash__ = None


class MConstraint:
    def __init__(self, expr: MExpr, lower, upper, name=None):
        self.name = name
        self.lower, self.upper = CyLPArray(lower), CyLPArray(upper)
        self.variables = list(expr.terms)
        self.varCoefs = dict(expr.terms)
        self.nRows = expr.rows

    def __bool__(self):
        return True


class MultiVarSimplex:
    """The calls cut_generating_lp.py makes on its CyClpSimplex, answered by HiGHS."""

    def __init__(self):
        self.variables: List[MVar] = []
        self.constraints: List[MConstraint] = []
        self._objective = None
        self.logLevel = 0
        self.iteration = 0
        self._status = -1
        self._solution = None
        self._basis = None
        self.objectiveValue = 0.0

    def addVariable(self, name, dim, isInt=False):
        v = MVar(name, dim)
        self.variables.append(v)
        return v

    def getVarByName(self, name):
        return next(v for v in self.variables if v.name == name)

    def __iadd__(self, stmt):
        if isinstance(stmt, MBounds):
            if stmt.lower is not None:
                stmt.var.lower = CyLPArray(stmt.lower)
            if stmt.upper is not None:
                stmt.var.upper = CyLPArray(stmt.upper)
        else:
            self.addConstraint(stmt)
        return self

    def addConstraint(self, cons, name=None, addMpsNames=True):
        cons.name = name
        self.constraints.append(cons)

    @property
    def nVariables(self):
        return sum(v.dim for v in self.variables)

    @property
    def nConstraints(self):
        return sum(c.nRows for c in self.constraints)

    @property
    def objective(self):
        return self._objective

    @objective.setter
    def objective(self, expr):
        self._objective = self._row(MExpr._of(expr))

    def _offsets(self):
        off, at = {}, 0
        for v in self.variables:
            off[v] = at
            at += v.dim
        return off, at

    def _row(self, expr: MExpr):
        off, n = self._offsets()
        out = np.zeros((expr.rows, n))
        for v, M in expr.terms.items():
            out[:, off[v]:off[v] + v.dim] += M.toarray()
        return CyLPArray(out[0]) if expr.rows == 1 else out

    def getStatusCode(self):
        return self._status

    def getBasisStatus(self):
        if self._basis is None:
            return np.full(self.nVariables, 3, dtype=np.int32), np.full(self.nConstraints, 1, dtype=np.int32)
        return self._basis[0].copy(), self._basis[1].copy()

    def setBasisStatus(self, cols, rows):
        self._basis = (np.asarray(cols, dtype=np.int32).copy(), np.asarray(rows, dtype=np.int32).copy())

    def primal(self):
        off, n = self._offsets()
        blocks = []
        for c in self.constraints:
            row = sp.lil_matrix((c.nRows, n))
            for v, M in c.varCoefs.items():
                row[:, off[v]:off[v] + v.dim] = M
            blocks.append(row.tocsr())
        A = sp.vstack(blocks, format='csr')
        fin = lambda a, big: np.where(np.abs(a) >= 1e30, big, a)
        lo = fin(np.concatenate([c.lower for c in self.constraints]), -HIGHS_INF)
        up = fin(np.concatenate([c.upper for c in self.constraints]), HIGHS_INF)
        l = fin(np.concatenate([v.lower for v in self.variables]), -HIGHS_INF)
        u = fin(np.concatenate([v.upper for v in self.variables]), HIGHS_INF)
        lp = HighsLP(A, np.asarray(self._objective, dtype=float), lo, up, l, u)
        r = lp.solve()
        self._status = r.status
        self.iteration = r.iterations
        if r.status == 0:
            self.objectiveValue = r.objective
            self._solution = r.x
            self._basis = (np.asarray(r.col_basis, dtype=np.int32), np.asarray(r.row_basis, dtype=np.int32))
        return r.status

    dual = primal

    @property
    def primalVariableSolution(self):
        off, _ = self._offsets()
        return {v.name: CyLPArray(self._solution[off[v]:off[v] + v.dim]) for v in self.variables}
