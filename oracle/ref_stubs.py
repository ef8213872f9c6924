"""TEST INFRASTRUCTURE — run the UNMODIFIED reference package on a HiGHS stand-in for CyLP/CLP.

The reference (`/root/reference/simple_mip_solver`) imports cylp, coinor.cuppy, coinor.gimpy and
coinor.grumpy, none of which exist in this image. ``install()`` registers
look-alike modules under those names whose LP arithmetic is HiGHS dual simplex
(oracle/highs_lp.py), after which ``import simple_mip_solver`` works in THIS container and the
reference's own control flow (BranchAndBound, BaseNode, PseudoCostBranchNode, ...) can be executed
to produce golden vectors (tests/golden/make_goldens.py). Nothing here is shipped or measured as
the product, and nothing on the GPU box needs it: the goldens are committed.

Only what the reference touches is implemented (the CyClpSimplex protocol subset of SURVEY.md
section 8b, cuppy's MILPInstance constructor, gimpy's BinaryTree calls, GrUMPy's generator).
"""
from __future__ import annotations

import random as _pyrandom
import sys
import types
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp

from oracle.dual_simplex import dual_simplex
from oracle.highs_lp import HIGHS_INF, HighsLP
from oracle.mps_py import read_mps
from oracle.ref_lookalikes import BinaryTree as _Tree
from oracle.ref_lookalikes import (COIN_INFINITY, CyLPArray, CyLPBounds, CyLPConstraint, CyLPExpr,
                                   CyLPVar)

WARM_START = True      # False: every solve is cold, a deterministic function of the LP data
# which exact simplex answers lp.dual(): 'highs' (HiGHS 1.12 dual simplex) or 'dual_simplex' (the
# textbook dual simplex of oracle/dual_simplex.py, the restatement of what the device runs)
LP_BACKEND = 'highs'


class _DenseCoefConstraint(CyLPConstraint):
    """Constraint whose coefficient block is a dense np.matrix, as CyLP keeps for dense models."""

    def __init__(self, cons: CyLPConstraint, name):
        self.name = name
        self.lower, self.upper = cons.lower, cons.upper
        self.variables = cons.variables
        var = cons.variables[0]
        self.varCoefs = {var: np.asmatrix(cons.varCoefs[var].toarray())}
        self.nRows = cons.nRows
        self.isRange = False


class CyClpSimplex:
    """HiGHS-backed subset of cylp.cy.CyClpSimplex."""

    def __init__(self):
        self._vars: List[CyLPVar] = []
        self._l = self._u = None
        self.constraints: List[_DenseCoefConstraint] = []
        self._objective = None
        self.logLevel = 0
        self.maxNumIteration = 2147483647
        self.iteration = 0
        self._status = -1
        self._obj = 0.0
        self._x = self._y = self._rc = None
        self._basis = None
        self._auto = 0

    # -- modelling
    def addVariable(self, name, dim, isInt=False):
        v = CyLPVar(name, dim)
        self._vars.append(v)
        self._l = CyLPArray(np.zeros(dim))
        self._u = CyLPArray(np.full(dim, COIN_INFINITY))
        return v

    def getVarByName(self, name):
        return next(v for v in self._vars if v.name == name)

    @property
    def variables(self):
        v = self._vars[0]
        v.lower, v.upper = self._l, self._u
        return self._vars

    def __iadd__(self, stmt):
        if isinstance(stmt, CyLPBounds):
            if stmt.lower is not None:
                self._l = CyLPArray(stmt.lower)
            if stmt.upper is not None:
                self._u = CyLPArray(stmt.upper)
        else:
            self.addConstraint(stmt)
        return self

    def addConstraint(self, cons, name=None, addMpsNames=True):
        if name is None:
            name = f'R_{self._auto}'
            self._auto += 1
        self.constraints.append(_DenseCoefConstraint(cons, name))

    def removeConstraint(self, name):
        k = next(i for i, c in enumerate(self.constraints) if c.name == name)
        del self.constraints[k]
        self._basis = None

    @property
    def objective(self):
        return self._objective

    @objective.setter
    def objective(self, value):
        if isinstance(value, CyLPExpr):
            value = np.asarray(value.coefs.todense()).ravel()
        self._objective = CyLPArray(np.asarray(value, dtype=float).ravel())

    @property
    def objectiveCoefficients(self):
        return self._objective

    # -- array views
    @property
    def nVariables(self):
        return self._vars[0].dim

    nCols = nVariables

    @property
    def nConstraints(self):
        return sum(c.nRows for c in self.constraints)

    nRows = nConstraints

    @property
    def variablesLower(self):
        return self._l

    @variablesLower.setter
    def variablesLower(self, v):
        self._l = CyLPArray(v)

    @property
    def variablesUpper(self):
        return self._u

    @variablesUpper.setter
    def variablesUpper(self, v):
        self._u = CyLPArray(v)

    @property
    def constraintsLower(self):
        return CyLPArray(np.concatenate([np.asarray(c.lower) for c in self.constraints]))

    @property
    def constraintsUpper(self):
        return CyLPArray(np.concatenate([np.asarray(c.upper) for c in self.constraints]))

    @property
    def coefMatrix(self):
        x = self._vars[0]
        return sp.csc_matrix(np.vstack([np.asarray(c.varCoefs[x]) for c in self.constraints]))

    @property
    def matrix(self):
        return types.SimpleNamespace(elements=self.coefMatrix.tocsc().data)

    @staticmethod
    def getCoinInfinity():
        return COIN_INFINITY

    def setInteger(self, idx):
        pass

    # -- solving
    def _solve_ds(self):
        n = self.nVariables
        c = np.zeros(n) if self._objective is None else np.asarray(self._objective)
        A = self.coefMatrix.toarray()
        assert (np.asarray(self.constraintsUpper) >= 1e300).all(), 'rows must read a.x >= b'
        cs = rs = None
        if WARM_START and self._basis is not None and len(self._basis[0]) == n \
                and len(self._basis[1]) == A.shape[0]:
            cs, rs = self._basis
        r = dual_simplex(A, np.asarray(self.constraintsLower), c, np.asarray(self._l), np.asarray(self._u),
                         col_status=cs, row_status=rs, max_pivots=self.maxNumIteration)
        self._status, self.iteration = r.status, r.pivots
        self._obj = r.objective if r.status != 1 else float('inf')
        self._x = None if r.status == 1 else CyLPArray(r.x)
        self._y, self._rc = r.y, r.rc
        self._basis = (r.col_status.astype(np.int32), r.row_status.astype(np.int32))
        return r.status

    def _solve(self):
        if LP_BACKEND == 'dual_simplex':
            return self._solve_ds()
        n = self.nVariables
        c = np.zeros(n) if self._objective is None else np.asarray(self._objective)
        A = self.coefMatrix
        h = HighsLP(A, c, self.constraintsLower, self.constraintsUpper, self._l, self._u)
        if WARM_START and self._basis is not None and len(self._basis[0]) == n \
                and len(self._basis[1]) == A.shape[0]:
            h.set_basis(*self._basis)
        r = h.solve(self.maxNumIteration if self.maxNumIteration < 2147483647 else None)
        self._status = r.status
        self.iteration = r.iterations
        self._obj = r.objective
        self._x = None if r.x is None else CyLPArray(r.x)
        self._y = r.row_dual
        self._rc = r.reduced_cost
        if r.col_basis is not None:
            self._basis = (r.col_basis, r.row_basis)
        return r.status

    def dual(self, *args, **kwargs):
        return self._solve()

    def primal(self, *args, **kwargs):
        return self._solve()

    def getStatusCode(self):
        return self._status

    @property
    def objectiveValue(self):
        return self._obj

    @property
    def primalVariableSolution(self):
        return {'x': self._x}

    @property
    def dualVariableSolution(self):
        return {'x': self._rc}

    @property
    def dualConstraintSolution(self):
        out, k = {}, 0
        for c in self.constraints:
            out[c.name] = CyLPArray(self._y[k:k + c.nRows]) if self._y is not None else None
            k += c.nRows
        return out

    def getBasisStatus(self):
        if self._basis is None:
            return (np.full(self.nVariables, 3, dtype=np.int32),
                    np.full(self.nConstraints, 1, dtype=np.int32))
        return self._basis[0].copy(), self._basis[1].copy()

    def setBasisStatus(self, cols, rows):
        self._basis = (np.asarray(cols, dtype=np.int32).copy(), np.asarray(rows, dtype=np.int32).copy())


class csc_matrixPlus(sp.csc_matrix):
    pass


class MILPInstance:
    """coinor.cuppy.milpInstance.MILPInstance as the reference uses it."""

    def __init__(self, A=None, b=None, c=None, l=None, u=None, sense=None, integerIndices=None,
                 numVars=None, file_name=None):
        if file_name is not None:
            mdl = read_mps(file_name)
            senses = set(mdl.row_senses)
            A = csc_matrixPlus(mdl.A)
            b, c, l = CyLPArray(mdl.rhs), CyLPArray(mdl.c), CyLPArray(mdl.l)
            u = CyLPArray(np.where(np.isinf(mdl.u), COIN_INFINITY, mdl.u))
            sense = ['Min', '>=' if senses == {'G'} else '<=']
            integerIndices = list(mdl.integer_indices)
            numVars = len(mdl.c)
        self.A, self.b, self.c = A, b, c
        self.numVars = numVars if numVars is not None else np.shape(A)[1]
        self.l = l if l is not None else CyLPArray(np.zeros(self.numVars))
        self.u = u if u is not None else CyLPArray(np.full(self.numVars, COIN_INFINITY))
        self.sense = sense[1]
        self.integerIndices = integerIndices or []
        self.lp = CyClpSimplex()
        x = self.lp.addVariable('x', self.numVars)
        dense = A.toarray() if sp.issparse(A) else np.asarray(A, dtype=float)
        self.lp += (dense * x >= np.asarray(b, dtype=float).ravel()) if self.sense == '>=' \
            else (dense * x <= np.asarray(b, dtype=float).ravel())
        self.lp += np.asarray(self.l, dtype=float) <= x <= np.asarray(self.u, dtype=float)
        cvec = np.asarray(c, dtype=float).ravel()
        self.lp.objective = cvec if sense[0] == 'Min' else -cvec


def GenerateRandomMIP(numVars=40, numCons=20, density=0.2, maxObjCoeff=10, maxConsCoeff=10,
                      tightness=2, rand_seed=2, layout='dot'):
    """GrUMPy's generator (draw order restated in SURVEY.md section 8c; it reproduces the
    reference's checked-in MPS fixtures exactly — tests/test_instances.py)."""
    rng = _pyrandom.Random()
    rng.seed(rand_seed)
    cons = ['C' + str(i) for i in range(numCons)]
    variables = ['x' + str(i) for i in range(numVars)]
    obj = {v: rng.randint(1, maxObjCoeff) for v in variables}
    mat = {v: [rng.randint(1, maxConsCoeff) if rng.random() <= density else 0 for _ in cons]
           for v in variables}
    lo = int(numVars * density * maxConsCoeff / tightness)
    hi = int(numVars * density * maxConsCoeff / 1.5)
    rhs = [rng.randint(lo, hi) for _ in cons]
    return cons, variables, obj, mat, rhs


def install(reference_root: str = '/root/reference'):
    """Register the stand-in modules and put the reference on sys.path."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    this = sys.modules[__name__]
    mod('cylp')
    cy = mod('cylp.cy', CyClpSimplex=CyClpSimplex)
    mod('cylp.cy.CyClpSimplex', CyClpSimplex=CyClpSimplex, CyLPArray=CyLPArray)
    mod('cylp.py')
    mod('cylp.py.modeling')
    mod('cylp.py.modeling.CyLPModel', CyLPArray=CyLPArray)
    mod('cylp.py.utils')
    mod('cylp.py.utils.sparseUtil', csc_matrixPlus=csc_matrixPlus)
    mod('coinor')
    mod('coinor.cuppy')
    mod('coinor.cuppy.milpInstance', MILPInstance=MILPInstance)
    mod('coinor.gimpy')
    mod('coinor.gimpy.tree', BinaryTree=_Tree)
    mod('coinor.grumpy')
    mod('coinor.grumpy.BranchAndBound', GenerateRandomMIP=GenerateRandomMIP)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return this


def enable_cglp():
    """Give the reference's cut generating LP (simple_mip_solver/utils/cut_generating_lp.py) the multi-variable
    modelling algebra it builds its LP with (oracle/cylp_multivar.py). Call after ``install()``; returns the
    reference module. Every other reference module keeps the single-variable look-alike above."""
    import simple_mip_solver.utils.cut_generating_lp as ref_cglp
    from oracle.cylp_multivar import MultiVarSimplex
    ref_cglp.CyClpSimplex = MultiVarSimplex
    return ref_cglp

