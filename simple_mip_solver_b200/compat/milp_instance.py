"""``MILPInstance`` — the model container the reference takes as input.

The reference imports it from ``coinor.cuppy.milpInstance`` (third party, not in the repo) and
reads ``.A .b .c .l .u .sense .integerIndices .lp`` from it
(simple_mip_solver/algorithms/base_algorithm.py:18-29, 53-59; test_simple_mip_solver/helpers.py:42).
Same constructor keywords here; ``.lp`` is the GPU-backed LP look-alike.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

from .cylp_like import COIN_INFINITY, CyClpSimplex, CyLPArray, SharedLP
from .mps import read_mps


class MILPInstance:
    def __init__(self, A=None, b=None, c=None, l=None, u=None, sense: Optional[Sequence[str]] = None,
                 integerIndices: Optional[List[int]] = None, numVars: Optional[int] = None,
                 file_name: Optional[str] = None, device: int = 0):
        if file_name is not None:
            self._from_mps(file_name, device)
            return
        assert A is not None and b is not None and c is not None, 'A, b and c are required'
        sense = list(sense) if sense is not None else ['Min', '<=']
        assert sense[0] in ('Min', 'Max') and sense[1] in ('<=', '>='), "sense is e.g. ['Min', '>=']"
        self.A = A
        Am = sp.csr_matrix(A, dtype=float) if sp.issparse(A) else sp.csr_matrix(np.asarray(A, dtype=float))
        self.numCons, n = Am.shape
        self.numVars = int(numVars) if numVars is not None else n
        assert self.numVars == n, 'numVars must match the number of columns of A'
        self.b = CyLPArray(np.asarray(b, dtype=float).ravel())
        self.c = CyLPArray(np.asarray(c, dtype=float).ravel())
        self.l = CyLPArray(np.zeros(n)) if l is None else CyLPArray(np.asarray(l, dtype=float).ravel())
        self.u = CyLPArray(np.full(n, COIN_INFINITY)) if u is None else \
            CyLPArray(np.asarray(u, dtype=float).ravel())
        self.sense = sense[1]
        self.integerIndices = list(integerIndices) if integerIndices is not None else []
        # every model is turned into a minimisation through the objective held by .lp
        obj = self.c if sense[0] == 'Min' else -self.c
        self._build_lp(Am, obj, device)

    def _build_lp(self, Am, obj, device):
        # the LP object keeps rows in ">=" form; a "<=" model is flipped by BaseAlgorithm
        # (base_algorithm.py:47-61) before any node sees it
        if self.sense == '>=':
            # CyLP broadcasts a one-element right-hand side over the rows of `A * x >= b` (the reference's
            # `unbounded` example model passes one, example_models.py:155-163)
            rhs = np.asarray(self.b, dtype=float)
            rhs = np.full(Am.shape[0], rhs[0]) if rhs.size == 1 and Am.shape[0] > 1 else rhs
            shared = SharedLP(Am, rhs, np.asarray(obj), device=device)
            self.lp = CyClpSimplex(shared, self.l.copy(), self.u.copy())
        else:
            # a "<=" model, as cuppy builds it: the same CyClpSimplex protocol with rows bounded from above.
            # It is modelling-only (BaseNode rejects it with 'must have Ax >= b', base_node.py:111) until
            # BaseAlgorithm rebuilds the instance in ">=" form
            shared = SharedLP(sp.csr_matrix((0, Am.shape[1])), np.zeros(0), np.asarray(obj), device=device)
            self.lp = CyClpSimplex(shared, self.l.copy(), self.u.copy())
            rhs = np.asarray(self.b, dtype=float)
            rhs = np.full(Am.shape[0], rhs[0]) if rhs.size == 1 and Am.shape[0] > 1 else rhs
            self.lp._base_off = True
            self.lp._foreign_base = (Am, np.full(Am.shape[0], -COIN_INFINITY), rhs)
        self.lp._objective = CyLPArray(obj)

    def _from_mps(self, file_name, device):
        mdl = read_mps(file_name)
        senses = set(mdl.row_senses)
        assert senses <= {'L'} or senses <= {'G'}, 'all rows must have the same sense'
        self.sense = '>=' if senses == {'G'} else '<='
        # the matrix stays sparse (the reference's cuppy hands out a csc_matrixPlus for MPS input,
        # base_algorithm.py:55-56): a 50 000 x 20 000 model is 8 GB dense
        self.A = mdl.A.tocsc()
        self.numCons, self.numVars = mdl.A.shape
        self.b = CyLPArray(mdl.rhs)
        self.c = CyLPArray(mdl.c)
        self.l = CyLPArray(mdl.l)
        self.u = CyLPArray(np.where(np.isinf(mdl.u), COIN_INFINITY, mdl.u))
        self.integerIndices = list(mdl.integer_indices)
        self._build_lp(mdl.A, self.c, device)
