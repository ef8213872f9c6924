"""TEST INFRASTRUCTURE — the four gurobipy calls the reference's tests make, answered by HiGHS.

The reference checks its B&B optimum on the 64 ``example_models`` against Gurobi
(``test_simple_mip_solver/helpers.py:30-53, 75-126``: ``gu.read(path)``, ``mdl.setParam(...)``,
``mdl.optimize()``, ``mdl.objVal``). gurobipy is not in this image; ``scipy.optimize.milp`` (HiGHS
branch and cut) on the oracle's own MPS reader gives the same optimum independently of every line
of the product. Only tests/tools/run_reference_tests_on_product.py and tests may import this.
"""
from __future__ import annotations

import types

import numpy as np
from scipy.optimize import Bounds, LinearConstraint, milp

from oracle.mps_py import read_mps


class _Model:
    def __init__(self, path: str):
        self._mdl = read_mps(path)
        self.objVal = None
        self.status = None

    def setParam(self, *args, **kwargs):
        pass

    def optimize(self):
        m = self._mdl
        sense = np.array(m.row_senses)
        lo = np.where(sense == 'L', -np.inf, m.rhs)
        hi = np.where(sense == 'G', np.inf, m.rhs)
        integrality = np.zeros(len(m.c))
        integrality[list(m.integer_indices)] = 1
        cons = [LinearConstraint(m.A, lo, hi)] if m.A.shape[0] else []
        res = milp(m.c, constraints=cons, integrality=integrality, bounds=Bounds(m.l, m.u))
        self.status = res.status
        self.objVal = float(res.fun) + m.obj_offset if res.status == 0 else None


def read(path: str) -> _Model:
    return _Model(path)


GRB = types.SimpleNamespace(Param=types.SimpleNamespace(LogToConsole='LogToConsole'))
