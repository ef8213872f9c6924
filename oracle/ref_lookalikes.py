"""TEST INFRASTRUCTURE — the oracle's OWN minimal look-alikes of the CyLP modelling objects and of
gimpy's BinaryTree that the unmodified reference touches.

Deliberately independent of ``simple_mip_solver_b200.compat``: the goldens produced by running the
reference on these (tests/golden/make_goldens.py) must not inherit a bug of the product's
modelling layer. Only what the reference calls is implemented.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

COIN_INFINITY = 1.7976931348623157e308


class CyLPArray(np.ndarray):
    __array_priority__ = 5.0

    def __new__(cls, data, info=None):
        return np.array(data, dtype=np.float64).view(cls)


class CyLPVar:
    """``x = lp.addVariable('x', n)``; supports ``M * x``, ``l <= x <= u``."""
    __array_ufunc__ = None

    def __init__(self, name, dim):
        self.name, self.dim = name, int(dim)
        self.lower = CyLPArray(np.zeros(self.dim))
        self.upper = CyLPArray(np.full(self.dim, COIN_INFINITY))
        self.indices = np.arange(self.dim)
        self._half = None

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __rmul__(self, coefs):
        return CyLPExpr(self, coefs)

    __mul__ = __rmul__

    def __ge__(self, lower):            # first half of "l <= x <= u"
        self._half = np.array(np.broadcast_to(np.asarray(lower, float), (self.dim,)))
        return CyLPBounds(self, self._half, None)

    def __le__(self, upper):
        lo, self._half = self._half, None
        return CyLPBounds(self, lo, np.array(np.broadcast_to(np.asarray(upper, float), (self.dim,))))


class CyLPBounds:
    def __init__(self, var, lower, upper):
        self.var, self.lower, self.upper = var, lower, upper

    def __bool__(self):
        return True


class CyLPExpr:
    __array_ufunc__ = None

    def __init__(self, var, coefs):
        self.var = var
        M = sp.csr_matrix(coefs, dtype=float) if sp.issparse(coefs) else \
            sp.csr_matrix(np.atleast_2d(np.asarray(coefs, dtype=float)))
        assert M.shape[1] == var.dim
        self.coefs = M
        self._half = None

    def __ge__(self, lower):
        k = self.coefs.shape[0]
        self._half = np.array(np.broadcast_to(np.asarray(lower, float), (k,)))
        return CyLPConstraint(self, self._half, np.full(k, COIN_INFINITY))

    def __le__(self, upper):
        k = self.coefs.shape[0]
        lo, self._half = self._half, None
        lo = np.full(k, -COIN_INFINITY) if lo is None else lo
        return CyLPConstraint(self, lo, np.array(np.broadcast_to(np.asarray(upper, float), (k,))))


class CyLPConstraint:
    def __init__(self, expr, lower, upper, name=None):
        self.name = name
        self.lower, self.upper = CyLPArray(lower), CyLPArray(upper)
        self.variables = [expr.var]
        self.varCoefs = {expr.var: expr.coefs}
        self.nRows = expr.coefs.shape[0]
        self.isRange = False

    def __bool__(self):
        return True


class _Vertex:
    def __init__(self, **attr):
        self.attr = dict(attr)


class BinaryTree:
    """coinor.gimpy.tree.BinaryTree as branch_and_bound.py:19-108, 193, 285-286 uses it."""

    def __init__(self):
        self.nodes, self.root = {}, None
        self._kids, self._up = {}, {}

    def __contains__(self, name):
        return name in self.nodes

    def add_root(self, name, **attr):
        self.root = name
        self.nodes[name] = _Vertex(**attr)
        self._kids[name] = {}

    def _child(self, name, parent, side, attr):
        assert parent in self.nodes and name not in self.nodes and side not in self._kids[parent]
        self.nodes[name] = _Vertex(direction=side, **attr)
        self._kids[parent][side] = name
        self._kids[name] = {}
        self._up[name] = parent

    def add_left_child(self, name, parent, **attr):
        self._child(name, parent, 'left', attr)

    def add_right_child(self, name, parent, **attr):
        self._child(name, parent, 'right', attr)

    def get_children(self, name):
        return [self._kids[name][s] for s in ('left', 'right') if s in self._kids[name]]

    def get_left_child(self, name):
        return self._kids[name].get('left')

    def get_right_child(self, name):
        return self._kids[name].get('right')

    def get_parent(self, name):
        return self._up.get(name)

    def get_node_attr(self, name, attr):
        return self.nodes[name].attr.get(attr)

    def set_node_attr(self, name, attr, value):
        self.nodes[name].attr[attr] = value

    def get_node(self, name):
        return self.nodes[name]
