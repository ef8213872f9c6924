"""BaseAlgorithm — model normalisation, Node API check and the kwargs bus
(reference simple_mip_solver/algorithms/base_algorithm.py:15-72)."""
from __future__ import annotations

import inspect
from typing import Any, Dict, List, Type

import scipy.sparse as sp

from simple_mip_solver_b200.compat.milp_instance import MILPInstance
from simple_mip_solver_b200.nodes.base_node import BaseNode


class BaseAlgorithm:

    def __init__(self, model: MILPInstance, Node: Type[BaseNode], node_attributes: List[str],
                 node_funcs: List[str], **kwargs: Any):
        assert isinstance(model, MILPInstance), 'model must be cuppy MILPInstance'
        # the model kept here may be a rebuilt copy of the one passed in (reference :21-24)
        self.model = self._convert_constraints_to_greq(model)
        self._swapped_constraint_direction = model.sense != self.model.sense

        assert inspect.isclass(Node), 'Node must be a class'
        root_node = Node(lp=self.model.lp, integer_indices=self.model.integerIndices, idx=0, **kwargs)
        for attribute in node_attributes:
            assert hasattr(root_node, attribute), f'Node needs a {attribute} attribute'
        for func in node_funcs:
            assert callable(getattr(root_node, func, None)), f'Node needs a {func} function'
        assert 'next_node_idx' not in kwargs, 'key next_node_idx is reserved for use by solver'

        self._Node = Node
        self.root_node = root_node
        self.evaluated_nodes = 0
        kwargs['next_node_idx'] = 1
        self._kwargs = kwargs
        self._M = 999999999

    @staticmethod
    def _convert_constraints_to_greq(model: MILPInstance) -> MILPInstance:
        """``A x <= b`` becomes ``-A x >= -b``; the objective is already a minimisation."""
        if model.sense != '<=':
            return model
        A = model.A        # negated as it is: a sparse matrix stays sparse
        return MILPInstance(A=-A, b=-model.b, c=model.lp.objective, l=model.l, u=model.u,
                            integerIndices=model.integerIndices, sense=['Min', '>='],
                            numVars=len(model.c))

    def _process_rtn(self, rtn: Dict[str, Any]):
        assert isinstance(rtn, dict), 'rtn must be a dictionary'
        assert all(isinstance(k, str) for k in rtn), 'rtn keys must be strings'
        self._kwargs.update(rtn)
