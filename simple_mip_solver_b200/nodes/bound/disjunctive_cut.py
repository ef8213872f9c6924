"""DisjunctiveCutBoundNode — one disjunctive (CGLP) cut per cut round on top of BaseNode's loop.

Same keyword arguments, counters and branching rules as the reference's
``simple_mip_solver/nodes/bound/disjunctive_cut.py`` (:66-238). The node is a *caller* of the bound
step: every re-solve after a cut is the batched GPU LP solve, the cut itself comes from
``utils.cut_generating_lp.CutGeneratingLP``.
"""
from __future__ import annotations

import re
from typing import Any, Dict, Iterable, List, Tuple, Union

import numpy as np

from simple_mip_solver_b200.compat.cylp_like import CyLPArray
from simple_mip_solver_b200.nodes.base_node import BaseNode
from simple_mip_solver_b200.utils.cut_generating_lp import CutGeneratingLP
from simple_mip_solver_b200.utils.floating_point import numerically_safe_cut
from simple_mip_solver_b200.utils.tolerance import min_cglp_norm, variable_epsilon


class DisjunctiveCutBoundNode(BaseNode):
    """Adds to BranchAndBound the kwargs ``max_cglp_calls``, ``warm_start_cglp``,
    ``cglp_cumulative_constraints`` and ``cglp_cumulative_bounds`` (reference :13-31)."""

    def __init__(self, cglp: CutGeneratingLP = None, prev_cglp_basis: Tuple[np.ndarray, np.ndarray] = None,
                 force_create_cglp: bool = False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert isinstance(force_create_cglp, bool), 'force_create_cglp is bool'
        if cglp is not None:
            assert isinstance(cglp, CutGeneratingLP), 'cglp must be CutGeneratingLP instance'
        else:
            assert not force_create_cglp, 'cannot force creation of CGLP that does not exist'
        self.cglp = cglp
        self.prev_cglp_basis = prev_cglp_basis
        self.force_create_cglp = force_create_cglp
        self.current_node_added_cglp = force_create_cglp
        self.previous_cglp_added = cglp is not None
        self.cglp_name_pattern = re.compile('^cut_cglp_')
        self.current_cglp_name_pattern = re.compile(f'^cut_cglp_{self.idx}_')
        self.sharable_cuts = {}
        self.number_cglp_created = 0
        self.number_cglp_added = 0
        self.number_cglp_removed = 0

    @staticmethod
    def prefetch(nodes: Iterable[BaseNode]) -> int:
        """BaseNode.prefetch, then the FIRST disjunctive cut of every node of the batch that is going
        to ask for one: nodes that share a CGLP differ only in the point they want cut off, so their
        CGLPs are one batched device call (``CutGeneratingLP.prefetch``; the reference solves one
        CGLP per node and round, :133). A node's first cut round then finds its answer cached."""
        nodes = list(nodes)
        sent = BaseNode.prefetch(nodes)
        groups: Dict[int, Tuple[CutGeneratingLP, list, list]] = {}
        for node in nodes:
            point = node._first_cglp_point() if isinstance(node, DisjunctiveCutBoundNode) else None
            if point is not None:
                _, points, bases = groups.setdefault(id(node.cglp), (node.cglp, [], []))
                points.append(point)
                bases.append(node.prev_cglp_basis)
        for cglp, points, bases in groups.values():
            if len(points) > 1:
                cglp.prefetch(points, bases)
        return sent

    def _first_cglp_point(self) -> Union[None, np.ndarray]:
        """The point this node's first cut round will hand to its CGLP (``_cut_generation_iteration``
        clips the LP solution at zero first), or None if the node will not get that far: no CGLP, LP not
        solved to optimality, or an integral solution."""
        if self.cglp is None or not self.previous_cglp_added or self.cut_generation_iterations:
            return None
        lp = self.lp
        if getattr(lp, '_solved_key', None) is None or lp.getStatusCode() != 0:
            return None
        x = np.maximum(np.asarray(lp.primalVariableSolution['x'], dtype=float), 0)
        ints = x[self._integer_indices]
        if not len(ints) or np.max(np.abs(np.round(ints) - ints)) <= variable_epsilon:
            return None
        return x

    def bound(self, total_number_cglp_created: int = 0, total_number_cglp_added: int = 0,
              total_number_cglp_removed: int = 0, **kwargs: Any) -> Dict[str, Any]:
        assert isinstance(total_number_cglp_added, int) and total_number_cglp_added >= 0, \
            "total_number_cglp_added is nonnegative integer"
        assert isinstance(total_number_cglp_created, int) and total_number_cglp_created >= 0, \
            "total_number_gmic_created is nonnegative integer"
        assert isinstance(total_number_cglp_removed, int) and total_number_cglp_removed >= 0, \
            "total_number_cglp_removed is nonnegative integer"
        rtn = super().bound(**kwargs)
        rtn['total_number_cglp_created'] = total_number_cglp_created + self.number_cglp_created
        rtn['total_number_cglp_added'] = total_number_cglp_added + self.number_cglp_added
        rtn['total_number_cglp_removed'] = total_number_cglp_removed + self.number_cglp_removed
        if self.sharable_cuts:
            rtn['cuts'] = self.sharable_cuts        # BranchAndBound hands these to every queued node
        return rtn

    def _remove_slack_cuts(self, **kwargs) -> List[str]:
        removed = super()._remove_slack_cuts(**kwargs)
        self.number_cglp_removed += sum(1 for name in removed if self.cglp_name_pattern.match(name))
        return removed

    def _generate_cuts(self, max_cglp_calls: int = None, min_cglp_norm: float = min_cglp_norm,
                       **kwargs) -> Dict[str, Tuple[CyLPArray, float]]:
        """BaseNode's cuts plus, while the previous round's disjunctive cut was used and the call
        budget allows, one new CGLP cut separating the current solution (reference :109-140)."""
        if max_cglp_calls is not None:
            assert isinstance(max_cglp_calls, int) and max_cglp_calls >= 0, 'max_cglp_calls is a nonnegative integer'
        assert isinstance(min_cglp_norm, (float, int)) and min_cglp_norm > 0, 'min_cglp_norm is a positive number'
        budget = float('inf') if max_cglp_calls is None else max_cglp_calls
        pool = super()._generate_cuts(**kwargs)
        if self.previous_cglp_added and self.cut_generation_iterations <= budget:
            pi, pi0 = self.cglp.solve(x_star=CyLPArray(self.solution),
                                      starting_basis=self._get_cglp_starting_basis(**kwargs))
            if pi is not None and pi0 is not None and np.linalg.norm(pi) > min_cglp_norm:
                name = f'cut_cglp_{self.idx}_{self.cut_generation_iterations}'
                pool[name] = numerically_safe_cut(pi=pi, pi0=pi0, estimate='over')
                self.number_cglp_created += 1
        return pool

    def _get_cglp_starting_basis(self, warm_start_cglp: bool = True, **kwargs) -> \
            Union[None, Tuple[np.ndarray, np.ndarray]]:
        assert isinstance(warm_start_cglp, bool), 'warm_start_cglp is boolean'
        if not warm_start_cglp:
            return (np.array([3] * self.cglp.lp.nVariables, dtype=np.int32),
                    np.array([1] * self.cglp.lp.nConstraints, dtype=np.int32))
        if self.cut_generation_iterations == 1:
            return self.prev_cglp_basis
        return None

    def _select_cuts(self, cglp_cumulative_constraints: bool = True, cglp_cumulative_bounds: bool = True,
                     **kwargs) -> Dict[str, Tuple[CyLPArray, float]]:
        """Track whether this round's disjunctive cut was appended; a cut built from the original
        disjunction and feasible regions is valid everywhere and becomes sharable (reference :164-196)."""
        assert isinstance(cglp_cumulative_constraints, bool), 'cglp_cumulative_constraints is bool'
        assert isinstance(cglp_cumulative_bounds, bool), 'cglp_cumulative_bounds is bool'
        self.previous_cglp_added = self.force_create_cglp
        added = super()._select_cuts(**kwargs)
        for name, (pi, pi0) in added.items():
            if self.cglp_name_pattern.match(name):
                self.number_cglp_added += 1
                if self.current_cglp_name_pattern.match(name):
                    self.current_node_added_cglp = True
                    self.previous_cglp_added = True
                    if not cglp_cumulative_bounds and not cglp_cumulative_constraints:
                        self.sharable_cuts[name] = (pi, pi0)
        return added

    def branch(self, cglp_cumulative_constraints: bool = False, cglp_cumulative_bounds: bool = False,
               cglp: CutGeneratingLP = None, **kwargs: Any) -> Dict[str, Any]:
        """Children inherit a CGLP only if this node's own disjunctive cut was useful; with
        cumulative options the CGLP is rebuilt on this node's rows / bounds (reference :198-238)."""
        assert isinstance(cglp_cumulative_constraints, bool), 'cglp_cumulative_constraints is bool'
        assert isinstance(cglp_cumulative_bounds, bool), 'cglp_cumulative_bounds is bool'
        if self.cglp is None or not self.current_node_added_cglp:
            return super().branch(force_create_cglp=self.force_create_cglp, **kwargs)
        if cglp_cumulative_constraints or cglp_cumulative_bounds:
            A = self.lp.coefMatrix.copy() if cglp_cumulative_constraints else None
            b = CyLPArray(self.lp.constraintsLower.copy()) if cglp_cumulative_constraints else None
            var_lb = CyLPArray(self.lp.variablesLower.copy()) if cglp_cumulative_bounds else None
            var_ub = CyLPArray(self.lp.variablesUpper.copy()) if cglp_cumulative_bounds else None
            child_cglp = CutGeneratingLP(bb=self.cglp.bb, root_id=self.cglp.root_id, A=A, b=b,
                                         var_lb=var_lb, var_ub=var_ub)
            return super().branch(cglp=child_cglp, force_create_cglp=self.force_create_cglp, **kwargs)
        return super().branch(cglp=self.cglp, prev_cglp_basis=self.cglp.lp.getBasisStatus(),
                              force_create_cglp=self.force_create_cglp, **kwargs)
