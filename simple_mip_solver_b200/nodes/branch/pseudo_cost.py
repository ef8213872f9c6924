"""PseudoCostBranchNode — pseudo-cost branching initialised by strong branching.

API and arithmetic of the reference's ``simple_mip_solver/nodes/branch/pseudo_cost.py``; the
strong-branching children of all uninitialised fractional variables are solved as ONE GPU batch
(``_strong_branch_batch``) instead of one ``lp.dual()`` per child (reference :57-62).
"""
from __future__ import annotations

from math import ceil, floor
from typing import Any, Dict, List, TypeVar, Union

from simple_mip_solver_b200.nodes.base_node import BaseNode
from simple_mip_solver_b200.utils.tolerance import variable_epsilon

T = TypeVar('T', bound='PseudoCostBranchNode')
pseudo_costs_hint = Dict[int, Dict[str, Dict[str, Union[float, int]]]]


class PseudoCostBranchNode(BaseNode):
    strong_branch_chunk = 256

    def __init__(self: T, *args: Any, **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.branch_method = 'pseudo cost'
        self.pseudo_costs = None
        self.strong_branch_iters = None

    def bound(self: T, pseudo_costs: pseudo_costs_hint, strong_branch_iters: int = 5,
              **kwargs: Any) -> Dict[str, Any]:
        problems = self._check_pseudo_costs(pseudo_costs)
        assert not problems, f'pseudo cost dict has following errors: {problems}'
        self.pseudo_costs = pseudo_costs
        self.strong_branch_iters = strong_branch_iters
        rtn = super().bound(**kwargs)
        if self.lp_feasible:
            self._update_pseudo_costs()
        rtn['pseudo_costs'] = self.pseudo_costs
        return rtn

    def _update_pseudo_costs(self: T) -> None:
        """Strong-branch every fractional integer variable that has no pseudo cost yet (one GPU
        batch), then update the cost of the variable this node was branched on (reference :46-66)."""
        sb_indices = [idx for idx in self._integer_indices
                      if self._is_fractional(float(self.solution[idx])) and idx not in self.pseudo_costs]
        if getattr(self._strong_branch, '__func__', None) is not BaseNode._strong_branch:
            # `_strong_branch` is the reference's plugin point for one variable (base_node.py:629-647):
            # a subclass or instance that replaces it is served by it, variable by variable (:60-62)
            pairs = ((idx, self._strong_branch(idx, self.strong_branch_iters)) for idx in sb_indices)
        else:
            pairs = self._strong_branch_chunks(sb_indices)
        for _, children in pairs:
            for child in children.values():
                self._calculate_costs(child)
        if self._b_idx is not None and self._b_idx not in sb_indices:
            self._calculate_costs(self)

    def _strong_branch_chunks(self: T, sb_indices: List[int]):
        """(variable, its two solved children) for all ``sb_indices``, one GPU call per chunk: up to 256
        candidates (512 child LPs) per call — wide enough for the kernels, and the child LP objects of a
        chunk (each holds its own bound vectors, as in the reference :592-608) are dropped before the next
        chunk is built, so a root with thousands of candidates stays small."""
        for first in range(0, len(sb_indices), self.strong_branch_chunk):
            chunk = sb_indices[first:first + self.strong_branch_chunk]
            children = self._strong_branch_batch(chunk, self.strong_branch_iters)
            for idx in chunk:
                yield idx, children[idx]

    def _calculate_costs(self: T, node: T) -> None:
        """Running mean of (objective change) / (variable change) for node's branching variable
        and direction; an infeasible child only counts a try (reference :68-100)."""
        idx, direction = node._b_idx, node._b_dir
        entry = self.pseudo_costs.setdefault(idx, {}).setdefault(direction, {'cost': 0, 'times': 0})
        if node.lp.getStatusCode() in [0, 3]:
            bound_change = max(node.lp.objectiveValue - node.dual_bound, 0)
            if direction == 'left':
                variable_change = node._b_val - node.lp.variablesUpper[idx]
            else:
                variable_change = node.lp.variablesLower[idx] - node._b_val
            entry['cost'] = float((entry['cost'] * entry['times'] + bound_change / variable_change)
                                  / (entry['times'] + 1))
        entry['times'] += 1

    def branch(self: T, pseudo_costs: pseudo_costs_hint, **kwargs: Any) -> Dict[str, T]:
        assert not self.mip_feasible, 'must have fractional value to branch'
        problems = self._check_pseudo_costs(pseudo_costs)
        assert not problems, f'pseudo cost dict has following errors: {problems}'
        return self._base_branch(self._best_pseudo_costs_index(pseudo_costs), **kwargs)

    def _best_pseudo_costs_index(self: T, pseudo_costs: pseudo_costs_hint) -> int:
        """argmax over fractional integer variables of min(up cost * up distance, down cost * down
        distance); ties resolved as a stable descending sort does (reference :118-133)."""
        scores = {}
        for i in self._integer_indices:
            v = float(self.solution[i])
            if self._is_fractional(v):
                scores[i] = min(pseudo_costs[i]['right']['cost'] * (ceil(v) - v),
                                pseudo_costs[i]['left']['cost'] * (v - floor(v)))
        return sorted(scores, key=scores.get, reverse=True)[0]

    def _check_pseudo_costs(self: T, pseudo_costs: pseudo_costs_hint) -> List[str]:
        problems = []
        for idx in pseudo_costs:
            if idx not in self._integer_indices:
                problems.append(f'index {idx} not integer index')
                continue
            for direction in ['right', 'left']:
                if direction not in pseudo_costs[idx]:
                    problems.append(f'index {idx} missing direction {direction}')
                    continue
                entry = pseudo_costs[idx][direction]
                if 'cost' not in entry:
                    problems.append(f'index {idx} direction {direction} missing cost')
                elif not (isinstance(entry['cost'], (int, float)) and entry['cost'] + variable_epsilon >= 0):
                    problems.append(f'index {idx} direction {direction} cost must be nonnegative number')
                if 'times' not in entry:
                    problems.append(f'index {idx} direction {direction} missing times')
                elif not (isinstance(entry['times'], int) and entry['times'] >= 0):
                    problems.append(f'index {idx} direction {direction} times must be nonnegative int')
        return problems
