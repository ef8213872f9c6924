"""BranchAndBound — the search loop, with a batched LP frontier.

Interface and decisions follow the reference's ``simple_mip_solver/algorithms/branch_and_bound.py``
(constructor :123-213, ``solve`` :215-241, ``_evaluate_node`` :243-266, return processing
:268-306): nodes are taken from the queue one at a time, pruned against the incumbent, bounded,
and either update the incumbent or branch. That order-dependent control flow stays in Python.

New here: before a node is bounded, the LP relaxations of the ``frontier_batch`` best open nodes
are solved together in one GPU call (``BaseNode.prefetch``) and cached in their LP objects; the
loop then proceeds exactly as the sequential one, most ``lp.dual()`` calls being cache hits. The
prefetch is speculative — a prefetched node may later be pruned — and never changes which node is
evaluated next or what any node's bound is, so the tree is the one the sequential loop builds.
"""
from __future__ import annotations

import heapq
import time
from queue import PriorityQueue
from typing import Any, Dict, Iterable, List, Tuple, Type, Union

import numpy as np

from simple_mip_solver_b200.algorithms.base_algorithm import BaseAlgorithm
from simple_mip_solver_b200.compat.binary_tree import BinaryTree
from simple_mip_solver_b200.compat.milp_instance import MILPInstance
from simple_mip_solver_b200.nodes.base_node import BaseNode


class BranchAndBoundTree(BinaryTree):
    """Tree of evaluated and open nodes; vertex attribute 'node' holds the Node instance."""

    def get_leaves(self, subtree_root_id: int, depth: int = None, keep: str = 'all') -> List[BaseNode]:
        """Leaves of the subtree under ``subtree_root_id``, optionally after cutting the subtree
        ``depth`` edges below its root (reference :22-61)."""
        assert subtree_root_id in self, 'subtree_root_id must belong to the tree'
        assert keep in ['all', 'feasible', 'not infeasible'], \
            "keep is one of 'all', 'feasible', or 'not infeasible'"
        everyone = [v.attr['node'] for v in self.nodes.values()]
        if depth is None:
            found = [n for n in everyone if n.is_leaf and subtree_root_id in n.lineage]
        else:
            assert isinstance(depth, int) and depth >= 0, 'depth is a nonnegative integer'
            if depth == 0:
                found = self.get_node_instances([subtree_root_id])
            elif depth == 1:
                found = self.get_node_instances(self.get_children(subtree_root_id))
            else:
                shallow = [n for n in everyone if n.is_leaf and subtree_root_id in n.lineage[-depth:]]
                at_depth = [n for n in everyone if len(n.lineage) >= depth + 1 and
                            n.lineage[-(depth + 1)] == subtree_root_id]
                found = shallow + at_depth
        if keep == 'feasible':
            return [n for n in found if n.lp_feasible]
        if keep == 'not infeasible':
            return [n for n in found if n.lp_feasible is not False]
        return found

    def get_disjunction(self, subtree_root_id: int) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
        return {n.idx: (n.lp.variablesLower.copy(), n.lp.variablesUpper.copy())
                for n in self.get_leaves(subtree_root_id, keep='not infeasible')}

    def get_node_instances(self, node_ids: Union[int, Iterable[int]]):
        single = isinstance(node_ids, int)
        if single:
            node_ids = [node_ids]
        else:
            assert isinstance(node_ids, Iterable) and not isinstance(node_ids, str), \
                'node_ids must be an integer or iterable (that is not a string)'
            node_ids = list(node_ids)
        missing = set(node_ids) - set(self.nodes)
        assert not missing, f'the following node_ids are not in the tree: {missing}'
        found = [self.nodes[i].attr.get('node') for i in node_ids]
        assert all(n is not None for n in found), \
            'each vertex in the branch and bound tree must have an attribute for a node instance'
        return found[0] if single else found

    def subtree_dual_bound(self, subtree_root_id: int, depth: int = None) -> Union[float, int]:
        assert subtree_root_id in self, 'subtree_root_id must belong to the tree'
        return min(n.objective_value if n.objective_value is not None else n.dual_bound
                   for n in self.get_leaves(subtree_root_id, depth=depth))


class BranchAndBound(BaseAlgorithm):
    """Solve a MILP by branch and bound with the search/branch/bound methods of ``Node``."""

    _node_attributes = ['dual_bound', 'objective_value', 'solution', 'lp_feasible', 'mip_feasible',
                        'search_method', 'branch_method', 'idx', 'lp', 'is_leaf', 'lineage']
    _node_funcs = ['bound', 'branch', '__lt__', '__eq__']
    _queue_funcs = ['put', 'get', 'empty']

    def __init__(self, model: MILPInstance, Node: Type[BaseNode] = BaseNode, node_queue: Any = None,
                 node_limit: int = float('inf'), mip_gap: float = .0001, logging: bool = False,
                 max_run_time: float = float('inf'), initial_primal_bound: float = float('inf'),
                 frontier_batch: int = 32, lp_cutoff: bool = False, **kwargs: Any):
        """``frontier_batch``: how many open nodes have their LP relaxation solved per GPU call
        (1 = one LP per call, as the reference). ``lp_cutoff``: hand the incumbent's value to the LP solver as
        an objective limit, so that a node whose dual bound already reaches it stops iterating (the
        reference solves it to the end and prunes it then, :251, :261; the tree is the same, the node's
        recorded LP value is a bound). All other arguments as in the reference."""
        node_queue = node_queue or PriorityQueue()
        super().__init__(model=model, Node=Node, node_attributes=self._node_attributes,
                         node_funcs=self._node_funcs, **kwargs)
        for func in self._queue_funcs:
            assert callable(getattr(node_queue, func, None)), f'node_queue needs a {func} function'
        assert node_limit == float('inf') or (isinstance(node_limit, int) and node_limit > 0), \
            "node limit must be positive integer or infinity"
        assert 0 <= mip_gap < 1, 'mip_gap is a ratio between 0 and 1'
        assert isinstance(logging, bool), 'logging is boolean'
        assert max_run_time > 0, 'max_run_time is positive value'
        assert initial_primal_bound > -float('inf'), 'initial_primal_bound is real or infinite'
        assert isinstance(frontier_batch, int) and frontier_batch >= 1, 'frontier_batch is a positive integer'
        special_keys = {'right', 'left', 'cuts'}
        assert set(kwargs.keys()).isdisjoint(special_keys), f'keys {special_keys} are saved for later use'
        assert all(isinstance(k, str) for k in kwargs), 'kwargs keys must be strings'

        self._node_queue = node_queue
        self._unbounded = None
        self._best_solution = None
        self.solution = None
        self.status = 'unsolved'
        self.objective_value = None
        self.primal_bound = initial_primal_bound
        self.node_limit = node_limit
        self.tree = BranchAndBoundTree()
        self.tree.add_root(self.root_node.idx, node=self.root_node)
        self.solve_time = 0
        self.mip_gap = mip_gap
        self.logging = logging
        self.max_run_time = max_run_time
        self.frontier_batch = frontier_batch
        self.lp_cutoff = bool(lp_cutoff)
        self.prefetch_calls = 0
        self.prefetched_lps = 0
        self.unsolved_nodes = 0         # nodes whose LP stopped on the solver's iteration budget
        self.frontier_bounds = (None, None)

    @property
    def dual_bound(self):
        return self.tree.subtree_dual_bound(self.root_node.idx)

    @property
    def current_gap(self):
        if self.primal_bound == self.dual_bound == 0:
            return 0
        if self.primal_bound == 0:
            return float('inf')
        if self.primal_bound == float('inf'):
            return None
        return abs(self.primal_bound - self.dual_bound) / abs(self.primal_bound)

    def solve(self) -> None:
        start = time.process_time()
        if self.status == 'unsolved':
            self._node_queue.put(self.root_node)

        def finished():
            gap = self.current_gap
            return (self._node_queue.empty() or self._unbounded or
                    self.evaluated_nodes >= self.node_limit or
                    (gap is not None and gap <= self.mip_gap) or
                    time.process_time() - start > self.max_run_time)

        while not finished():
            if self.logging and self.evaluated_nodes % 100 == 0:
                print(f'{self.evaluated_nodes} nodes evaluated gap: {self.current_gap}')
            node = self._node_queue.get()
            if self.lp_cutoff and self.primal_bound < float('inf') and hasattr(node.lp, 'solver_opts'):
                node.lp.solver_opts['obj_cutoff'] = float(self.primal_bound)     # shared by the model's LPs
            self._prefetch_frontier(node)
            self._evaluate_node(node)

        self.solve_time += time.process_time() - start
        if self._unbounded:
            self.status = 'unbounded'
        elif self._node_queue.empty() and self.primal_bound == float('inf') and not self.unsolved_nodes:
            self.status = 'infeasible'
        elif self.primal_bound < float('inf') and self.current_gap is not None and self.current_gap <= self.mip_gap:
            self.status = 'optimal'
        else:
            self.status = 'stopped on iterations or time'
        self.solution = self._best_solution
        self.objective_value = self.primal_bound

    def _prefetch_frontier(self, node: BaseNode) -> None:
        """If ``node``'s LP is not solved yet, solve it together with the LPs of the best open
        nodes still in the queue (those that would not be pruned right now) in one GPU call."""
        if self.frontier_batch <= 1 or node.lp._solved_key is not None:
            return
        if not node.dual_bound < self.primal_bound:
            return
        pending = getattr(self._node_queue, 'queue', None)
        batch = [node]
        if pending:
            want = self.frontier_batch - 1
            open_nodes = [n for n in pending if n.lp._solved_key is None and
                          n.dual_bound < self.primal_bound]
            batch += heapq.nsmallest(want, open_nodes)
        sent = type(node).prefetch(batch)
        self.prefetch_calls += 1
        self.prefetched_lps += sent
        # with several GPUs the shards of the batch agree on [incumbent, dual bound] through
        # blp_allreduce_min; kept for monitoring — pruning stays in the reference's node order
        self.frontier_bounds = getattr(getattr(node.lp, '_shared', None), 'last_global_bounds', (None, None))

    def _evaluate_node(self, node: BaseNode) -> None:
        """Bound the node unless the incumbent prunes it; then record a new incumbent or branch
        (reference :243-266)."""
        if not node.dual_bound < self.primal_bound:
            return
        self.evaluated_nodes += 1
        self._process_bound_rtn(node.bound(**self._kwargs))
        if node.unbounded:
            self._unbounded = True
        if getattr(node, 'lp_unsolved', False):
            # the LP hit the solver's iteration budget: the node stays an open leaf carrying the dual
            # bound it reached (never dropped as infeasible, never branched on an unconverged point)
            self.unsolved_nodes += 1
            return
        if node.lp_feasible and node.objective_value < self.primal_bound:
            if node.mip_feasible:
                self._best_solution = node.solution
                self.primal_bound = node.objective_value
            else:
                self._process_branch_rtn(node.idx, node.branch(**self._kwargs))

    def _process_branch_rtn(self, parent_id: int, rtn: Dict[str, Any]):
        assert isinstance(rtn, dict), 'rtn must be a dictionary'
        assert isinstance(parent_id, int), 'parent_id must be integer'
        assert parent_id in self.tree, 'parent must already exist in tree'
        for direction in ['left', 'right']:
            assert direction in rtn, f'{direction} must be in the returned dict'
            child = rtn[direction]
            assert isinstance(child, self._Node), f'{direction} value must be type {type(self._Node)}'
            assert child.idx not in self.tree, 'please give unique node ID'
            del rtn[direction]          # only a child that was accepted leaves the caller's dict
            self._node_queue.put(child)
            getattr(self.tree, f'add_{direction}_child')(child.idx, parent_id, node=child)
        self._process_rtn(rtn)

    def _process_bound_rtn(self, rtn: Dict[str, Any]):
        """Cuts returned under key 'cuts' are shared with every queued node's cut pool
        (reference :291-306); the rest goes on the kwargs bus."""
        assert isinstance(rtn, dict), 'rtn must be a dictionary'
        cuts = rtn.pop('cuts', None) if 'cuts' in rtn else None
        if cuts:
            for name, (pi, pi0) in cuts.items():
                for queued in self._node_queue.queue:
                    queued.cut_pool[name] = (pi, pi0)
        self._process_rtn(rtn)
