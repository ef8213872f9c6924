"""DepthFirstSearchNode — deepest node first when kept in a priority queue
(reference simple_mip_solver/nodes/search/depth_first.py:16-28)."""
from __future__ import annotations

from typing import Any

from simple_mip_solver_b200.nodes.base_node import BaseNode


class DepthFirstSearchNode(BaseNode):

    def __init__(self, *args: Any, **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.search_method = 'depth first'

    def __eq__(self, other):
        if isinstance(other, DepthFirstSearchNode):
            return self.depth == other.depth
        raise TypeError('A DFS Node can only be compared with another DFS Node')

    def __lt__(self, other):
        if isinstance(other, DepthFirstSearchNode):
            return self.depth > other.depth
        raise TypeError('A DFS Node can only be compared with another DFS Node')

    __hash__ = object.__hash__
