"""Minimal binary tree with the ``coinor.gimpy.tree.BinaryTree`` calls the reference makes
(simple_mip_solver/algorithms/branch_and_bound.py:19-108, 193, 285-286): ``add_root``,
``add_left_child``, ``add_right_child``, ``nodes[idx].attr``, ``get_children``, ``get_parent``,
``get_node_attr`` and ``in``."""
from __future__ import annotations

from typing import Any, Dict, List, Optional


class TreeVertex:
    def __init__(self, name, **attr):
        self.name = name
        self.attr: Dict[str, Any] = dict(attr)


class BinaryTree:
    def __init__(self):
        self.nodes: Dict[Any, TreeVertex] = {}
        self.root = None
        self._children: Dict[Any, Dict[str, Any]] = {}
        self._parent: Dict[Any, Any] = {}

    def __contains__(self, name):
        return name in self.nodes

    def __len__(self):
        return len(self.nodes)

    def add_root(self, name, **attr):
        assert self.root is None, 'the tree already has a root'
        self.root = name
        self.nodes[name] = TreeVertex(name, **attr)
        self._children[name] = {}

    def _add_child(self, name, parent, direction, **attr):
        assert parent in self.nodes, 'parent must already exist in tree'
        assert name not in self.nodes, 'node ids are unique'
        assert direction not in self._children[parent], f'parent already has a {direction} child'
        self.nodes[name] = TreeVertex(name, direction=direction, **attr)
        self._children[parent][direction] = name
        self._children[name] = {}
        self._parent[name] = parent

    def add_left_child(self, name, parent, **attr):
        self._add_child(name, parent, 'left', **attr)

    def add_right_child(self, name, parent, **attr):
        self._add_child(name, parent, 'right', **attr)

    def get_children(self, name) -> List[Any]:
        ch = self._children[name]
        return [ch[d] for d in ('left', 'right') if d in ch]

    def get_left_child(self, name):
        return self._children[name].get('left')

    def get_right_child(self, name):
        return self._children[name].get('right')

    def get_parent(self, name) -> Optional[Any]:
        return self._parent.get(name)

    def get_node_attr(self, name, attr):
        return self.nodes[name].attr.get(attr)

    def set_node_attr(self, name, attr, value):
        self.nodes[name].attr[attr] = value

    def get_node(self, name) -> TreeVertex:
        return self.nodes[name]
