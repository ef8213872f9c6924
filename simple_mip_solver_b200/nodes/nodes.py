"""Node classes combining search, branch and bound variants (reference nodes/nodes.py:9-14)."""
from simple_mip_solver_b200.nodes.bound.disjunctive_cut import DisjunctiveCutBoundNode
from simple_mip_solver_b200.nodes.branch.pseudo_cost import PseudoCostBranchNode
from simple_mip_solver_b200.nodes.search.depth_first import DepthFirstSearchNode


class PseudoCostBranchDepthFirstSearchNode(PseudoCostBranchNode, DepthFirstSearchNode):
    pass


class DisjunctiveCutBoundPseudoCostBranchNode(DisjunctiveCutBoundNode, PseudoCostBranchNode):
    pass
