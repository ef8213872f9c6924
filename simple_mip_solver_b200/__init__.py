"""B200-native batched LP bound step behind simple_mip_solver's Node / BranchAndBound API."""
from simple_mip_solver_b200.algorithms.branch_and_bound import BranchAndBound, BranchAndBoundTree
from simple_mip_solver_b200.compat import CyClpSimplex, CyLPArray, MILPInstance
from simple_mip_solver_b200.nodes.base_node import BaseNode
from simple_mip_solver_b200.nodes.branch.pseudo_cost import PseudoCostBranchNode
from simple_mip_solver_b200.nodes.bound.disjunctive_cut import DisjunctiveCutBoundNode
from simple_mip_solver_b200.nodes.nodes import (DisjunctiveCutBoundPseudoCostBranchNode,
                                                PseudoCostBranchDepthFirstSearchNode)
from simple_mip_solver_b200.utils.cut_generating_lp import CutGeneratingLP
from simple_mip_solver_b200.nodes.search.depth_first import DepthFirstSearchNode

__version__ = '0.1.0'
