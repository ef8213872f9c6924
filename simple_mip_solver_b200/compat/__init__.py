from .binary_tree import BinaryTree
from .cylp_like import COIN_INFINITY, CyClpSimplex, CyLPArray, SharedLP, solve_lps
from .milp_instance import MILPInstance
from .mps import read_mps
