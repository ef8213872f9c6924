"""Run the reference's OWN unittest modules, unmodified, against THIS package (drop-in check).

`simple_mip_solver[.x.y]` resolves to `simple_mip_solver_b200[.x.y]`, `cylp` / `coinor.cuppy` /
`coinor.gimpy` to the product's look-alikes in `simple_mip_solver_b200.compat`. The LP engine is the
CPU stand-in of the test suite (tests/helpers.OracleBatchLP: the numpy restatement of the device's
dual simplex, HiGHS for the first-order path) unless --device is given on a box with a GPU.
`gurobipy` (the reference's independent MIP check on the 64 example models) is answered by HiGHS
branch and cut (oracle/gurobi_like.py). The reference's tests live in /root/reference, which exists only in the authoring container, so this
is a tool, not part of the suites.

    python tests/tools/run_reference_tests_on_product.py [--device] [--pdhg] [module ...]

--pdhg routes every node LP through the host code of the first-order path (`SharedLP.default_method = 'pdhg'`;
with the CPU stand-in that is HiGHS behind `solve_batch` / `solve_children`).
"""
import importlib
import os
import sys
import types
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

SUBMODULES = ['', '.algorithms', '.algorithms.base_algorithm', '.algorithms.branch_and_bound', '.nodes',
              '.nodes.base_node', '.nodes.nodes', '.nodes.bound', '.nodes.bound.disjunctive_cut', '.nodes.branch',
              '.nodes.branch.pseudo_cost', '.nodes.search', '.nodes.search.depth_first', '.utils',
              '.utils.cut_generating_lp', '.utils.floating_point', '.utils.tolerance']
DEFAULT = ['test_simple_mip_solver.test_nodes.test_base_node',
           'test_simple_mip_solver.test_nodes.test_nodes',
           'test_simple_mip_solver.test_nodes.test_branch.test_pseudo_cost',
           'test_simple_mip_solver.test_nodes.test_search.test_depth_first',
           'test_simple_mip_solver.test_nodes.test_bound.test_disjunctive_cut',
           'test_simple_mip_solver.test_algorithms.test_base_algorithm',
           'test_simple_mip_solver.test_algorithms.test_branch_and_bound',
           'test_simple_mip_solver.test_utils.test_cut_generating_lp',
           'test_simple_mip_solver.test_utils.test_floating_point']


def install(reference_root='/root/reference', device=False):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if not device:
        import simple_mip_solver_b200.engine as engine
        from helpers import OracleBatchLP
        engine.BatchLP = OracleBatchLP
        engine.default_opts = lambda **kw: kw
    for sub in SUBMODULES:
        sys.modules['simple_mip_solver' + sub] = importlib.import_module('simple_mip_solver_b200' + sub)
    from simple_mip_solver_b200 import compat
    from simple_mip_solver_b200.compat import cylp_like
    from oracle.ref_stubs import GenerateRandomMIP              # GrUMPy's generator (fixtures only)
    import scipy.sparse as sp
    mod('cylp')
    mod('cylp.cy', CyClpSimplex=compat.CyClpSimplex)
    mod('cylp.cy.CyClpSimplex', CyClpSimplex=compat.CyClpSimplex, CyLPArray=compat.CyLPArray)
    mod('cylp.py')
    mod('cylp.py.modeling')
    mod('cylp.py.modeling.CyLPModel', CyLPArray=compat.CyLPArray)
    mod('cylp.py.utils')
    mod('cylp.py.utils.sparseUtil', csc_matrixPlus=getattr(cylp_like, 'csc_matrixPlus', sp.csc_matrix))
    mod('coinor')
    mod('coinor.cuppy')
    mod('coinor.cuppy.milpInstance', MILPInstance=compat.MILPInstance)
    mod('coinor.gimpy')
    mod('coinor.gimpy.tree', BinaryTree=compat.BinaryTree)
    mod('coinor.grumpy')
    mod('coinor.grumpy.BranchAndBound', GenerateRandomMIP=GenerateRandomMIP)
    from oracle import gurobi_like                              # HiGHS MIP behind the gurobipy calls of the tests
    sys.modules['gurobipy'] = gurobi_like
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def main():
    args = [a for a in sys.argv[1:] if a not in ('--device', '--pdhg')]
    install(device='--device' in sys.argv)
    if '--pdhg' in sys.argv:
        from simple_mip_solver_b200.compat.cylp_like import SharedLP
        SharedLP.default_method = 'pdhg'
    import numpy as np
    np.random.seed(0)        # the reference's helpers sample 10 % of (model, option set) pairs with the global RNG
    if not hasattr(unittest.TestCase, 'assertRegexpMatches'):
        unittest.TestCase.assertRegexpMatches = unittest.TestCase.assertRegex
    suite = unittest.TestSuite()
    for m in args or DEFAULT:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromName(m))
    res = unittest.TextTestRunner(verbosity=1).run(suite)
    print('RAN', res.testsRun, 'FAILURES', len(res.failures), 'ERRORS', len(res.errors), 'SKIPPED', len(res.skipped))
    for t, tb in res.failures + res.errors:
        print('---', t.id())
        print('    ' + tb.strip().splitlines()[-1][:300])


if __name__ == '__main__':
    main()
