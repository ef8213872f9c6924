"""Run the reference's own unittest modules on the HiGHS stand-in (oracle/ref_stubs.py).
Usage: python tools/run_reference_tests.py test_simple_mip_solver.test_nodes.test_base_node ..."""
import os, sys, unittest
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_stubs
ref_stubs.install()
ref_stubs.enable_cglp()
if not hasattr(unittest.TestCase, 'assertRegexpMatches'):
    unittest.TestCase.assertRegexpMatches = unittest.TestCase.assertRegex
mods = sys.argv[1:] or ['test_simple_mip_solver.test_nodes.test_base_node',
                        'test_simple_mip_solver.test_nodes.test_branch.test_pseudo_cost',
                        'test_simple_mip_solver.test_nodes.test_search.test_depth_first',
                        'test_simple_mip_solver.test_algorithms.test_base_algorithm',
                        'test_simple_mip_solver.test_algorithms.test_branch_and_bound',
                        'test_simple_mip_solver.test_utils.test_floating_point']
suite = unittest.TestSuite()
for m in mods:
    suite.addTests(unittest.defaultTestLoader.loadTestsFromName(m))
res = unittest.TextTestRunner(verbosity=1).run(suite)
print('RAN', res.testsRun, 'FAILURES', len(res.failures), 'ERRORS', len(res.errors), 'SKIPPED', len(res.skipped))
for t, tb in res.failures + res.errors:
    print('---', t.id()); print(tb.strip().splitlines()[-1][:300])
