"""The reference's OWN unit tests, unmodified, against this package (drop-in check).

tests/tools/run_reference_tests_on_product.py aliases ``simple_mip_solver`` / ``cylp`` / ``coinor`` to this
package and runs the reference's 136 unittest cases on the CPU stand-in engine. They live in
/root/reference, which exists only in the authoring container: elsewhere this test is skipped.
What may fail, and why, is listed here; anything else failing is a regression of the drop-in surface.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_TESTS = '/root/reference/test_simple_mip_solver'

ALLOWED = {
    # CLP artefacts: the "objective" -6.25e13 CLP reports for an unbounded LP (do not chase), and the
    # point CLP's primal simplex happens to stop at on unbounded random LPs
    'TestBaseNode.test_bound_lp_unbounded',
    'TestFloatingPoint.test_numerically_safe_cut',
    # CLP's choice among alternative optimal vertices of cut1 (the root ends integral there, so the CGLP
    # has one term; on a textbook simplex and on HiGHS the root branches, golden 'reference' / 'reference_ds')
    'TestNode.test_get_cglp_starting_basis',
    # "rough values" of the reference's own helper (helpers.py:62-73): a pool cut created in one round and
    # appended in a later one makes iterations_gmic_added exceed iterations_gmic_created by one on 3 of 818
    # (model, option set) pairs; which pairs a run draws depends on the global RNG
    'TestNode.test_zmodels', 'TestDisjunctiveCutBoundPseudoCostBranchNode.test_models',
    # out of scope: the parametric dual bound, and tests that take CyLP's multi-variable
    # model of the CGLP apart (pi, pi0, u_t, w_t, v_t as named CyLP variables)
    'TestBranchAndBound.test_bound_parameterized_dual', 'TestBranchAndBound.test_bound_parameterized_dual_fails_asserts',
    'TestBranchAndBound.test_find_parameterized_dual_bound',
    'TestBranchAndBound.test_find_parameterized_dual_bound_fails_asserts',
    'TestBranchAndBound.test_find_parameterized_dual_bound_many_times',
    'TestCutGeneratingLP.test_create_cglp_depth_1', 'TestCutGeneratingLP.test_create_cglp_fails_asserts',
    'TestCutGeneratingLP.test_create_cglp_infinite_bounds',
    'TestCutGeneratingLP.test_create_cglp_new_coef_matrix_and_var_bounds',
    'TestCutGeneratingLP.test_create_cglp_standard',
}


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason='the reference checkout is not on this machine')
def test_the_references_unit_tests_pass_on_this_package():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'tools', 'run_reference_tests_on_product.py')],
                         capture_output=True, text=True, timeout=1500, cwd=ROOT)
    text = out.stdout + out.stderr
    summary = re.search(r'^RAN (\d+) FAILURES (\d+) ERRORS (\d+)', text, re.M)
    assert summary, text[-2000:]
    ran, failures, errors = (int(g) for g in summary.groups())
    failed = {'.'.join(line.split()[1].split('.')[-2:]) for line in text.splitlines() if line.startswith('--- ')}
    assert ran == 136
    assert failed <= ALLOWED, sorted(failed - ALLOWED)
    assert ran - failures - errors >= 121
