"""utils/floating_point against input/output pairs recorded from the reference's
simple_mip_solver/utils/floating_point.py (tests/golden/make_goldens.py)."""
import json
import os

import numpy as np
import pytest

from simple_mip_solver_b200.compat import CyLPArray
from simple_mip_solver_b200.utils.floating_point import get_fraction, numerically_safe_cut, scale_cut

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'floating_point.json')))


def test_get_fraction_matches_reference():
    for g in GOLD['get_fraction']:
        n, d = get_fraction(g['x'], max_term=g['max_term'], estimate=g['estimate'])
        assert (n, d) == (g['n'], g['d']), g


def test_numerically_safe_cut_matches_reference():
    for g in GOLD['numerically_safe_cut']:
        pi, pi0 = numerically_safe_cut(CyLPArray(g['pi']), g['pi0'], estimate=g['estimate'])
        assert np.array_equal(np.asarray(pi), np.asarray(g['safe_pi'])), g
        assert pi0 == g['safe_pi0']


def test_safe_cut_is_an_outer_approximation():
    # property of the reference's tests (test_floating_point.py:44-111): the rounded cut is implied
    rng = np.random.default_rng(3)
    for _ in range(200):
        k = int(rng.integers(2, 7))
        pi = rng.uniform(0, 10, k)
        pi0 = float(rng.uniform(0, 10))
        spi, spi0 = numerically_safe_cut(CyLPArray(pi), pi0, estimate='over')
        scaled, scaled0 = scale_cut(pi, pi0)
        assert (np.asarray(spi) >= scaled - 1e-15).all() and spi0 <= scaled0 + 1e-15


def test_argument_checks():
    with pytest.raises(AssertionError, match='pi is a CyLPArray'):
        numerically_safe_cut(np.ones(2), 1.0)
    with pytest.raises(AssertionError, match='estimate must be over or under'):
        numerically_safe_cut(CyLPArray([1, 2]), 1.0, estimate='x')
    assert scale_cut(np.zeros(3), 1.0) == (None, None)
    assert get_fraction(2.0) == (2, 1)
