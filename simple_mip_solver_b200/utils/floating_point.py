"""Rational rounding of cut coefficients so that a rounded cut stays valid.

Same contract as the reference's ``simple_mip_solver/utils/floating_point.py``
(``scale_cut`` :11, ``numerically_safe_cut`` :40, ``get_fraction`` :106): a cut ``pi.x >= pi0`` is
scaled to unit max coefficient, every coefficient is replaced by a nearby fraction with numerator
and denominator at most ``max_term`` that errs on the safe side, and the right-hand side is
rounded the other way. Scalar host arithmetic; not GPU work.
"""
from __future__ import annotations

import warnings
from math import ceil, floor
from typing import Optional, Tuple, Union

import numpy as np

from simple_mip_solver_b200.compat.cylp_like import CyLPArray
from simple_mip_solver_b200.utils.tolerance import (
    exact_coefficient_approximation_epsilon, good_coefficient_approximation_epsilon, max_term)


def scale_cut(pi: np.ndarray, pi0: float, max_abs: float = 1, **kwargs) -> \
        Union[Tuple[np.ndarray, float], Tuple[None, None]]:
    """Scale (pi, pi0) so the largest |coefficient| equals ``max_abs``; (None, None) if pi == 0."""
    assert isinstance(pi, np.ndarray), 'pi is an nd.array'
    assert isinstance(pi0, (float, int)), 'pi0 is a number'
    assert isinstance(max_abs, (int, float)) and max_abs > 0, 'max_abs should be positive'
    biggest = float(np.max(np.abs(pi))) if pi.size else 0.0
    if biggest == 0:
        return None, None
    factor = max_abs / biggest
    return pi * factor, pi0 * factor


def get_fraction(x: float, max_term: int = max_term, estimate: Optional[str] = None, **kwargs) -> Tuple[int, int]:
    """Continued-fraction convergent of ``x`` with numerator and denominator <= ``max_term``.

    ``estimate='over'`` returns a fraction >= x, ``'under'`` one <= x, None the last convergent
    within the size limit. Convergents alternate: even-indexed ones lie below x, odd-indexed above.
    """
    assert isinstance(x, (int, float)), 'x should be an int or float'
    assert isinstance(max_term, (int, float)) and max_term > 0, 'max_term should be positive'
    if estimate is not None:
        assert estimate in ['over', 'under'], "estimate should be 'over' or 'under' when provided"
    if abs(x) > max_term:
        whole = ceil(x) if estimate == 'over' else floor(x) if estimate == 'under' else round(x)
        return whole, 1

    nums, dens = [], []                 # convergents h_i / k_i
    h_prev2, k_prev2, h_prev1, k_prev1 = 0, 1, 1, 0
    value = x
    exact = False
    while True:
        a = floor(value)
        h, k = a * h_prev1 + h_prev2, a * k_prev1 + k_prev2
        nums.append(h)
        dens.append(k)
        if h > max_term or k > max_term:
            break
        rest = value - a
        if not rest:
            exact = True
            break
        value = 1 / rest
        h_prev2, k_prev2, h_prev1, k_prev1 = h_prev1, k_prev1, h, k
    last = len(nums) - 1                # index of the convergent that stopped the loop
    if exact:
        return nums[last], dens[last]

    def conv(i):
        return nums[i], dens[i]

    if estimate == 'over':
        if (last - 1) % 2:
            return conv(last - 1)
        return conv(last - 2) if last - 2 >= 0 else (ceil(x), 1)
    if estimate == 'under':
        if not (last - 1) % 2:
            return conv(last - 1) if last - 1 >= 0 else (floor(x), 1)
        return conv(last - 2) if last - 2 >= 0 else (floor(x), 1)
    return conv(last - 1) if last - 1 >= 0 else (round(x), 1)


def numerically_safe_cut(pi: CyLPArray, pi0: float, estimate: str = 'over',
                         make_integer: bool = False, **kwargs) -> Tuple[CyLPArray, float]:
    """Outer approximation of ``pi.x >= pi0`` (estimate 'over') or ``pi.x <= pi0`` ('under')
    with small rational coefficients."""
    assert isinstance(pi, CyLPArray), 'pi is a CyLPArray'
    assert isinstance(pi0, (float, int)), 'pi0 is a number'
    assert estimate in ['over', 'under'], 'estimate must be over or under to ensure safety'
    scaled, scaled0 = scale_cut(pi, pi0, **kwargs)
    if scaled is None:
        return pi, pi0
    nums, dens = [], []
    for coef in scaled:
        coef = float(coef)
        n, d = get_fraction(coef, estimate=estimate, **kwargs)
        if coef != 0 and abs(1 - ((n / d) / coef)) > good_coefficient_approximation_epsilon:
            # the one-sided fraction is poor: accept the nearest one if it is exact
            n2, d2 = get_fraction(coef, estimate=None, **kwargs)
            if abs(n2 / d2 - coef) < exact_coefficient_approximation_epsilon:
                n, d = n2, d2
        nums.append(n)
        dens.append(d)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        lcm = np.lcm.reduce(np.asarray(dens, dtype=np.int64))
    mult = lcm if make_integer else 1
    safe_pi = CyLPArray(mult * np.array(nums, dtype=float) / np.array(dens, dtype=float))
    other = 'under' if estimate == 'over' else 'over'
    n, d = get_fraction(x=float(scaled0 * lcm) if make_integer else float(scaled0), estimate=other)
    return safe_pi, n / d
