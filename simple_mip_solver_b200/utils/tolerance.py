"""Numerical constants of the bound step.

Values follow the reference's ``simple_mip_solver/utils/tolerance.py:1-43`` so that
integrality tests, cut loops and stall tests make the same decisions.
"""

# |x - round(x)| above this => the value counts as fractional (tolerance.py:2)
variable_epsilon = 1e-4

# a fractional part closer than this to 0/1 is skipped when building GMI cuts (tolerance.py:5)
good_coefficient_approximation_epsilon = 1e-2

# two fractions closer than this are the same (tolerance.py:8)
exact_coefficient_approximation_epsilon = 1e-14

# slack allowed when validating a cut (tolerance.py:11)
cut_tolerance = 1e-14

# cuts with more nonzeros than this are ignored (tolerance.py:14)
max_nonzero_coefs = 1000000

# two cuts closer than this many degrees are "parallel" (tolerance.py:17)
parallel_cut_tolerance = 10

# relative objective progress a cut round must make to continue (tolerance.py:22)
cutting_plane_progress_tolerance = 1e-4

# cut rounds per node before branching (tolerance.py:26)
max_cut_generation_iterations = 10

# |cut coef| may exceed the largest root coefficient by at most this factor (tolerance.py:29)
max_relative_cut_term_ratio = 1000

# a cut must be violated by more than this euclidean depth (tolerance.py:34)
min_cut_depth = 1e-8

# smallest acceptable norm of a disjunctive cut (tolerance.py:37)
min_cglp_norm = 1e-4

# largest numerator/denominator in rational cut rounding (tolerance.py:43)
max_term = 1e3
