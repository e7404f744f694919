"""Numerical constants of the solve path (reference core/primitives/constants.py:5-26,
core/primitives/soft_math.py:16-27)."""

EPS_NUMERICAL = 1e-15
EPS_GEOMETRIC = 1e-6
MIN_CHIRALITY_VOLUME = 1e-6

SOLVE_TOLERANCE_VALUE = 1e-5
SOLVE_TOLERANCE_STEP = 1e-9
SOLVE_TOLERANCE_GRAD = 1e-9
SOLVE_ACCEPT_RESIDUAL = 1e-3

TEST_TOLERANCE = 1e-3
MM_PER_INCH = 25.4

# softnorm(s) = sqrt(s + EPS_SQ) - EPS
EPS = EPS_GEOMETRIC
EPS_SQ = EPS * EPS
