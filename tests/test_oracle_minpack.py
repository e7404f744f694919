"""Pin the MINPACK lmder restatement (oracle/minpack_lm.py) against SciPy's own _lmder -- the
third-party routine the reference's solve calls (core/solver.py:169) -- on the reference's
problems: same termination code, same number of residual evaluations, same iterate."""

import numpy as np
import pytest
from scipy.optimize import _minpack, least_squares

from helpers import authored_positions, build_case, load_golden, oracle_problem
from oracle.minpack_lm import lmder
from oracle.solve import ResidualComputer, design_setup, solve_sweep, target_bases


def _problem(case):
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    problem, values = oracle_problem(sus, sweep)
    pos0, consts = design_setup(problem, authored_positions(sus))
    rc = ResidualComputer(problem, pos0, consts)
    x0 = np.concatenate([pos0[k] for k in problem.free_order])
    return problem, rc, x0, target_bases(problem, pos0), values, arr


def test_rosenbrock_like_small_problem_matches_scipy():
    def fun(x):
        return np.array([10 * (x[1] - x[0] ** 2), 1 - x[0], x[0] * x[1] - 1.0])

    def jac(x):
        return np.array([[-20 * x[0], 10.0], [-1.0, 0.0], [x[1], x[0]]])

    x0 = np.array([-1.2, 1.0])
    x, fvec, info, nfev, njev = lmder(fun, jac, x0, ftol=1e-10, xtol=1e-10, gtol=1e-10)
    ref, out, status = _minpack._lmder(fun, jac, x0.copy(), (), True, False, 1e-10, 1e-10, 1e-10, 200, 100.0, None)
    assert info == status and nfev == out["nfev"] and njev == out["njev"]
    np.testing.assert_allclose(x, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", ["c1_dw_corner_bump", "c2_macpherson_bump_steer"])
def test_first_sweep_steps_match_scipy_lmder(case):
    """Rank-deficient kinematics systems: the restated iteration follows SciPy's on the first
    sweep steps (MINPACK info, nfev within a few evaluations, answer within the reference's own
    termination noise).  Bit-identical paths are not expected: the LM path on a rank-deficient
    Jacobian is sensitive to the last bit of the Householder sums."""
    problem, rc, x0, bases, values, arr = _problem(case)
    x = x0.copy()
    for s in range(3):
        tabs = bases + values[:, s]
        mine = lmder(lambda v: rc.compute(v, tabs), lambda v: rc.compute_jacobian(v, tabs), x)
        ref, out, status = _minpack._lmder(lambda v: rc.compute(v, tabs), lambda v: rc.compute_jacobian(v, tabs),
                                           x.copy(), (), True, False, 1e-5, 1e-9, 1e-9, 100 * x.size, 100.0, None)
        assert mine[2] in (1, 2, 3, 4) and status in (1, 2, 3, 4)
        assert abs(mine[3] - out["nfev"]) <= max(8, 0.3 * out["nfev"])
        assert np.abs(mine[0] - ref).max() <= 5e-5
        assert np.abs(mine[1]).max() < 1e-4
        x = ref


def test_oracle_sweep_with_restated_lm_matches_reference_tight_run():
    meta, arr = load_golden("c1_dw_corner_bump")
    sus, sweep = build_case(meta)
    problem, values = oracle_problem(sus, sweep)
    out = solve_sweep(problem, authored_positions(sus), values[:, :6], ftol=1e-15, xtol=1e-15, gtol=1e-15,
                      lm="restated")
    assert out["status"] == 0
    np.testing.assert_allclose(out["positions"], arr["positions_tight"][:6], rtol=0, atol=1e-7)
