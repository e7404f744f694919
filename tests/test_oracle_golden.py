"""Pin the CPU oracle against vectors produced by the reference itself
(tests/golden/generate_golden.py).  CPU only."""

import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, SWEEP_CASES, authored_positions, build_case, load_golden, oracle_problem
from oracle import families as F
from oracle.solve import ResidualComputer, design_setup, solve_sweep, state_tangents, target_bases


def test_families_match_reference_rows():
    """Residual and analytical gradient of every family vs core/constraints.py + core/jacobians.py
    at seeded random points (the reference pins its own rows against central differences at
    atol 1e-6, tests/core/test_jacobians.py:30-35; here the two closed forms must agree to 1e-12)."""
    recs = json.load(open(os.path.join(GOLDEN, "families.json")))
    assert set(recs) == set(F.FAMILIES)
    for fam, rec in recs.items():
        for pts, cst, res, jac in zip(rec["points"], rec["consts"], rec["residual"], rec["jacobian"]):
            r, g = F.FAMILIES[fam](np.array(pts), cst)
            scale = max(1.0, np.abs(jac).max())
            assert abs(r - res) <= 1e-9 * max(1.0, abs(res)), fam
            np.testing.assert_allclose(g, np.array(jac), rtol=0, atol=1e-12 * scale, err_msg=fam)


@pytest.mark.parametrize("case", SWEEP_CASES)
def test_design_constants_match_reference(case):
    """Constants recomputed from hardpoints == the reference's Suspension.constraints() values."""
    meta, _ = load_golden(case)
    sus, sweep = build_case(meta)
    problem, _ = oracle_problem(sus, sweep)
    _, consts = design_setup(problem, authored_positions(sus))
    assert len(consts) == len(meta["constraints"])
    for mine, ref in zip(consts, meta["constraints"]):
        for attr, idx in (("target_distance", 0), ("target_angle", 0), ("target_volume", 0)):
            if attr in ref:
                assert abs(mine[idx] - ref[attr]) <= 1e-12 * max(1.0, abs(ref[attr]))
        if "scale" in ref:
            assert abs(1.0 / mine[1] - ref["scale"]) <= 1e-9 * ref["scale"]
        if "line_point" in ref:
            np.testing.assert_allclose(mine[0:3], ref["line_point"], atol=1e-12)


@pytest.mark.parametrize("case", ["c1_dw_corner_bump", "c2_macpherson_bump_steer", "c3_rocker_ubar_roll_shipped",
                                  "dw_axle_direct"])
def test_oracle_sweep_matches_reference_default_run(case):
    """Default tolerances (ftol 1e-5, xtol/gtol 1e-9).  The reference stops ~1e-5 mm from the root
    of its own equations (termination noise of MINPACK's xtol test on a rank-deficient system,
    SURVEY.md Appendix D: default<->tight = 0.8-2.3e-5 mm) and the LM path is sensitive to
    round-off in the callbacks, so two default runs agree only to that noise: 5e-5 mm here.
    The sharp pin is the tight-tolerance test below."""
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    problem, values = oracle_problem(sus, sweep)
    out = solve_sweep(problem, authored_positions(sus), values)
    assert out["status"] == 0
    order = [problem.point_keys.index(k) for k in sorted(problem.point_keys)]
    np.testing.assert_allclose(out["positions"][:, order], arr["positions_default"], rtol=0, atol=5e-5)
    ratio = out["nfev"].mean() / arr["nfev_default"].mean()
    assert 0.6 < ratio < 1.6
    assert out["max_residual"].max() < 1e-4 and arr["max_residual_default"].max() < 1e-4


@pytest.mark.parametrize("case", ["c1_dw_corner_bump", "c2_macpherson_bump_steer", "c4_tbar_roll"])
def test_oracle_tight_run_and_tangents(case):
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    problem, values = oracle_problem(sus, sweep)
    auth = authored_positions(sus)
    out = solve_sweep(problem, auth, values, ftol=1e-15, xtol=1e-15, gtol=1e-15)
    assert out["status"] == 0
    # two tight-tolerance LM runs agree to their own termination noise (~2e-8 mm)
    np.testing.assert_allclose(out["positions"], arr["positions_tight"], rtol=0, atol=1e-7)
    pos0, consts = design_setup(problem, auth)
    rc = ResidualComputer(problem, pos0, consts)
    bases = target_bases(problem, pos0)
    for s in (0, values.shape[1] // 2, values.shape[1] - 1):
        t = state_tangents(problem, rc, out["x"][s], bases + values[:, s])
        np.testing.assert_allclose(t["tangents"], arr["tangents"][s], rtol=0, atol=1e-7)
        assert t["rank"] == arr["tangent_rank"][s]
        assert abs(t["condition_number"] - arr["tangent_cond"][s]) <= 1e-6 * arr["tangent_cond"][s]


def test_oracle_failure_flags():
    """First failed step and failure class of out-of-reach sweeps (solver.py:726-747)."""
    cases = json.load(open(os.path.join(GOLDEN, "failures.json")))
    for label in ("c1_bump_to_+400", "dw_corner_rocker_bump_-60_+80", "c1_bump_to_-600"):
        rec = cases[label]
        sus, sweep = build_case(rec)
        problem, values = oracle_problem(sus, sweep)
        out = solve_sweep(problem, authored_positions(sus), values)
        assert (out["status"], out["failed_step"]) == (rec["status"], rec["failed_step"]), label


def test_metric_registry_columns_and_units_match_reference():
    """Host-only mirror of flat_specs_for_suspension (reference core/metrics/registry.py:201-215): the
    flat columns of every golden case, in export order, with the reference's unit symbols; and
    flatten_positions (core/export.py)."""
    import numpy as np
    from helpers import SWEEP_CASES, build_case, load_golden
    from open_kinematics_b200.core.export import flatten_positions
    from open_kinematics_b200.core.metrics.registry import flat_specs_for_suspension
    units = json.load(open(os.path.join(GOLDEN, "result_files.json")))["metric_units"]
    for case in SWEEP_CASES:
        meta, arr = load_golden(case)
        sus, sweep = build_case(meta)
        specs = flat_specs_for_suspension(sus, [dim[0] for dim in sweep.target_sweeps])
        assert list(specs) == meta["metric_names"], case
        for name, spec in specs.items():
            assert spec.unit == units[name], (case, name)
    meta, arr = load_golden("c1_dw_corner_bump")
    sus, _ = build_case(meta)
    flat = flatten_positions(sus.authored_state().positions, sus.output_points())
    assert list(flat) == [k.name.lower() for k in sus.output_points() if k in sus.authored_state().positions]
    assert all(isinstance(v, tuple) and len(v) == 3 for v in flat.values())
