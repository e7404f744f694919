"""Parity of the CUDA path (through the C ABI) against the reference golden vectors and the
oracle.  Runs on the B200 box: ``pytest -m gpu``."""

import json
import os

import numpy as np
import pytest

from helpers import (GOLDEN, SWEEP_CASES, authored_positions, build_case, key_from_name, load_golden,
                     oracle_problem, perturbed_hardpoints)
from test_emu_core import POS_TOL_MM, _nominal, _program, check_failure_flags

pytestmark = pytest.mark.gpu


def gpu_solve(prog, hardpoints, values, **cfg):
    from open_kinematics_b200 import _lib
    topo = _lib.DeviceTopology(prog)
    try:
        out = topo.solve_batch(hardpoints, values, _lib.default_cfg(**cfg), want_tangents=True,
                               want_metrics=bool(prog.metric_names))
    finally:
        topo.close()
    return out


@pytest.mark.parametrize("case", SWEEP_CASES)
def test_positions_match_reference_tight_run(case):
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    prog, values = _program(sus, sweep)
    out = gpu_solve(prog, _nominal(sus, prog), values)
    assert out["status"][0] == 0 and out["failed_step"][0] == -1
    order = [prog.out_keys.index(key_from_name(n)) for n in meta["point_keys"]]
    diff = np.abs(out["positions"][0][:, order] - arr["positions_tight"]).max()
    assert diff <= POS_TOL_MM, diff
    scale = max(1.0, np.abs(arr["tangents"]).max())
    assert np.abs(out["tangents"][0] - arr["tangents"]).max() <= 1e-7 * scale
    assert out["max_residual"].max() < 1e-5


@pytest.mark.parametrize("case", SWEEP_CASES)
def test_metric_rows_match_reference(case):
    """State, mechanism and derivative metrics evaluated on the device vs the reference's
    compute_sweep_metrics rows (angles within 1e-9 rad)."""
    from open_kinematics_b200.core.topology import compile_suspension
    from test_emu_metrics import check_metrics
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    prog = compile_suspension(sus, sweep)
    assert prog.metric_names == meta["metric_names"]
    out = gpu_solve(prog, _nominal(sus, prog), arr["sweep_values"])
    assert out["status"][0] == 0
    check_metrics(prog.metric_names, out["metrics"][0], arr["metrics"])


@pytest.mark.parametrize("batch,case", [("batch_c1", "c1_dw_corner_bump"), ("batch_c2", "c2_macpherson_bump_steer"),
                                        ("batch_c3", "c3_rocker_ubar_coilover_roll")])
def test_perturbed_instances_match_reference(batch, case):
    meta, _ = load_golden(case)
    bmeta, barr = load_golden(batch)
    sus, sweep = build_case(meta)
    prog, values = _program(sus, sweep)
    out = gpu_solve(prog, perturbed_hardpoints(bmeta, prog), values)
    assert (out["status"] == 0).all()
    order = [prog.out_keys.index(key_from_name(n)) for n in bmeta["point_keys"]]
    assert np.abs(out["positions"][:, :, order] - barr["positions_tight"]).max() <= POS_TOL_MM


def test_failure_flags_match_reference():
    cases = json.load(open(os.path.join(GOLDEN, "failures.json")))
    check_failure_flags(lambda prog, hp, values: gpu_solve(prog, hp, values), cases)


def test_reference_boundary_solve_suspension_sweep():
    """Boundary B1 with the reference's own call shape (tests/core/test_solver.py:156-208 style)."""
    from open_kinematics_b200.core.points.derived.manager import DerivedPointsManager
    from open_kinematics_b200.core.solver import SolverInfo, solve_suspension_sweep
    from open_kinematics_b200.core.sweep import solve_sweep

    meta, arr = load_golden("c1_dw_corner_bump")
    sus, sweep = build_case(meta)
    initial = sus.initial_state()
    before = {k: p.data.copy() for k, p in initial.positions.items()}
    states, stats = solve_suspension_sweep(initial, sus.constraints(), sweep, DerivedPointsManager(sus.derived_spec()))
    assert len(states) == len(stats) == sweep.n_steps
    assert all(isinstance(s, SolverInfo) and s.converged and s.max_residual < 1e-3 for s in stats)
    assert all(np.array_equal(before[k], initial.positions[k].data) for k in before)  # inputs not mutated
    keys = [key_from_name(n) for n in meta["point_keys"]]
    got = np.array([[st.positions[k].data for k in keys] for st in states])
    assert np.abs(got - arr["positions_tight"]).max() <= POS_TOL_MM
    states2, _ = solve_sweep(sus, sweep)
    assert np.abs(np.array([[st.positions[k].data for k in keys] for st in states2]) - got).max() < 1e-9


def test_reference_boundary_errors():
    """Infeasible target -> RuntimeError naming the step (tests/core/test_solver.py:223-244);
    underdetermined system -> ValueError (tests/core/test_solver.py:211-220)."""
    from open_kinematics_b200.core.sweep import solve_sweep
    cases = json.load(open(os.path.join(GOLDEN, "failures.json")))
    rec = cases["dw_corner_rocker_bump_-60_+80"]
    sus, sweep = build_case(rec)
    with pytest.raises(RuntimeError, match=r"sweep step 3[456] did not reach an acceptable residual"):
        solve_sweep(sus, sweep)


def test_large_batch_properties():
    """Size-independent properties at batch scale (2e4 perturbed C3 instances):
    every accepted state satisfies the rigid-link lengths of its own instance, repeated
    instances give bit-identical answers, and instance order does not matter."""
    meta, _ = load_golden("c3_rocker_ubar_coilover_roll")
    sus, sweep = build_case(meta)
    prog, values = _program(sus, sweep)
    rng = np.random.default_rng(11)
    nominal = _nominal(sus, prog)[0]
    n = 20000
    hp = nominal[None, :] + rng.normal(0.0, 0.25, size=(n, nominal.size))
    hp[n // 2:] = hp[: n // 2]                      # duplicates
    out = gpu_solve(prog, hp, values)
    ok = out["status"] == 0
    assert ok.mean() > 0.99
    assert np.array_equal(out["positions"][: n // 2][ok[: n // 2]], out["positions"][n // 2:][ok[n // 2:]])
    perm = rng.permutation(n)
    out2 = gpu_solve(prog, hp[perm], values)
    assert np.array_equal(out2["positions"][np.argsort(perm)][ok], out["positions"][ok])
    # link lengths: |p_a - p_b| constant over the sweep for every distance row (softnorm bias 1e-6)
    from open_kinematics_b200.core.constraints import DistanceConstraint
    pos = out["positions"][ok]
    for c in sus.constraints():
        if isinstance(c, DistanceConstraint):
            a, b = prog.out_keys.index(c.p1), prog.out_keys.index(c.p2)
            ia, ib = prog.in_keys.index(c.p1), prog.in_keys.index(c.p2)
            design = np.linalg.norm(hp[ok][:, 3 * ia: 3 * ia + 3] - hp[ok][:, 3 * ib: 3 * ib + 3], axis=1)
            length = np.linalg.norm(pos[:, :, a] - pos[:, :, b], axis=2)
            assert np.abs(length - design[:, None]).max() < 1e-4
    assert out["max_residual"][ok].max() <= 1e-3


def test_empty_batch_and_zero_steps():
    meta, _ = load_golden("c1_dw_corner_bump")
    sus, sweep = build_case(meta)
    prog, values = _program(sus, sweep)
    out = gpu_solve(prog, np.zeros((0, 3 * prog.n_in)), values)
    assert out["positions"].shape[0] == 0 and out["status"].shape == (0,)
    out = gpu_solve(prog, _nominal(sus, prog), values[:, :0])
    assert out["status"][0] == 0 and out["positions"].shape[1] == 0


def test_invalid_geometry_is_flagged_not_propagated():
    """A NaN / collapsed instance must not poison its warp-mates."""
    meta, arr = load_golden("c1_dw_corner_bump")
    sus, sweep = build_case(meta)
    prog, values = _program(sus, sweep)
    hp = np.repeat(_nominal(sus, prog), 8, axis=0)
    hp[3, :] = np.nan
    hp[5, :] = 0.0
    out = gpu_solve(prog, hp, values)
    assert out["status"][3] != 0 and out["status"][5] != 0
    good = [0, 1, 2, 4, 6, 7]
    assert (out["status"][good] == 0).all()
    assert np.abs(out["positions"][good] - out["positions"][0]).max() == 0.0


from helpers import SHIM_CASES  # noqa: E402


@pytest.mark.parametrize("case", SHIM_CASES)
def test_camber_shim_presolve_matches_reference(case):
    """Setup pose from the on-device shim assembly pre-solve, the sweep solved from it and its
    metric rows vs the reference; also the host model's initial_state(), which asks the device
    for the setup pose (no host-side shim arithmetic)."""
    from open_kinematics_b200.core.topology import compile_suspension
    from test_emu_metrics import check_metrics
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    prog = compile_suspension(sus, sweep)
    out = gpu_solve(prog, _nominal(sus, prog), arr["sweep_values"])
    assert out["status"][0] == 0
    order = [prog.out_keys.index(key_from_name(n)) for n in meta["point_keys"]]
    assert np.abs(out["positions"][0][:, order] - arr["positions_tight"]).max() <= POS_TOL_MM
    check_metrics(prog.metric_names, out["metrics"][0], arr["metrics"])
    state = sus.initial_state()
    got = np.array([state.positions[key_from_name(n)].data for n in meta["point_keys"]])
    assert np.abs(got - arr["design_positions"]).max() <= POS_TOL_MM
    # single-instance boundary on the shimmed model
    from open_kinematics_b200.core.sweep import solve_sweep
    states, _ = solve_sweep(sus, sweep)
    got = np.array([[st.positions[key_from_name(n)].data for n in meta["point_keys"]] for st in states])
    assert np.abs(got - arr["positions_tight"]).max() <= POS_TOL_MM


def test_per_instance_shim_thickness_batch():
    """Monte-Carlo over shim thickness (BASELINE configs[3]): each instance's setup pose is
    solved on the device; thickness == design reproduces the un-shimmed sweep."""
    from open_kinematics_b200.core.sweep import BatchSolver
    meta, arr = load_golden("c4_tbar_heave_shim_roll")
    sus, sweep = build_case(meta)
    solver = BatchSolver(sus, sweep)
    prog = solver.program
    n = 64
    hp = np.repeat(solver.nominal_hardpoints()[None, :], n, axis=0)
    params = np.repeat(prog.param_default[None, :], n, axis=0)
    setup_cols = [i for i, name in enumerate(prog.param_names) if name.endswith("setup_thickness")]
    rng = np.random.default_rng(4)
    params[:, setup_cols] = rng.uniform(29.5, 30.5, size=(n, len(setup_cols)))
    params[0, setup_cols] = 31.0      # the golden instance (left shim; the mirrored right follows)
    params[1, setup_cols] = 30.0      # no-op shim
    res = solver.solve(hp, params=params)
    assert (res.status == 0).all()
    order = [prog.out_keys.index(key_from_name(k)) for k in meta["point_keys"]]
    assert np.abs(res.positions[0][:, order] - arr["positions_tight"]).max() <= POS_TOL_MM
    meta0, arr0 = load_golden("c4_tbar_roll")      # same T-bar axle without heave link / shim
    cam = prog.metric_names.index("camber_left")
    met = solver.solve(hp, params=params, want_metrics=True).metrics
    # camber at the design step is a monotone function of the shim thickness
    mid = arr["sweep_values"].shape[1] // 2
    # (instances within 0.01 mm of the design thickness are left out: below a 1e-6 rad upright
    # rotation the reference skips the attachment rotation, double_wishbone.py:554)
    far = np.abs(params[:, setup_cols[0]] - 30.0) > 0.01
    order_t = np.argsort(params[:, setup_cols[0]])
    order_t = order_t[far[order_t]]
    steps = np.diff(met[order_t, mid, cam])
    assert np.all(steps > 0) or np.all(steps < 0)
    assert abs(met[order_t[-1], mid, cam] - met[order_t[0], mid, cam]) > 0.1
    solver.close()
