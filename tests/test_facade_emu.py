"""Host-side mirror of the reference facade (B1-B3, core/sweep.py) exercised on the CPU with the
lane-emulation build standing in for the device.  The same test bodies run against the real
library in tests/test_gpu_parity.py."""

import pytest

import test_gpu_parity as G


@pytest.mark.parametrize("case", ["c1_dw_corner_bump_steer"])
def test_compute_state_tangents(emu_device, case):
    G.test_reference_boundary_compute_state_tangents(case)


@pytest.mark.parametrize("case", ["c1_dw_corner_bump_steer", "c4_tbar_heave_shim_roll"])
def test_metrics_of_solved_states(emu_device, case):
    G.test_reference_boundary_metrics_of_solved_states(case)


def test_solve_suspension_sweep_boundary(emu_device):
    G.test_reference_boundary_solve_suspension_sweep()
    G.test_reference_boundary_errors()


def test_sweep_diagnostics(emu_device):
    G.test_sweep_diagnostics_match_reference()


def test_result_files(emu_device, tmp_path):
    G.test_result_files_match_reference(tmp_path)


@pytest.mark.parametrize("case", ["c1_dw_corner_bump_steer", "c3_rocker_ubar_coilover_roll", "c4_tbar_roll"])
def test_analyze_sweep(emu_device, case):
    G.test_analyze_sweep_matches_reference(case)


@pytest.mark.parametrize("case", ["c1_dw_corner_bump_steer", "c3_rocker_ubar_coilover_roll"])
def test_metrics_of_default_tolerance_states(emu_device, case):
    """States of reference accuracy (its default ftol = 1e-5 run, up to 2.3e-5 mm from the tight
    solution) are valid inputs of the B2/B3 facade: nothing raises, and the metric rows agree with
    the reference's rows for the tight states to that accuracy."""
    import numpy as np
    from helpers import build_case, key_from_name, load_golden
    from open_kinematics_b200.core.metrics.main import AxleMetricRows
    from open_kinematics_b200.core.primitives.geometry import Point3
    from open_kinematics_b200.core.state import SuspensionState
    from open_kinematics_b200.core.sweep import compute_sweep_metrics
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    keys = [key_from_name(n) for n in meta["point_keys"]]
    free = set(sus.initial_state().free_points)
    states = [SuspensionState(positions={k: Point3(arr["positions_default"][s, i]) for i, k in enumerate(keys)},
                              free_points=set(free)) for s in range(sweep.n_steps)]
    assert np.abs(arr["positions_default"] - arr["positions_tight"]).max() > 5e-6   # the case is meaningful
    result = compute_sweep_metrics(sus, sweep, states)
    flat = [row.flat_row() if isinstance(row, AxleMetricRows) else row for row in result.rows]
    got = np.array([[np.nan if r[k] is None else r[k] for k in meta["metric_names"]] for r in flat])
    ref = arr["metrics"]
    assert (np.isnan(got) == np.isnan(ref)).all()
    scale = np.maximum(1.0, np.abs(ref))
    assert np.nanmax(np.abs(got - ref) / scale) <= 1e-4


def test_per_instance_target_tables(emu_device):
    G.test_per_instance_target_tables()


def test_worst_residual_row_is_reported(emu_device):
    G.test_worst_residual_row_is_reported()


def test_vectors_parallel_and_spherical_families(emu_device):
    G.test_vectors_parallel_and_spherical_families_solve_like_the_reference()
