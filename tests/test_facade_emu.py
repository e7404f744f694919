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


@pytest.mark.parametrize("case", ["c1_dw_corner_bump_steer", "c3_rocker_ubar_coilover_roll"])
def test_analyze_sweep(emu_device, case):
    G.test_analyze_sweep_matches_reference(case)
