"""Physical invariants over solved sweeps, in the manner of the reference's own integration tests
(tests/test_t_bar_arb.py:89-169, tests/test_macpherson.py:68-107, tests/test_axle_rocker.py:96-120,
tests/test_sensitivity.py:39-89, tests/test_steering.py:122-160): rigid links stay rigid, the strut
telescopes along its axis, wheel travel reaches the swept values, tangents agree with finite
differences of re-solved states.  Run through the public facade; on the CPU the lane-emulation
build stands in for the device (``emu_device``), ``-m gpu`` runs the same bodies on the B200."""

import numpy as np
import pytest

from helpers import build_case, load_golden
from open_kinematics_b200.core.enums import PointID as P
from open_kinematics_b200.core.metrics.main import AxleMetricRows
from open_kinematics_b200.core.primitives.point_ref import PointRef, Side
from open_kinematics_b200.core.sweep import compute_sweep_metrics, compute_sweep_tangents, solve_sweep


def _dist(state, a, b) -> float:
    return float(np.linalg.norm(state.positions[a].data - state.positions[b].data))


def t_bar_invariants():
    for case, opposed in (("c4_tbar_bump", False), ("c4_tbar_roll", True)):
        meta, _ = load_golden(case)
        sus, sweep = build_case(meta)
        states, stats = solve_sweep(sus, sweep)
        assert all(s.converged for s in stats)
        design = sus.initial_state()
        pivot = PointRef(Side.CENTER, P.ARB_T_BAR_PIVOT)
        left, right = PointRef(Side.LEFT, P.DROPLINK_T_BAR), PointRef(Side.RIGHT, P.DROPLINK_T_BAR)
        pairs = [(left, right), (left, pivot), (right, pivot)]
        lengths = {pair: _dist(design, *pair) for pair in pairs}
        travel = {side: [] for side in (Side.LEFT, Side.RIGHT)}
        center_x = []
        for st in states:
            center = st.positions[left].data + (st.positions[right].data - st.positions[left].data) / 2.0
            assert abs(center[1]) <= 1e-7                       # crossbar midpoint stays on the centre plane
            for pair in pairs:
                assert abs(_dist(st, *pair) - lengths[pair]) <= 1e-5
            center_x.append(center[0])
            for side in travel:
                wc = PointRef(side, P.WHEEL_CENTER)
                travel[side].append(st.positions[wc].data[2] - design.positions[wc].data[2])
        for side in travel:
            assert abs(min(travel[side]) + 50.0) <= 1e-5 and abs(max(travel[side]) - 50.0) <= 1e-5
        if not opposed:
            assert max(center_x) - min(center_x) > 3.0          # the stem swings through its arc
        else:
            result = compute_sweep_metrics(sus, sweep, states)
            assert result.derivative_error is None
            wanted = {"deriv_t_bar_center_x_wrt_hub_z_left", "deriv_t_bar_center_x_wrt_hub_z_right",
                      "deriv_arb_twist_wrt_hub_z_left", "deriv_arb_twist_wrt_hub_z_right"}
            twists = []
            for row in result.rows:
                assert isinstance(row, AxleMetricRows) and wanted <= row.axle.keys()
                assert all(row.axle[k] is not None for k in wanted) and "t_bar_twist" not in row.axle
                twists.append(row.axle["arb_twist"])
            assert max(twists) - min(twists) > 1.0


def macpherson_invariants():
    meta, _ = load_golden("c2_macpherson_bump_steer")
    sus, sweep = build_case(meta)
    states, stats = solve_sweep(sus, sweep)
    assert all(s.converged and s.max_residual < 1e-3 for s in stats)
    design = sus.initial_state()
    rigid = [(P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_OUTBOARD),
             (P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD),
             (P.TRACKROD_INBOARD, P.TRACKROD_OUTBOARD), (P.AXLE_INBOARD, P.AXLE_OUTBOARD),
             (P.LOWER_WISHBONE_OUTBOARD, P.STRUT_BOTTOM), (P.STRUT_BOTTOM, P.AXLE_OUTBOARD)]
    strut = []
    for st in states:
        for a, b in rigid:
            assert abs(_dist(st, a, b) - _dist(design, a, b)) <= 1e-3
        lbj, top, bottom = (st.positions[k].data for k in (P.LOWER_WISHBONE_OUTBOARD, P.STRUT_TOP, P.STRUT_BOTTOM))
        axis = (top - lbj) / np.linalg.norm(top - lbj)
        off_axis = (bottom - lbj) - ((bottom - lbj) @ axis) * axis
        assert np.linalg.norm(off_axis) <= 1e-3                 # strut bottom stays on the strut axis
        strut.append(float(np.linalg.norm(top - bottom)))
    assert max(strut) - min(strut) > 10.0                       # and the strut telescopes


def rocker_axle_invariants():
    meta, _ = load_golden("c3_rocker_ubar_roll_shipped")
    sus, sweep = build_case(meta)
    states, stats = solve_sweep(sus, sweep)
    assert all(s.converged for s in stats)
    design = sus.initial_state()
    for side in (Side.LEFT, Side.RIGHT):
        link = (PointRef(side, P.DROPLINK_ROCKER), PointRef(side, P.DROPLINK_U_BAR))
        pushrod = (PointRef(side, P.PUSHROD_INBOARD), PointRef(side, P.PUSHROD_OUTBOARD))
        for st in states:
            assert abs(_dist(st, *link) - _dist(design, *link)) <= 1e-5
            assert abs(_dist(st, *pushrod) - _dist(design, *pushrod)) <= 1e-5


def tangents_vs_resolved_states():
    """dq/dt from the device against central differences of neighbouring sweep states (the sweep
    steps are the finite-difference steps), and d(hub z)/d(hub z target) = 1."""
    meta, _ = load_golden("c1_dw_corner_bump")
    sus, sweep = build_case(meta)
    states, _ = solve_sweep(sus, sweep)
    tangents = compute_sweep_tangents(sus, sweep, states)
    assert all(not info.rank_deficient for info in tangents.solve_infos)
    hub = [j for j, dim in enumerate(sweep.target_sweeps) if dim[0].point_id == P.WHEEL_CENTER][0]
    values = [t.value for t in sweep.target_sweeps[hub]]
    for s in range(1, len(states) - 1):
        field = tangents.per_step[s][hub]
        assert abs(field.velocity(P.WHEEL_CENTER)[2] - 1.0) <= 1e-9
        h = values[s + 1] - values[s - 1]
        for key in (P.UPPER_WISHBONE_OUTBOARD, P.TRACKROD_OUTBOARD, P.CONTACT_PATCH_CENTER):
            fd = (states[s + 1].positions[key].data - states[s - 1].positions[key].data) / h
            assert np.abs(fd - field.velocity(key)).max() <= 2e-3   # O(h^2) with 4 mm steps


def known_metric_values_at_design():
    """Hand-checked values of the reference's test geometry at its design pose (reference
    tests/test_metrics.py:321-400): camber -1.909 deg (top tilted inward), caster +4.764 deg (top tilted
    rearward), roadwheel angle 0; all at the reference's tolerance of 1e-3."""
    meta, _ = load_golden("c1_dw_corner_bump")
    sus, _ = build_case(meta)
    row = sus.compute_state_metrics(sus.initial_state())
    assert abs(row["camber"] + 1.909) <= 1e-3 and row["camber"] < 0
    assert abs(row["caster"] - 4.764) <= 1e-3 and row["caster"] > 0
    assert abs(row["roadwheel_angle"]) <= 1e-3
    assert abs(row["wheel_travel"]) <= 1e-6
    assert not any(key.startswith("deriv_") for key in row)      # no tangents given: no derivative columns


BODIES = [known_metric_values_at_design, t_bar_invariants, macpherson_invariants, rocker_axle_invariants, tangents_vs_resolved_states]


@pytest.mark.parametrize("body", BODIES, ids=lambda f: f.__name__)
def test_invariants_on_the_lane_emulation(emu_device, body):
    body()


@pytest.mark.gpu
@pytest.mark.parametrize("body", BODIES, ids=lambda f: f.__name__)
def test_invariants_on_the_device(body):
    body()


def test_tied_derivative_drivers_raise_like_the_reference():
    """The device marks a derivative column whose driver tangents tie with +inf (NaN = the
    reference's None); the row builder turns it into the reference's ValueError
    (metrics/derivatives.py:299-304)."""
    import numpy as np
    import pytest
    from open_kinematics_b200.core.metrics.main import rows_from_columns
    locations = [(None, "camber"), (None, "deriv_camber_wrt_wheel_travel")]
    row = rows_from_columns(np.array([1.5, np.nan]), locations, is_axle=False)
    assert row["camber"] == 1.5 and row["deriv_camber_wrt_wheel_travel"] is None
    with pytest.raises(ValueError, match="Ambiguous derivative driver for column 'deriv_camber_wrt_wheel_travel'"):
        rows_from_columns(np.array([1.5, np.inf]), locations, is_axle=False)


def test_resolve_positions_errors_like_the_reference():
    """presentation.py:304-329: a missing assembly point or a zero-length rotation axis raises."""
    import numpy as np
    import pytest
    from helpers import build_case, load_golden
    from open_kinematics_b200.core.enums import PointID
    from open_kinematics_b200.core.presentation import presentation_points, resolve_positions
    from open_kinematics_b200.core.primitives.point_ref import PointRef, Side
    meta, arr = load_golden("c4_tbar_roll")
    sus, _ = build_case(meta)
    projections, midpoints = presentation_points(sus)
    assert len(projections) == 4 and len(midpoints) == 1
    from helpers import key_from_name
    keys = [key_from_name(n) for n in meta["point_keys"]]
    positions = {k: arr["positions_tight"][0, i] for i, k in enumerate(keys)}
    named = resolve_positions(positions, sus)
    mid = np.array(named["left_droplink_t_bar_right_droplink_t_bar_midpoint"])
    assert np.allclose(mid, 0.5 * (positions[PointRef(Side.LEFT, PointID.DROPLINK_T_BAR)]
                                   + positions[PointRef(Side.RIGHT, PointID.DROPLINK_T_BAR)]))
    broken = dict(positions)
    del broken[PointRef(Side.LEFT, PointID.PUSHROD_INBOARD)]
    with pytest.raises(ValueError, match="Cannot resolve missing assembly points"):
        resolve_positions(broken, sus)
    flat = dict(positions)
    flat[PointRef(Side.LEFT, PointID.ROCKER_AXIS_B)] = flat[PointRef(Side.LEFT, PointID.ROCKER_AXIS_A)]
    with pytest.raises(ValueError, match="zero-length rotation axis"):
        resolve_positions(flat, sus)
