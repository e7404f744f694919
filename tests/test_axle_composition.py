"""The axle composer accepts any CornerSuspension, not one architecture (reference
tests/test_axle_composition.py:58-339): a minimal trailing-arm corner without AXLE_INBOARD /
AXLE_OUTBOARD points is composed into AxleSuspension and driven through composition, solving,
metrics and diagnostics, so every metric path must resolve the wheel axis through the role hooks.
CPU: the lane emulation stands in for the device; ``-m gpu`` runs the same body on the library."""

from dataclasses import dataclass, field

import numpy as np
import pytest

from helpers import build_case, load_golden
from open_kinematics_b200.core.constraints import DistanceConstraint
from open_kinematics_b200.core.diagnostics import DiagnosticCategory, DiagnosticIssue, DiagnosticSeverity
from open_kinematics_b200.core.enums import Axis, PointID, SteeringType, SuspensionType
from open_kinematics_b200.core.metrics.main import AxleMetricRows
from open_kinematics_b200.core.points.derived.manager import DerivedPointsSpec
from open_kinematics_b200.core.primitives.geometry import Point3
from open_kinematics_b200.core.primitives.point_ref import PointRef, Side
from open_kinematics_b200.core.state import SuspensionState
from open_kinematics_b200.core.suspensions.axle import AxleSuspension
from open_kinematics_b200.core.suspensions.base import CornerSuspension
from open_kinematics_b200.core.sweep import compute_sweep_metrics, solve_sweep
from open_kinematics_b200.core.targeting import PointTarget, PointTargetAxis, SweepConfig

CHASSIS_FRONT = PointID.LOWER_WISHBONE_INBOARD_FRONT
CHASSIS_REAR = PointID.LOWER_WISHBONE_INBOARD_REAR
KNUCKLE = PointID.LOWER_WISHBONE_OUTBOARD
STUB_DIAGNOSTIC_MESSAGE = "stub corner diagnostic"


@dataclass
class TrailingArmCorner(CornerSuspension):
    """Rigid knuckle body hinged on two chassis anchors; deliberately not a double wishbone."""

    name: str = "stub"
    side: Side = Side.LEFT
    config: object = None
    hardpoints: dict = field(default_factory=dict)
    TYPE_KEY = SuspensionType.DOUBLE_WISHBONE
    REQUIRED_POINTS = frozenset({CHASSIS_FRONT, CHASSIS_REAR, KNUCKLE, PointID.WHEEL_CENTER,
                                 PointID.CONTACT_PATCH_CENTER})
    FREE_POINTS = (KNUCKLE, PointID.WHEEL_CENTER, PointID.CONTACT_PATCH_CENTER)
    OUTPUT_POINTS = FREE_POINTS

    def initial_state(self) -> SuspensionState:
        return SuspensionState(positions=self.get_hardpoints_copy(), free_points=set(self.FREE_POINTS))

    def free_points(self):
        return self.FREE_POINTS

    def output_points(self):
        return self.OUTPUT_POINTS

    def constraints(self) -> list:
        positions = self.initial_state().positions

        def distance(a, b):
            return DistanceConstraint(a, b, float(np.linalg.norm(positions[a].data - positions[b].data)))

        rows = [distance(moving, anchor) for moving in self.FREE_POINTS for anchor in (CHASSIS_FRONT, CHASSIS_REAR)]
        rows += [distance(KNUCKLE, PointID.WHEEL_CENTER), distance(PointID.WHEEL_CENTER, PointID.CONTACT_PATCH_CENTER),
                 distance(KNUCKLE, PointID.CONTACT_PATCH_CENTER)]
        return rows

    def derived_spec(self) -> DerivedPointsSpec:
        return DerivedPointsSpec({}, {})

    def wheel_axis_points(self):
        return (KNUCKLE, PointID.WHEEL_CENTER)      # not the double-wishbone convention

    def steering_axis_points(self):
        return (KNUCKLE, CHASSIS_FRONT)

    def rack_attachment_point(self):
        return None

    def topology_diagnostics(self, states):
        return [DiagnosticIssue(None, DiagnosticCategory.CHIRALITY, DiagnosticSeverity.WARNING,
                                STUB_DIAGNOSTIC_MESSAGE, None)]


@dataclass
class SteeredTrailingArmCorner(TrailingArmCorner):
    def rack_attachment_point(self):
        return PointID.TRACKROD_INBOARD


def build_stub_corner(side, corner_class=TrailingArmCorner, config=None):
    lateral = 600.0 if side is Side.LEFT else -600.0
    return corner_class(name=f"stub_{side.name.lower()}", side=side, config=config, hardpoints={
        CHASSIS_FRONT: Point3(np.array([100.0, 0.3 * lateral, 150.0])),
        CHASSIS_REAR: Point3(np.array([-100.0, 0.3 * lateral, 150.0])),
        KNUCKLE: Point3(np.array([0.0, 0.9 * lateral, 50.0])),
        PointID.WHEEL_CENTER: Point3(np.array([0.0, lateral, 0.0])),
        PointID.CONTACT_PATCH_CENTER: Point3(np.array([0.0, lateral, -200.0])),
    })


def donor_config():
    """Vehicle configuration of the shipped corner geometry with the steering switched off."""
    import dataclasses
    meta, _ = load_golden("c1_dw_corner_bump")
    sus, _ = build_case(meta)
    cfg = sus.config
    steering = dataclasses.replace(cfg.steering, type=SteeringType.NONE)
    return dataclasses.replace(cfg, steering=steering)


def build_stub_axle(config=None):
    return AxleSuspension(type_key=SuspensionType.DOUBLE_WISHBONE, name="stub_axle", side=Side.CENTER, hardpoints={},
                          config=config, corners={Side.LEFT: build_stub_corner(Side.LEFT, config=config),
                                                  Side.RIGHT: build_stub_corner(Side.RIGHT, config=config)})


def test_axle_composes_non_double_wishbone_corners():
    axle = build_stub_axle()
    state = axle.initial_state()
    assert state.free_points == {PointRef(side, point) for side in (Side.LEFT, Side.RIGHT)
                                 for point in TrailingArmCorner.FREE_POINTS}
    constraints = axle.constraints()
    assert len(constraints) == 2 * len(build_stub_corner(Side.LEFT).constraints())
    assert all(isinstance(p, PointRef) for c in constraints for p in c.involved_points)
    assert axle.rack_attachment_points() is None


def test_axle_rejects_mixed_rack_attachment():
    with pytest.raises(ValueError, match="disagree on rack attachment"):
        AxleSuspension(type_key=SuspensionType.DOUBLE_WISHBONE, name="stub_axle", side=Side.CENTER, hardpoints={},
                       corners={Side.LEFT: build_stub_corner(Side.LEFT, SteeredTrailingArmCorner),
                                Side.RIGHT: build_stub_corner(Side.RIGHT)})


def stub_axle_solves_and_reports_metrics_through_role_hooks():
    axle = build_stub_axle(config=donor_config())
    bump_values = [-10.0, 0.0, 10.0]
    sweep = SweepConfig([[PointTarget(PointRef(side, PointID.WHEEL_CENTER), PointTargetAxis(Axis.Z), v)
                          for v in bump_values] for side in (Side.LEFT, Side.RIGHT)])
    states, infos = solve_sweep(axle, sweep)
    assert all(info.converged for info in infos) and all(info.max_residual < 1e-6 for info in infos)
    # the arm is a rigid body about the chassis hinge: every authored distance is preserved
    for state in states:
        for c in axle.constraints():
            a, b = (state.positions[k].data for k in c.point_keys)
            assert abs(np.linalg.norm(a - b) - c.target_distance) < 1e-5
    metrics = compute_sweep_metrics(axle, sweep, states)
    assert metrics.derivative_error is None
    final = metrics.rows[-1]
    assert isinstance(final, AxleMetricRows)
    for side in (Side.LEFT, Side.RIGHT):
        row = final.corners[side]
        assert row["camber"] is not None and row["caster"] is not None
        assert row["deriv_camber_wrt_hub_z"] is not None
        assert "deriv_roadwheel_angle_wrt_rack_displacement" not in row     # no rack, no rack-driven derivatives
        assert row["svic_x"] is None and row["fvic_y"] is None                # the stub declares no instant centres
    assert final.axle["heave"] == pytest.approx(bump_values[-1], abs=1e-6)
    assert final.axle["rack_displacement"] is None
    stub_issues = [i for i in axle.topology_diagnostics(states) if i.message == STUB_DIAGNOSTIC_MESSAGE]
    assert len(stub_issues) == 2           # corner-owned diagnostics survive axle composition
    assert axle.reported_type_key() is SuspensionType.DOUBLE_WISHBONE


def test_stub_axle_emu(emu_device):
    stub_axle_solves_and_reports_metrics_through_role_hooks()


@pytest.mark.gpu
def test_stub_axle_gpu():
    stub_axle_solves_and_reports_metrics_through_role_hooks()
