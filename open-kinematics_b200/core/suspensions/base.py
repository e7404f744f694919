"""Suspension model base classes (reference core/suspensions/base.py:36-253,
corner/base.py:20-101).  Only the solve-path surface is mirrored: point sets,
initial state, constraints, derived spec, actuator DOFs and target resolution."""

from __future__ import annotations

from typing import Sequence

from ..enums import PointID, SuspensionType
from ..points.derived.manager import DerivedPointsSpec
from ..primitives.point_ref import PointRef, Side
from ..state import SuspensionState
from ..targeting import ActuatorDOF, WorldAxisSystem


class Suspension:
    name: str = "unnamed"
    version: str = "0.0.0"
    side: Side
    config = None
    hardpoints: dict

    @property
    def is_axle(self) -> bool:
        return False

    def reported_type_key(self) -> SuspensionType:
        raise NotImplementedError

    def initial_state(self) -> SuspensionState:
        raise NotImplementedError

    def free_points(self) -> Sequence:
        raise NotImplementedError

    def constraints(self) -> list:
        raise NotImplementedError

    def derived_spec(self) -> DerivedPointsSpec:
        raise NotImplementedError

    def output_points(self) -> tuple:
        raise NotImplementedError

    def actuator_dofs(self) -> tuple:
        return ()

    def damper_points(self):
        return None

    def get_hardpoints_copy(self) -> dict:
        return {k: p.copy() for k, p in self.hardpoints.items()}

    def resolve_target_key(self, point: PointID, side: Side | None):
        if side is not None:
            raise ValueError(
                f"Sweep target for '{point.name}' specifies side '{side.name.lower()}', but "
                f"suspension type '{self.reported_type_key()}' is a single corner and does "
                "not accept a side."
            )
        return point

    def structure(self) -> tuple:
        """``(state, constraints)`` describing the topology without touching the device: the
        authored pose stands in for the design pose (a setup shim would need the device), which
        is all the topology compiler needs when design constants are recomputed per instance."""
        return self.initial_state(), self.constraints()

    def default_state_targets(self) -> list:
        """Targets that pin every degree of freedom of the mechanism at a state: hub height of each
        wheel plus every physical actuator (used to re-pin a state when no tangents are given)."""
        from ..enums import Axis
        from ..targeting import PointTarget, PointTargetAxis, PointTargetVector
        hubs = [PointRef(side, PointID.WHEEL_CENTER) for side in (Side.LEFT, Side.RIGHT)] if self.is_axle \
            else [PointID.WHEEL_CENTER]
        targets = [PointTarget(hub, PointTargetAxis(Axis.Z), 0.0) for hub in hubs]
        targets += [PointTarget(dof.point_keys[0], PointTargetVector(dof.direction), 0.0)
                    for dof in self.actuator_dofs()]
        return targets

    def compute_state_metrics(self, state: SuspensionState, tangents=None):
        """Metric row(s) of one solved state (reference suspensions/base.py:198-204,
        corner/base.py:93-101, axle/suspension.py:313-326): a ``MetricRow`` for a corner,
        ``AxleMetricRows`` for an axle.  Derivative columns are present only when ``tangents``
        (the state's ``TangentField`` list) is given; their targets select the drivers.  The
        numbers come from the device's metric program evaluated at the state."""
        if self.config is None:
            raise ValueError("Suspension has no configuration")
        from ..metrics.main import rows_from_columns
        from ..sweep import evaluate_states_on_device
        heads = [field.target for field in tangents] if tangents else self.default_state_targets()
        res, solver, ramp = evaluate_states_on_device(self, heads, [state], want_metrics=True, ramp=5)
        return rows_from_columns(res.metrics[0, ramp], solver.program.metric_locations, self.is_axle,
                                 with_derivatives=bool(tangents))

    def topology_diagnostics(self, states: list) -> list:
        """Advisory checks owned by the concrete topology (reference suspensions/base.py:220-225,
        axle/suspension.py:241-251), evaluated by the device's diagnostic program."""
        from ..diagnostics import evaluate_diagnostics_on_device, topology_issues
        if not states:
            return []
        diag, _, program = evaluate_diagnostics_on_device(self, states)
        return topology_issues(program.diagnostic_checks, diag)

    def all_point_keys(self) -> set:
        """Every point present in a solved state (authored + derived)."""
        return set(self.initial_state().positions)


class CornerSuspension(Suspension):
    TYPE_KEY: SuspensionType
    REQUIRED_POINTS: frozenset = frozenset()

    def reported_type_key(self) -> SuspensionType:
        return self.TYPE_KEY

    def required_points(self) -> frozenset:
        return self.REQUIRED_POINTS

    def validate_hardpoints(self) -> None:
        present = set(self.hardpoints)
        missing = self.required_points() - present
        if missing:
            raise ValueError("Missing required hardpoints: " + ", ".join(sorted(p.name for p in missing)))
        unknown = present - self.required_points()
        if unknown:
            raise ValueError("Invalid hardpoints: " + ", ".join(sorted(p.name for p in unknown)))

    # -- role hooks: what the axle composer, the metric program and the diagnostics ask of ANY corner
    # (reference corner/base.py:20-101); user-defined corners override them -----------------------
    def wheel_axis_points(self) -> tuple:
        return (PointID.AXLE_INBOARD, PointID.AXLE_OUTBOARD)

    def authored_state(self) -> SuspensionState:
        """Design pose before any setup shim; a corner without shims has only one pose."""
        return self.initial_state()

    def constraints_at(self, positions) -> list:
        """Constraints with their design constants taken at ``positions``.  Corners that compute
        their constants from their own initial state need not override this."""
        return self.constraints()

    def instant_center_points(self):
        """``(kind, points)`` of the planes whose intersection is the instant axis, or None when the
        architecture declares no instant centres (then the side- / front-view instant-centre and
        anti-geometry columns are None, as for the reference's compute_*_instant_center -> None)."""
        return None

    def steering_axis_points(self) -> tuple:
        raise NotImplementedError

    def rack_attachment_point(self):
        raise NotImplementedError

    def actuator_dofs(self) -> tuple:
        rack = self.rack_attachment_point()
        if rack is None:
            return ()
        return (ActuatorDOF(name="steering rack", point_keys=(rack,), direction=WorldAxisSystem.Y),)
