"""Suspension model base classes (reference core/suspensions/base.py:36-253,
corner/base.py:20-101).  Only the solve-path surface is mirrored: point sets,
initial state, constraints, derived spec, actuator DOFs and target resolution."""

from __future__ import annotations

from typing import Sequence

from ..enums import PointID, SuspensionType
from ..points.derived.manager import DerivedPointsSpec
from ..primitives.point_ref import Side
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

    def wheel_axis_points(self) -> tuple:
        return (PointID.AXLE_INBOARD, PointID.AXLE_OUTBOARD)

    def steering_axis_points(self) -> tuple:
        raise NotImplementedError

    def rack_attachment_point(self):
        raise NotImplementedError

    def actuator_dofs(self) -> tuple:
        rack = self.rack_attachment_point()
        if rack is None:
            return ()
        return (ActuatorDOF(name="steering rack", point_keys=(rack,), direction=WorldAxisSystem.Y),)
