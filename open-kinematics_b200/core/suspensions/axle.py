"""Two-corner axle composition and shared mechanisms.

Behavioural references: axle/suspension.py:78-311 (state merge under
``PointRef`` keys, rack coupling distance, derived-spec remap), axle/mechanisms.py
:230-342 (U-bar), :600-716 (T-bar), :881-899 (rocker-to-rocker heave link: no
constraint, only a pickup rigid to each rocker).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from ..constraints import DistanceConstraint, MidpointOnPlaneConstraint
from ..enums import Axis, PointID, SuspensionType
from ..points.derived.manager import DerivedPointsSpec
from ..primitives.constants import EPS_GEOMETRIC, MIN_CHIRALITY_VOLUME
from ..primitives.geometry import Direction3, Point3
from ..primitives.point_ref import PointRef, Side, side_qualified
from ..state import SuspensionState
from ..targeting import ActuatorDOF, WorldAxisSystem
from .base import Suspension
from .corner import _line_distance, _triple, _unit

P = PointID
_SIDES = (Side.LEFT, Side.RIGHT)


def _dist(a: Point3, b: Point3) -> float:
    return float(np.linalg.norm(b.data - a.data))


@dataclass(frozen=True)
class ArbNone:
    free_points = ()
    output_points = ()

    def validate(self, axle) -> None:
        pass

    def add_to_state(self, state) -> None:
        pass

    def constraints(self, axle) -> list:
        return []


@dataclass(frozen=True)
class ArbUBar:
    """U-bar: each arm pickup holds its distance to the two axis points and to the
    rocker droplink pickup; the arms do not couple the sides (mechanisms.py:307-342)."""

    center_points: dict = field(default_factory=dict)
    droplink_points: dict = field(default_factory=dict)

    free_points = (PointRef(Side.LEFT, P.DROPLINK_U_BAR), PointRef(Side.RIGHT, P.DROPLINK_U_BAR))
    output_points = free_points

    def validate(self, axle) -> None:
        for side, corner in axle.corners.items():
            if P.DROPLINK_ROCKER not in corner.free_points():
                raise ValueError(f"{side.name} U-bar corner does not expose DROPLINK_ROCKER as a moving pickup")
        if set(self.center_points) != {P.ARB_U_BAR_AXIS_A, P.ARB_U_BAR_AXIS_B}:
            raise ValueError("U-bar requires center ARB_U_BAR_AXIS_A and ARB_U_BAR_AXIS_B")
        if set(self.droplink_points) != set(_SIDES):
            raise ValueError("U-bar requires DROPLINK_U_BAR on both sides")
        a = self.center_points[P.ARB_U_BAR_AXIS_A].data
        b = self.center_points[P.ARB_U_BAR_AXIS_B].data
        if np.linalg.norm(b - a) <= EPS_GEOMETRIC:
            raise ValueError("ARB_U_BAR_AXIS_A and ARB_U_BAR_AXIS_B must be distinct points")
        axis = _unit(b - a)
        for side, droplink in self.droplink_points.items():
            if _line_distance(droplink.data, a, axis) <= EPS_GEOMETRIC:
                raise ValueError(f"{side.name} DROPLINK_U_BAR lies on the U-bar axis; it must be off-axis")
            rocker = axle.corners[side].hardpoints[P.DROPLINK_ROCKER].data
            if abs(_triple(b - a, rocker - a, droplink.data - a)) < MIN_CHIRALITY_VOLUME:
                raise ValueError(f"{side.name} U-bar arm geometry does not define reliable handedness")

    def add_to_state(self, state) -> None:
        for point, position in self.center_points.items():
            state.positions[PointRef(Side.CENTER, point)] = position.copy()
        for side, position in self.droplink_points.items():
            key = PointRef(side, P.DROPLINK_U_BAR)
            state.positions[key] = position.copy()
            state.free_points.add(key)

    def constraints(self, axle) -> list:
        axis_a = self.center_points[P.ARB_U_BAR_AXIS_A]
        axis_b = self.center_points[P.ARB_U_BAR_AXIS_B]
        a_key = PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_A)
        b_key = PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_B)
        rows = []
        for side in _SIDES:
            arm = self.droplink_points[side]
            arm_key = PointRef(side, P.DROPLINK_U_BAR)
            rocker = axle.corner_design(side).positions[P.DROPLINK_ROCKER]
            rows += [
                DistanceConstraint(arm_key, a_key, _dist(arm, axis_a)),
                DistanceConstraint(arm_key, b_key, _dist(arm, axis_b)),
                DistanceConstraint(PointRef(side, P.DROPLINK_ROCKER), arm_key, _dist(rocker, arm)),
            ]
        return rows


T_BAR_PIVOT_KEY = PointRef(Side.CENTER, P.ARB_T_BAR_PIVOT)
T_BAR_LEFT_KEY = PointRef(Side.LEFT, P.DROPLINK_T_BAR)
T_BAR_RIGHT_KEY = PointRef(Side.RIGHT, P.DROPLINK_T_BAR)


@dataclass(frozen=True)
class ArbTBar:
    """Rigid T: crossbar ends + chassis pivot form a rigid triangle whose crossbar
    midpoint stays on the vehicle XZ plane (mechanisms.py:669-716)."""

    center_points: dict = field(default_factory=dict)
    droplink_points: dict = field(default_factory=dict)

    free_points = (T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY)
    output_points = free_points

    def validate(self, axle) -> None:
        for side, corner in axle.corners.items():
            if P.DROPLINK_ROCKER not in corner.free_points():
                raise ValueError(f"{side.name} T-bar corner does not expose DROPLINK_ROCKER as a moving pickup")
        if set(self.center_points) != {P.ARB_T_BAR_PIVOT}:
            raise ValueError("T-bar requires center ARB_T_BAR_PIVOT")
        if set(self.droplink_points) != set(_SIDES):
            raise ValueError("T-bar requires DROPLINK_T_BAR on both sides")
        pivot = self.center_points[P.ARB_T_BAR_PIVOT].data
        if abs(pivot[Axis.Y]) > EPS_GEOMETRIC:
            raise ValueError("ARB_T_BAR_PIVOT must lie on the vehicle centerline Y = 0")
        left, right = self.droplink_points[Side.LEFT].data, self.droplink_points[Side.RIGHT].data
        center = left + (right - left) / 2.0
        if abs(center[Axis.Y]) > EPS_GEOMETRIC:
            raise ValueError("The T-bar crossbar midpoint must lie on the vehicle centerline Y = 0")
        crossbar, stem = right - left, center - pivot
        if np.linalg.norm(crossbar) <= EPS_GEOMETRIC:
            raise ValueError("T-bar crossbar points must be distinct")
        if np.linalg.norm(stem) <= EPS_GEOMETRIC:
            raise ValueError("T-bar pivot and crossbar midpoint must be distinct")
        if np.linalg.norm(np.cross(crossbar, stem)) <= EPS_GEOMETRIC:
            raise ValueError("T-bar points must define a non-degenerate triangle")

    def add_to_state(self, state) -> None:
        state.positions[T_BAR_PIVOT_KEY] = self.center_points[P.ARB_T_BAR_PIVOT].copy()
        for side, position in self.droplink_points.items():
            key = PointRef(side, P.DROPLINK_T_BAR)
            state.positions[key] = position.copy()
            state.free_points.add(key)

    def constraints(self, axle) -> list:
        design = axle.design_state()
        g = design.get
        rows = [
            DistanceConstraint(T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY, _dist(g(T_BAR_LEFT_KEY), g(T_BAR_RIGHT_KEY))),
            DistanceConstraint(T_BAR_LEFT_KEY, T_BAR_PIVOT_KEY, _dist(g(T_BAR_LEFT_KEY), g(T_BAR_PIVOT_KEY))),
            DistanceConstraint(T_BAR_RIGHT_KEY, T_BAR_PIVOT_KEY, _dist(g(T_BAR_RIGHT_KEY), g(T_BAR_PIVOT_KEY))),
            MidpointOnPlaneConstraint(T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY,
                                      Point3([0.0, 0.0, 0.0]), Direction3([0.0, 1.0, 0.0])),
        ]
        for side in _SIDES:
            arm_key = PointRef(side, P.DROPLINK_T_BAR)
            rocker_key = PointRef(side, P.DROPLINK_ROCKER)
            rows.append(DistanceConstraint(rocker_key, arm_key, _dist(g(rocker_key), g(arm_key))))
        return rows


@dataclass(frozen=True)
class HeaveLink:
    """kind: 'none' | 'rocker_to_rocker' (variable-length link: no constraint row)."""

    kind: str = "none"

    def validate(self, axle) -> None:
        if self.kind != "rocker_to_rocker":
            return
        for side, corner in axle.corners.items():
            if P.HEAVE_LINK_ROCKER not in corner.free_points():
                raise ValueError(f"{side.name} corner does not expose HEAVE_LINK_ROCKER as a moving pickup")
        # authored pose: the separation check does not need the (device-computed) setup pose
        left = axle.corners[Side.LEFT].authored_state().get(P.HEAVE_LINK_ROCKER)
        right = axle.corners[Side.RIGHT].authored_state().get(P.HEAVE_LINK_ROCKER)
        if _dist(left, right) <= EPS_GEOMETRIC:
            raise ValueError("Rocker-to-rocker heave-link pickups must be separated in the design state")


@dataclass
class AxleSuspension(Suspension):
    type_key: SuspensionType = SuspensionType.DOUBLE_WISHBONE
    corners: dict = field(default_factory=dict)
    anti_roll: object = field(default_factory=ArbNone)
    heave_link: HeaveLink = field(default_factory=HeaveLink)
    name: str = "unnamed"
    version: str = "0.0.0"
    config: object = None
    side: Side = Side.CENTER
    hardpoints: dict = field(default_factory=dict)
    _initial_state: SuspensionState | None = field(default=None, init=False, repr=False)
    _authored_mode: bool = field(default=False, init=False, repr=False)

    def __post_init__(self) -> None:
        self.validate_hardpoints()

    @property
    def is_axle(self) -> bool:
        return True

    # -- design pose access; in "authored mode" (structure()) no setup shim is applied --------
    def corner_design(self, side: Side) -> SuspensionState:
        corner = self.corners[side]
        return corner.authored_state() if self._authored_mode else corner.initial_state()

    def design_state(self) -> SuspensionState:
        return self._merged_state() if self._authored_mode else self.initial_state()

    def structure(self) -> tuple:
        self._authored_mode = True
        try:
            return self._merged_state(), self.constraints()
        finally:
            self._authored_mode = False

    def _merged_state(self) -> SuspensionState:
        positions: dict = {}
        free: set = set()
        for side in self.corners:
            cs = self.corner_design(side)
            positions.update({PointRef(side, k): p.copy() for k, p in cs.positions.items()})
            free.update(PointRef(side, k) for k in cs.free_points)
        state = SuspensionState(positions, free)
        self.anti_roll.add_to_state(state)
        state.free_points_order = sorted(state.free_points)
        return state

    def reported_type_key(self) -> SuspensionType:
        return self.type_key

    def validate_hardpoints(self) -> None:
        if set(self.corners) != set(_SIDES):
            raise ValueError("Axle requires exactly LEFT and RIGHT corner models.")
        for side, corner in self.corners.items():
            if corner.side is not side:
                raise ValueError(f"Axle {side.name.lower()} corner must declare side '{side.name.lower()}'.")
            corner.validate_hardpoints()
        self.rack_attachment_points()
        self.anti_roll.validate(self)
        self.heave_link.validate(self)

    def rack_attachment_points(self):
        left = self.corners[Side.LEFT].rack_attachment_point()
        right = self.corners[Side.RIGHT].rack_attachment_point()
        if (left is None) != (right is None):
            raise ValueError("Axle corners disagree on rack attachment: one corner is steered and the other is not.")
        return None if left is None else (left, right)

    def actuator_dofs(self) -> tuple:
        rack = self.rack_attachment_points()
        if rack is None:
            return ()
        return (ActuatorDOF(
            name="steering rack",
            point_keys=(PointRef(Side.LEFT, rack[0]), PointRef(Side.RIGHT, rack[1])),
            direction=WorldAxisSystem.Y,
        ),)

    def initial_state(self) -> SuspensionState:
        if self._initial_state is None:
            self._initial_state = self._merged_state()
        return self._initial_state

    def free_points(self) -> tuple:
        corner_points = tuple(PointRef(s, k) for s, c in self.corners.items() for k in c.free_points())
        return (*corner_points, *self.anti_roll.free_points)

    def output_points(self) -> tuple:
        corner_points = tuple(side_qualified(s, k) for s in _SIDES for k in self.corners[s].output_points())
        return tuple(dict.fromkeys((*corner_points, *self.anti_roll.output_points)))

    def constraints(self) -> list:
        rows = [
            c.remap(lambda k, side=side: side_qualified(side, k))
            for side, corner in self.corners.items()
            for c in corner.constraints_at(self.corner_design(side).positions)
        ]
        rack = self.rack_attachment_points()
        if rack is not None:
            left = self.corner_design(Side.LEFT).positions[rack[0]]
            right = self.corner_design(Side.RIGHT).positions[rack[1]]
            # The rigid rack keeps its two ends a fixed distance apart (suspension.py:196-209).
            rows.append(DistanceConstraint(PointRef(Side.LEFT, rack[0]), PointRef(Side.RIGHT, rack[1]),
                                           _dist(left, right)))
        rows += self.anti_roll.constraints(self)
        return rows

    def topology_diagnostics(self, states: list) -> list:
        """Corner-owned diagnostics followed by the shared axle checks (reference
        axle/suspension.py:241-251).  Checks of the shipped architectures run in the device's
        diagnostic program (base class); a user-defined corner that overrides
        ``topology_diagnostics`` is asked directly, with its own side's states."""
        issues: list = []
        for side in _SIDES:
            corner = self.corners[side]
            if type(corner).topology_diagnostics is not Suspension.topology_diagnostics:
                issues.extend(corner.topology_diagnostics([self.corner_state(state, side) for state in states]))
        issues.extend(super().topology_diagnostics(states))
        return issues

    def derived_spec(self) -> DerivedPointsSpec:
        functions: dict = {}
        dependencies: dict = {}
        for side, corner in self.corners.items():
            spec = corner.derived_spec()
            for key, fn in spec.functions.items():
                functions[PointRef(side, key)] = fn.remap(lambda k, side=side: PointRef(side, k))
            for key, deps in spec.dependencies.items():
                dependencies[PointRef(side, key)] = {PointRef(side, d) for d in deps}
        return DerivedPointsSpec(functions, dependencies)

    def corner_state(self, state: SuspensionState, side: Side) -> SuspensionState:
        positions = {k.point: p for k, p in state.positions.items() if isinstance(k, PointRef) and k.side is side}
        free = {k.point for k in state.free_points if isinstance(k, PointRef) and k.side is side}
        return SuspensionState(positions, free)

    def resolve_target_key(self, point: PointID, side: Side | None):
        if side not in _SIDES:
            raise ValueError(f"Axle sweep target for '{point.name}' requires side left or right.")
        return PointRef(side, point)
