"""Corner topologies and their composed mechanisms.

Behavioural references (constraint lists, free points, derived points):
  double wishbone   corner/double_wishbone.py:74-350
  MacPherson        corner/macpherson.py:83-323
  track rod/toe link corner/track_rod.py:27-97, corner/toe_link.py:22-87
  rigid attachments corner/attachments.py:23-120
  actuation/springs corner/mechanisms.py:78-314, :437-623

Every design constant is the constraint's own geometric quantity evaluated at
the design pose (true norm / angle / signed volume), which is what lets the
device recompute them per perturbed instance.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from ..constraints import (
    AngleConstraint,
    Constraint,
    DistanceConstraint,
    PointOnLineConstraint,
    ScalarTripleProductConstraint,
)
from ..enums import MountBody, PointID, ShimType, SteeringType, SuspensionType
from ..points.derived.definitions import PointAlongLine, build_wheel_derived_spec
from ..points.derived.manager import DerivedPointsManager, DerivedPointsSpec
from ..primitives.constants import EPS_GEOMETRIC, MIN_CHIRALITY_VOLUME
from ..primitives.point_ref import Side
from ..state import SuspensionState
from ..targeting import WorldAxisSystem
from .base import CornerSuspension

P = PointID


# ---------------------------------------------------------------------------
# Design-pose measurements (true norms, no softnorm: reference geometric.py:17-28,
# :71-104, :197-214).
# ---------------------------------------------------------------------------
def _d(positions, a, b) -> float:
    return float(np.linalg.norm(positions[b].data - positions[a].data))


def _unit(v: np.ndarray) -> np.ndarray:
    n = float(np.linalg.norm(v))
    if n < EPS_GEOMETRIC:
        raise ValueError("Cannot normalize a zero-length vector")
    return v / n


def _angle(v1: np.ndarray, v2: np.ndarray) -> float:
    u1, u2 = _unit(v1), _unit(v2)
    return float(np.arctan2(np.linalg.norm(np.cross(u1, u2)), np.dot(u1, u2)))


def _triple(v1, v2, v3) -> float:
    return float(np.dot(v1, np.cross(v2, v3)))


def _line_distance(p, a, direction) -> float:
    return float(np.linalg.norm(np.cross(p - a, direction)))


def distance_constraint(positions, a, b) -> DistanceConstraint:
    return DistanceConstraint(a, b, _d(positions, a, b))


def rigid_point_constraints(positions, point, references) -> list:
    """Three design-length distances to a body (attachments.py:23-42)."""
    return [distance_constraint(positions, point, ref) for ref in references]


def chiral_rigid_point_constraints(positions, point, references) -> list:
    """Rigid attachment plus a signed-volume row preserving handedness
    (attachments.py:45-74)."""
    out = rigid_point_constraints(positions, point, references)
    a, b, c = (positions[r].data for r in references)
    volume = _triple(b - a, c - a, positions[point].data - a)
    if abs(volume) < MIN_CHIRALITY_VOLUME:
        raise ValueError(f"{point.name} and its rigid-body references do not define reliable handedness")
    out.append(ScalarTripleProductConstraint(*references, point, target_volume=volume, scale=abs(volume)))
    return out


def anchored_rigid_point_constraints(positions, point, anchors) -> list:
    """First three anchors chiral, further anchors plain distances (attachments.py:97-120)."""
    out = chiral_rigid_point_constraints(positions, point, tuple(anchors[:3]))
    out.extend(distance_constraint(positions, point, a) for a in anchors[3:])
    return out


def validate_rigid_anchor_points(hardpoints, anchors, label: str) -> None:
    if len(anchors) < 3:
        raise ValueError(f"{label} requires at least three mounting body anchors")
    a, b, c = (hardpoints[k].data for k in anchors[:3])
    if np.linalg.norm(b - a) <= EPS_GEOMETRIC:
        raise ValueError(f"{label} mounting body anchors must be distinct")
    if _line_distance(c, a, _unit(b - a)) <= EPS_GEOMETRIC:
        raise ValueError(f"The first three {label} mounting body anchors must not be collinear")


# ---------------------------------------------------------------------------
# Wheel-heading links.
# ---------------------------------------------------------------------------
@dataclass(frozen=True)
class HeadingLink:
    """Track rod (rack driven, inboard end slides on a world-Y line) or fixed toe link."""

    steered: bool
    upright_anchors: tuple
    preserve_attachment_handedness: bool = True

    @property
    def inboard_point(self):
        return P.TRACKROD_INBOARD if self.steered else P.TOE_LINK_INBOARD

    @property
    def outboard_point(self):
        return P.TRACKROD_OUTBOARD if self.steered else P.TOE_LINK_OUTBOARD

    @property
    def required_points(self) -> frozenset:
        return frozenset({self.inboard_point, self.outboard_point})

    @property
    def output_points(self) -> tuple:
        return (self.inboard_point, self.outboard_point)

    @property
    def free_points(self) -> tuple:
        return (self.outboard_point, self.inboard_point) if self.steered else (self.outboard_point,)

    def validate(self, hardpoints) -> None:
        validate_rigid_anchor_points(hardpoints, self.upright_anchors, "Track rod" if self.steered else "Toe link")

    def constraints(self, positions) -> list:
        out = self.outboard_point
        if self.preserve_attachment_handedness:
            attach = anchored_rigid_point_constraints(positions, out, self.upright_anchors)
        else:
            attach = [distance_constraint(positions, out, a) for a in self.upright_anchors]
        rows: list = [distance_constraint(positions, self.inboard_point, out), *attach]
        if self.steered:
            rows.append(
                PointOnLineConstraint(
                    point_id=self.inboard_point,
                    line_point=positions[self.inboard_point],
                    line_direction=WorldAxisSystem.Y,
                )
            )
        return rows


# ---------------------------------------------------------------------------
# Actuation and spring mechanisms of a double-wishbone corner.
# ---------------------------------------------------------------------------
ROCKER_BODY = (P.ROCKER_AXIS_A, P.ROCKER_AXIS_B, P.PUSHROD_INBOARD)


@dataclass(frozen=True)
class ActuationDirect:
    spring_pickup_body: tuple

    moving_pickup_point = P.STRUT_BOTTOM
    required_points = frozenset()
    free_points = ()
    output_points = ()
    torsion_axis = None

    @property
    def moving_pickup_body(self):
        return self.spring_pickup_body

    def validate(self, hardpoints) -> None:
        validate_rigid_anchor_points(hardpoints, self.spring_pickup_body, "Direct spring actuation")

    def constraints(self, positions) -> list:
        return []

    def spring_constraints(self, positions) -> list:
        return anchored_rigid_point_constraints(positions, P.STRUT_BOTTOM, self.spring_pickup_body)


@dataclass(frozen=True)
class ActuationPushrodRocker:
    pushrod_outboard_body: tuple
    external_point_ids: tuple = ()

    moving_pickup_point = P.PUSHROD_OUTBOARD
    torsion_axis = (P.ROCKER_AXIS_A, P.ROCKER_AXIS_B)

    @property
    def moving_pickup_body(self):
        return self.pushrod_outboard_body

    @property
    def rocker_mounted_point_ids(self) -> tuple:
        return (P.PUSHROD_INBOARD, *self.external_point_ids)

    @property
    def required_points(self) -> frozenset:
        return frozenset({P.PUSHROD_OUTBOARD, P.PUSHROD_INBOARD, P.ROCKER_AXIS_A, P.ROCKER_AXIS_B,
                          *self.external_point_ids})

    @property
    def free_points(self) -> tuple:
        return (P.PUSHROD_OUTBOARD, P.PUSHROD_INBOARD, *self.external_point_ids)

    output_points = free_points

    def validate(self, hardpoints) -> None:
        validate_rigid_anchor_points(hardpoints, self.pushrod_outboard_body, "Pushrod actuation")
        a, b = hardpoints[P.ROCKER_AXIS_A].data, hardpoints[P.ROCKER_AXIS_B].data
        if np.linalg.norm(b - a) <= EPS_GEOMETRIC:
            raise ValueError("Rocker axis points must be distinct")
        axis = _unit(b - a)
        for point in self.rocker_mounted_point_ids:
            if _line_distance(hardpoints[point].data, a, axis) <= EPS_GEOMETRIC:
                raise ValueError(f"{point.name} must not lie on the rocker axis")

    def constraints(self, positions) -> list:
        rows = anchored_rigid_point_constraints(positions, P.PUSHROD_OUTBOARD, self.pushrod_outboard_body)
        rows += [
            distance_constraint(positions, P.PUSHROD_OUTBOARD, P.PUSHROD_INBOARD),
            distance_constraint(positions, P.PUSHROD_INBOARD, P.ROCKER_AXIS_A),
            distance_constraint(positions, P.PUSHROD_INBOARD, P.ROCKER_AXIS_B),
        ]
        for point in self.external_point_ids:
            rows += chiral_rigid_point_constraints(positions, point, ROCKER_BODY)
        return rows

    def spring_constraints(self, positions) -> list:
        return chiral_rigid_point_constraints(positions, P.STRUT_BOTTOM, ROCKER_BODY)


@dataclass(frozen=True)
class CornerSpring:
    """kind: 'none' | 'coilover' | 'torsion_bar' (corner/mechanisms.py:437-623)."""

    kind: str = "none"

    @property
    def required_points(self) -> frozenset:
        return frozenset({P.STRUT_TOP, P.STRUT_BOTTOM}) if self.kind == "coilover" else frozenset()

    @property
    def free_points(self) -> tuple:
        return (P.STRUT_BOTTOM,) if self.kind == "coilover" else ()

    @property
    def output_points(self) -> tuple:
        return (P.STRUT_TOP, P.STRUT_BOTTOM) if self.kind == "coilover" else ()

    @property
    def damper_points(self):
        return (P.STRUT_TOP, P.STRUT_BOTTOM) if self.kind == "coilover" else None

    @property
    def rocker_mounted_points(self) -> tuple:
        return (P.STRUT_BOTTOM,) if self.kind == "coilover" else ()

    def validate(self, actuation) -> None:
        if self.kind == "torsion_bar" and actuation.torsion_axis is None:
            raise ValueError("Corner torsion bar is not supported by direct actuation yet")

    def constraints(self, positions, actuation) -> list:
        return actuation.spring_constraints(positions) if self.kind == "coilover" else []


# ---------------------------------------------------------------------------
# Corners.
# ---------------------------------------------------------------------------
def _validate_side_sign(hardpoints: dict, side: Side) -> None:
    """AXLE_OUTBOARD.y > 0 on the left, < 0 on the right (build.py:326-341)."""
    ao = hardpoints.get(P.AXLE_OUTBOARD)
    if ao is None:
        return
    y = float(ao.data[1])
    if side is Side.LEFT and y <= 0.0:
        raise ValueError(f"Side 'left' requires AXLE_OUTBOARD Y > 0 (got {y}); check the hardpoint handedness.")
    if side is Side.RIGHT and y >= 0.0:
        raise ValueError(f"Side 'right' requires AXLE_OUTBOARD Y < 0 (got {y}); check the hardpoint handedness.")


@dataclass
class _Corner(CornerSuspension):
    hardpoints: dict = field(default_factory=dict)
    config: object = None
    side: Side = Side.LEFT
    name: str = "unnamed"
    version: str = "0.0.0"
    _initial_state: SuspensionState | None = field(default=None, init=False, repr=False)

    def _heading_link(self, anchors, chiral: bool) -> HeadingLink:
        if self.config is None:
            raise ValueError(f"{type(self).__name__} requires configuration")
        return HeadingLink(
            steered=self.config.steering.type is SteeringType.RACK,
            upright_anchors=anchors,
            preserve_attachment_handedness=chiral,
        )

    def rack_attachment_point(self):
        link = self.wheel_heading_link
        return link.inboard_point if link.steered else None

    def output_points(self) -> tuple:
        raise NotImplementedError


@dataclass
class DoubleWishboneSuspension(_Corner):
    TYPE_KEY = SuspensionType.DOUBLE_WISHBONE
    REQUIRED_POINTS = frozenset({
        P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD,
        P.UPPER_WISHBONE_INBOARD_FRONT, P.UPPER_WISHBONE_INBOARD_REAR, P.UPPER_WISHBONE_OUTBOARD,
        P.AXLE_INBOARD, P.AXLE_OUTBOARD,
    })
    LOWER_WISHBONE_BODY = (P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD)
    UPRIGHT_BODY = (P.UPPER_WISHBONE_OUTBOARD, P.LOWER_WISHBONE_OUTBOARD, P.AXLE_INBOARD, P.AXLE_OUTBOARD)
    MOUNT_BODIES = {MountBody.LOWER_WISHBONE: LOWER_WISHBONE_BODY, MountBody.UPRIGHT: UPRIGHT_BODY}
    SUPPORTED_SHIMS = frozenset({ShimType.OUTBOARD_CAMBER})
    LOCATING_OUTPUT_POINTS = (
        P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD,
        P.UPPER_WISHBONE_INBOARD_FRONT, P.UPPER_WISHBONE_INBOARD_REAR, P.UPPER_WISHBONE_OUTBOARD,
    )
    WHEEL_OUTPUT_POINTS = (
        P.AXLE_INBOARD, P.AXLE_OUTBOARD, P.AXLE_MIDPOINT, P.WHEEL_CENTER, P.WHEEL_INBOARD,
        P.WHEEL_OUTBOARD, P.CONTACT_PATCH_CENTER,
    )
    FREE_POINTS = (P.UPPER_WISHBONE_OUTBOARD, P.LOWER_WISHBONE_OUTBOARD, P.AXLE_INBOARD, P.AXLE_OUTBOARD)

    actuation: object = None
    spring: CornerSpring = field(default_factory=CornerSpring)

    def __post_init__(self) -> None:
        # Four upright anchors already over-determine the attachment; the upright
        # angle row keeps the authored branch (double_wishbone.py:166-177).
        self.wheel_heading_link = self._heading_link(self.UPRIGHT_BODY, chiral=False)
        if self.actuation is None:
            self.actuation = ActuationDirect(spring_pickup_body=self.LOWER_WISHBONE_BODY)
        self.validate_hardpoints()

    def required_points(self) -> frozenset:
        return (self.REQUIRED_POINTS | self.wheel_heading_link.required_points
                | self.actuation.required_points | self.spring.required_points)

    def validate_hardpoints(self) -> None:
        super().validate_hardpoints()
        self.wheel_heading_link.validate(self.hardpoints)
        self.actuation.validate(self.hardpoints)
        self.spring.validate(self.actuation)

    def free_points(self) -> tuple:
        return (*self.FREE_POINTS, *self.wheel_heading_link.free_points,
                *self.actuation.free_points, *self.spring.free_points)

    def output_points(self) -> tuple:
        return tuple(dict.fromkeys((
            *self.LOCATING_OUTPUT_POINTS, *self.wheel_heading_link.output_points,
            *self.WHEEL_OUTPUT_POINTS, *self.actuation.output_points, *self.spring.output_points,
        )))

    def damper_points(self):
        return self.spring.damper_points

    def steering_axis_points(self) -> tuple:
        return (P.LOWER_WISHBONE_OUTBOARD, P.UPPER_WISHBONE_OUTBOARD)

    def derived_spec(self) -> DerivedPointsSpec:
        return build_wheel_derived_spec(self.config.wheel)

    def shim_rocker_coupled(self) -> bool:
        """An upright-mounted pushrod couples the rocker group into the shim assembly
        (double_wishbone.py:518-533)."""
        return (isinstance(self.actuation, ActuationPushrodRocker)
                and self.actuation.moving_pickup_body == self.UPRIGHT_BODY)

    def upright_attachment_points(self) -> tuple:
        """Points carried by the upright during camber-shim setup (double_wishbone.py:573-581)."""
        base = (P.AXLE_INBOARD, P.AXLE_OUTBOARD, self.wheel_heading_link.outboard_point)
        if self.actuation.moving_pickup_body == self.UPRIGHT_BODY:
            return (*base, self.actuation.moving_pickup_point)
        return base

    def authored_state(self) -> SuspensionState:
        """Hardpoints + derived points before any setup shim is applied."""
        positions = self.get_hardpoints_copy()
        DerivedPointsManager(self.derived_spec()).update_in_place(positions)
        return SuspensionState(positions=positions, free_points=set(self.free_points()))

    def initial_state(self) -> SuspensionState:
        if self._initial_state is None:
            shim = self.config.camber_shim
            if shim is not None and abs(shim.setup_thickness - shim.design_thickness) >= EPS_GEOMETRIC:
                # The split-body shim assembly pre-solve (reference config/shims.py:284-501) runs on
                # the device like every other piece of arithmetic: ask it for the setup pose.
                from ..shim_setup import device_setup_pose
                positions = device_setup_pose(self)
                self._initial_state = SuspensionState(positions=positions, free_points=set(self.free_points()))
            else:
                self._initial_state = self.authored_state()
        return self._initial_state

    def constraints(self) -> list:
        return self.constraints_at(self.initial_state().positions)

    def structure(self) -> tuple:
        state = self.authored_state()
        return state, self.constraints_at(state.positions)

    def constraints_at(self, pos: dict) -> list:
        """Constraint declarations with constants measured at the given pose."""
        rows: list[Constraint] = [
            distance_constraint(pos, a, b)
            for a, b in (
                (P.UPPER_WISHBONE_INBOARD_FRONT, P.UPPER_WISHBONE_OUTBOARD),
                (P.UPPER_WISHBONE_INBOARD_REAR, P.UPPER_WISHBONE_OUTBOARD),
                (P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_OUTBOARD),
                (P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD),
                (P.UPPER_WISHBONE_OUTBOARD, P.LOWER_WISHBONE_OUTBOARD),
                (P.AXLE_INBOARD, P.AXLE_OUTBOARD),
                (P.AXLE_INBOARD, P.UPPER_WISHBONE_OUTBOARD),
                (P.AXLE_INBOARD, P.LOWER_WISHBONE_OUTBOARD),
                (P.AXLE_OUTBOARD, P.UPPER_WISHBONE_OUTBOARD),
                (P.AXLE_OUTBOARD, P.LOWER_WISHBONE_OUTBOARD),
            )
        ]
        # Upright rigidity: angle between the kingpin vector and the axle vector.
        kingpin = pos[P.LOWER_WISHBONE_OUTBOARD].data - pos[P.UPPER_WISHBONE_OUTBOARD].data
        axle = pos[P.AXLE_OUTBOARD].data - pos[P.AXLE_INBOARD].data
        rows.append(AngleConstraint(
            v1_start=P.UPPER_WISHBONE_OUTBOARD, v1_end=P.LOWER_WISHBONE_OUTBOARD,
            v2_start=P.AXLE_INBOARD, v2_end=P.AXLE_OUTBOARD,
            target_angle=_angle(kingpin, axle),
        ))
        rows += self.wheel_heading_link.constraints(pos)
        rows += self.actuation.constraints(pos)
        rows += self.spring.constraints(pos, self.actuation)
        return rows


STRUT_AXIS_ALIGNMENT_TOLERANCE_MM = 1.0


@dataclass
class MacPhersonSuspension(_Corner):
    TYPE_KEY = SuspensionType.MACPHERSON
    UPRIGHT_BODY = (P.LOWER_WISHBONE_OUTBOARD, P.AXLE_INBOARD, P.AXLE_OUTBOARD)
    REQUIRED_POINTS = frozenset({
        P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD,
        P.STRUT_TOP, P.STRUT_BOTTOM, P.AXLE_INBOARD, P.AXLE_OUTBOARD,
    })
    SUPPORTED_SHIMS = frozenset()
    LOCATING_OUTPUT_POINTS = (
        P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD,
        P.STRUT_TOP, P.STRUT_BOTTOM,
    )
    WHEEL_OUTPUT_POINTS = DoubleWishboneSuspension.WHEEL_OUTPUT_POINTS
    FREE_POINTS = (P.LOWER_WISHBONE_OUTBOARD, P.AXLE_INBOARD, P.AXLE_OUTBOARD)

    def __post_init__(self) -> None:
        self.wheel_heading_link = self._heading_link(self.UPRIGHT_BODY, chiral=True)
        self.validate_hardpoints()

    def required_points(self) -> frozenset:
        return self.REQUIRED_POINTS | self.wheel_heading_link.required_points

    def validate_hardpoints(self) -> None:
        super().validate_hardpoints()
        self.wheel_heading_link.validate(self.hardpoints)
        ball, top = self.hardpoints[P.LOWER_WISHBONE_OUTBOARD].data, self.hardpoints[P.STRUT_TOP].data
        axis_length = float(np.linalg.norm(top - ball))
        if axis_length <= EPS_GEOMETRIC:
            raise ValueError("STRUT_TOP must not coincide with LOWER_WISHBONE_OUTBOARD; "
                             "the steering axis would be undefined.")
        off_axis = _line_distance(self.hardpoints[P.STRUT_BOTTOM].data, ball, _unit(top - ball))
        if off_axis > STRUT_AXIS_ALIGNMENT_TOLERANCE_MM:
            raise ValueError(
                f"STRUT_BOTTOM sits {off_axis:.3f} mm off the line from LOWER_WISHBONE_OUTBOARD to "
                "STRUT_TOP. This model treats the strut axis as coincident with the steering axis; "
                "an intentionally offset strut is not supported.")
        axial = self._strut_clamp_offset()
        if axial <= EPS_GEOMETRIC or axial >= axis_length - EPS_GEOMETRIC:
            raise ValueError("STRUT_BOTTOM must lie between LOWER_WISHBONE_OUTBOARD and STRUT_TOP "
                             "along the strut axis")

    def _strut_clamp_offset(self) -> float:
        ball = self.hardpoints[P.LOWER_WISHBONE_OUTBOARD].data
        axis = _unit(self.hardpoints[P.STRUT_TOP].data - ball)
        return float(np.dot(self.hardpoints[P.STRUT_BOTTOM].data - ball, axis))

    def free_points(self) -> tuple:
        return (*self.FREE_POINTS, *self.wheel_heading_link.free_points)

    def output_points(self) -> tuple:
        return tuple(dict.fromkeys((
            *self.LOCATING_OUTPUT_POINTS, *self.wheel_heading_link.output_points, *self.WHEEL_OUTPUT_POINTS,
        )))

    def steering_axis_points(self) -> tuple:
        return (P.LOWER_WISHBONE_OUTBOARD, P.STRUT_TOP)

    def damper_points(self):
        return (P.STRUT_TOP, P.STRUT_BOTTOM)

    def derived_spec(self) -> DerivedPointsSpec:
        wheel = build_wheel_derived_spec(self.config.wheel)
        clamp = PointAlongLine(P.LOWER_WISHBONE_OUTBOARD, P.STRUT_TOP, self._strut_clamp_offset())
        # ``design_projection`` tells the topology compiler that the distance is the
        # authored STRUT_BOTTOM projected on the strut axis, so a perturbed instance
        # recomputes it on the device (macpherson.py:199-204).
        clamp.design_projection = P.STRUT_BOTTOM
        functions = {P.STRUT_BOTTOM: clamp, **wheel.functions}
        dependencies = {P.STRUT_BOTTOM: set(clamp.inputs), **wheel.dependencies}
        return DerivedPointsSpec(functions, dependencies)

    def initial_state(self) -> SuspensionState:
        if self._initial_state is None:
            positions = self.get_hardpoints_copy()
            DerivedPointsManager(self.derived_spec()).update_in_place(positions)
            self._initial_state = SuspensionState(positions=positions, free_points=set(self.free_points()))
        return self._initial_state

    def authored_state(self) -> SuspensionState:
        return self.initial_state()

    def constraints(self) -> list:
        return self.constraints_at(self.initial_state().positions)

    def constraints_at(self, pos: dict) -> list:
        rows: list[Constraint] = [
            distance_constraint(pos, a, b)
            for a, b in (
                (P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_OUTBOARD),
                (P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD),
                (P.AXLE_INBOARD, P.AXLE_OUTBOARD),
                (P.AXLE_INBOARD, P.LOWER_WISHBONE_OUTBOARD),
                (P.AXLE_OUTBOARD, P.LOWER_WISHBONE_OUTBOARD),
            )
        ]
        # The derived strut clamp rides the ball-joint-to-top line; holding the upright
        # to it leaves the clamp-to-top distance free to telescope (macpherson.py:288-297).
        rows += chiral_rigid_point_constraints(pos, P.STRUT_BOTTOM, self.UPRIGHT_BODY)
        rows += self.wheel_heading_link.constraints(pos)
        return rows
