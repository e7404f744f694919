"""Constraint declarations accepted by ``solve_suspension_sweep``.

Same class names, constructor arguments and validation errors as the reference
(src/kinematics/core/constraints.py:89-733).  Here they are *declarations*: the
topology compiler (``core/topology.py``) turns them into constraint-table rows
and the residual/Jacobian arithmetic runs in the generated CUDA device
functions (``csrc/okin_gen_constraints.cuh``).
"""

from __future__ import annotations

import copy
from typing import Callable, ClassVar

import numpy as np

from .enums import Axis
from .primitives.geometry import Direction3, Point3
from .primitives.point_ref import PointKey


class Constraint:
    """Base declaration.  ``_POINT_ATTRS`` names the point-key attributes."""

    _POINT_ATTRS: ClassVar[tuple[str, ...]] = ()
    #: constraint-family code understood by the device tables (csrc/okin_types.h)
    FAMILY: ClassVar[str] = ""

    @property
    def involved_points(self) -> set:
        return {getattr(self, a) for a in self._POINT_ATTRS}

    @property
    def point_keys(self) -> tuple:
        return tuple(getattr(self, a) for a in self._POINT_ATTRS)

    def remap(self, mapping: Callable[[PointKey], PointKey]) -> "Constraint":
        new = copy.copy(self)
        for attr in self._POINT_ATTRS:
            setattr(new, attr, mapping(getattr(self, attr)))
        return new

    def __repr__(self) -> str:
        names = ", ".join(getattr(k, "name", str(k)) for k in self.point_keys)
        return f"{type(self).__name__}({names})"


class DistanceConstraint(Constraint):
    _POINT_ATTRS = ("p1", "p2")
    FAMILY = "distance"

    def __init__(self, p1: PointKey, p2: PointKey, target_distance: float):
        if target_distance < 0:
            raise ValueError(f"Target distance must be non-negative, got {target_distance}")
        self.p1, self.p2 = p1, p2
        self.target_distance = float(target_distance)


class SphericalJointConstraint(Constraint):
    _POINT_ATTRS = ("p1", "p2")
    FAMILY = "spherical"

    def __init__(self, p1: PointKey, p2: PointKey):
        self.p1, self.p2 = p1, p2


def _check_angle(target_angle: float) -> float:
    if not (0 <= target_angle <= np.pi):
        raise ValueError(f"Target angle must be in [0, pi], got {target_angle}")
    return float(target_angle)


class AngleConstraint(Constraint):
    _POINT_ATTRS = ("v1_start", "v1_end", "v2_start", "v2_end")
    FAMILY = "angle"

    def __init__(self, v1_start, v1_end, v2_start, v2_end, target_angle: float):
        self.target_angle = _check_angle(target_angle)
        self.v1_start, self.v1_end = v1_start, v1_end
        self.v2_start, self.v2_end = v2_start, v2_end


class ThreePointAngleConstraint(Constraint):
    _POINT_ATTRS = ("p1", "p2", "p3")
    FAMILY = "three_point_angle"

    def __init__(self, p1, p2, p3, target_angle: float):
        self.target_angle = _check_angle(target_angle)
        self.p1, self.p2, self.p3 = p1, p2, p3


class _TwoVectorConstraint(Constraint):
    _POINT_ATTRS = ("v1_start", "v1_end", "v2_start", "v2_end")

    def __init__(self, v1_start, v1_end, v2_start, v2_end):
        self.v1_start, self.v1_end = v1_start, v1_end
        self.v2_start, self.v2_end = v2_start, v2_end


class VectorsParallelConstraint(_TwoVectorConstraint):
    FAMILY = "vectors_parallel"


class VectorsPerpendicularConstraint(_TwoVectorConstraint):
    FAMILY = "vectors_perpendicular"


class EqualDistanceConstraint(Constraint):
    _POINT_ATTRS = ("p1", "p2", "p3", "p4")
    FAMILY = "equal_distance"

    def __init__(self, p1, p2, p3, p4):
        self.p1, self.p2, self.p3, self.p4 = p1, p2, p3, p4


class FixedAxisConstraint(Constraint):
    _POINT_ATTRS = ("point_id",)
    FAMILY = "fixed_axis"

    def __init__(self, point_id, axis: Axis, value: float):
        self.point_id = point_id
        self.axis = Axis(axis)
        self.value = float(value)


class PointOnLineConstraint(Constraint):
    _POINT_ATTRS = ("point_id",)
    FAMILY = "point_on_line"

    def __init__(self, point_id, line_point: Point3, line_direction: Direction3):
        if not isinstance(line_point, Point3):
            raise TypeError("line_point must be a Point3")
        if not isinstance(line_direction, Direction3):
            raise TypeError("line_direction must be a Direction3")
        self.point_id = point_id
        self.line_point = line_point.copy()
        self.line_direction = line_direction


class PointOnPlaneConstraint(Constraint):
    _POINT_ATTRS = ("point_id",)
    FAMILY = "point_on_plane"

    def __init__(self, point_id, plane_point: Point3, plane_normal: Direction3):
        if not isinstance(plane_point, Point3):
            raise TypeError("plane_point must be a Point3")
        if not isinstance(plane_normal, Direction3):
            raise TypeError("plane_normal must be a Direction3")
        self.point_id = point_id
        self.plane_point = plane_point.copy()
        self.plane_normal = plane_normal


class MidpointOnPlaneConstraint(Constraint):
    _POINT_ATTRS = ("point_a", "point_b")
    FAMILY = "midpoint_on_plane"

    def __init__(self, point_a, point_b, plane_point: Point3, plane_normal: Direction3):
        if not isinstance(plane_point, Point3):
            raise TypeError("plane_point must be a Point3")
        if not isinstance(plane_normal, Direction3):
            raise TypeError("plane_normal must be a Direction3")
        self.point_a, self.point_b = point_a, point_b
        self.plane_point = plane_point.copy()
        self.plane_normal = plane_normal


class CoplanarPointsConstraint(Constraint):
    _POINT_ATTRS = ("p1", "p2", "p3", "p4")
    FAMILY = "coplanar"

    def __init__(self, p1, p2, p3, p4):
        self.p1, self.p2, self.p3, self.p4 = p1, p2, p3, p4


class ScalarTripleProductConstraint(CoplanarPointsConstraint):
    FAMILY = "scalar_triple"

    def __init__(self, p1, p2, p3, p4, target_volume: float, scale: float = 1.0):
        if scale <= 0.0:
            raise ValueError(f"scale must be strictly positive, got {scale}")
        super().__init__(p1, p2, p3, p4)
        self.target_volume = float(target_volume)
        self.scale = float(scale)
