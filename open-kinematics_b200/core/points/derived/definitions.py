"""Declarative derived-point functions (formulas: reference
core/points/derived/definitions.py:24-180, corner/macpherson.py:307-313).

Each class names a device op (``OP``) of ``csrc/okin_core.cuh``; ``inputs`` are
the point keys it reads and ``param`` its scalar parameter.
"""

from __future__ import annotations

import numpy as np

from ...enums import PointID
from ...primitives.constants import EPS_GEOMETRIC
from ...primitives.geometry import Point3
from .manager import DerivedPointsSpec


def _unit(v: np.ndarray) -> np.ndarray:
    length = float(np.sqrt(np.dot(v, v)))
    if length < EPS_GEOMETRIC:
        raise ValueError("Cannot normalize a zero-length vector")
    return v / length


class DerivedFn:
    OP = ""
    inputs: tuple = ()
    param: float = 0.0
    #: when set, ``param`` is the authored position of this point projected on the
    #: op's line at the design pose, recomputed per instance on the device
    design_projection = None

    def remap(self, mapping) -> "DerivedFn":
        import copy

        new = copy.copy(self)
        new.inputs = tuple(mapping(k) for k in self.inputs)
        if self.design_projection is not None:
            new.design_projection = mapping(self.design_projection)
        return new

    def __call__(self, positions: dict) -> Point3:
        raise NotImplementedError


class Midpoint(DerivedFn):
    """a + (b - a)/2  (definitions.py:76-89)."""

    OP = "midpoint"

    def __init__(self, a, b):
        self.inputs = (a, b)

    def __call__(self, positions):
        a, b = (positions[k].data for k in self.inputs)
        return Point3.from_trusted(a + (b - a) / 2.0)


class PointAlongLine(DerivedFn):
    """start + unit(end - start) * distance  (definitions.py:24-33).

    The wheel centre, rim faces and MacPherson strut clamp are all this op
    (definitions.py:92-155 use ``p - unit(p - q) * d == p + unit(q - p) * d``).
    """

    OP = "along_line"

    def __init__(self, start, end, distance: float):
        self.inputs = (start, end)
        self.param = float(distance)

    def __call__(self, positions):
        s, e = (positions[k].data for k in self.inputs)
        return Point3.from_trusted(s + _unit(e - s) * self.param)


class ContactPatch(DerivedFn):
    """wheel_center + unit(down - (down.a) a) * tire_radius, a = unit(ao - ai),
    down = -Z  (definitions.py:36-73, :158-180)."""

    OP = "contact_patch"

    def __init__(self, wheel_center, axle_inboard, axle_outboard, tire_radius: float):
        self.inputs = (wheel_center, axle_inboard, axle_outboard)
        self.param = float(tire_radius)

    def __call__(self, positions):
        wc, ai, ao = (positions[k].data for k in self.inputs)
        a = _unit(ao - ai)
        down = np.array([0.0, 0.0, -1.0])
        return Point3.from_trusted(wc + _unit(down - np.dot(down, a) * a) * self.param)


def build_wheel_derived_spec(wheel) -> DerivedPointsSpec:
    """Standard wheel points of a corner whose spin axis is AXLE_INBOARD -> AXLE_OUTBOARD
    (definitions.py:183-216)."""
    P = PointID
    half_width = wheel.tire.section_width / 2
    functions = {
        P.AXLE_MIDPOINT: Midpoint(P.AXLE_INBOARD, P.AXLE_OUTBOARD),
        # wc = ao - unit(ao - ai) * offset
        P.WHEEL_CENTER: PointAlongLine(P.AXLE_OUTBOARD, P.AXLE_INBOARD, wheel.offset),
        # wi = wc - unit(wc - ai) * w/2 ; wo = wc + unit(wc - ai) * w/2
        P.WHEEL_INBOARD: PointAlongLine(P.WHEEL_CENTER, P.AXLE_INBOARD, half_width),
        P.WHEEL_OUTBOARD: PointAlongLine(P.WHEEL_CENTER, P.AXLE_INBOARD, -half_width),
        P.CONTACT_PATCH_CENTER: ContactPatch(
            P.WHEEL_CENTER, P.AXLE_INBOARD, P.AXLE_OUTBOARD, wheel.tire.nominal_radius
        ),
    }
    dependencies = {key: set(fn.inputs) for key, fn in functions.items()}
    return DerivedPointsSpec(functions=functions, dependencies=dependencies)
