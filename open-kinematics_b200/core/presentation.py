"""Named positions of a solved state, including the presentation points the reference derives from
its element model (reference core/presentation.py:26-47, :97-118, :252-348): every rocker pickup
projected onto the rocker's rotation axis, and the midpoint of a T-bar's two droplink attachments.

The reference collects those points by walking ``suspension.assembly()`` (element paths).  Here they
are read off the mechanism objects of the host model, which carry the same information: a
pushrod-rocker actuation knows its axis and its pickups (pushrod + externally mounted droplink /
heave-link points), a T-bar knows its attachments.  Names, order and arithmetic follow the reference,
so ``analyze_sweep`` frames carry the same keys and values.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .enums import PointID
from .primitives.point_ref import PointRef, Side, point_key_name


@dataclass(frozen=True)
class AxisProjection:
    """Presentation point projected onto a physical rotation axis."""

    point: object
    rotation_axis: tuple


@dataclass(frozen=True)
class PointMidpoint:
    """Presentation midpoint of two physical element points."""

    point_a: object
    point_b: object


def axis_projection_name(projection: AxisProjection) -> str:
    axis_names = sorted(point_key_name(p) for p in projection.rotation_axis)
    return f"{point_key_name(projection.point)}_axis_projection_{axis_names[0]}_{axis_names[1]}"


def point_midpoint_name(midpoint: PointMidpoint) -> str:
    names = sorted((point_key_name(midpoint.point_a), point_key_name(midpoint.point_b)))
    return f"{names[0]}_{names[1]}_midpoint"


def _corner_projections(corner, qualify) -> list:
    from .suspensions.corner import ActuationPushrodRocker
    actuation = getattr(corner, "actuation", None)
    if not isinstance(actuation, ActuationPushrodRocker):
        return []
    axis = tuple(qualify(p) for p in actuation.torsion_axis)
    pickups = (PointID.PUSHROD_INBOARD, *actuation.external_point_ids)
    return [AxisProjection(qualify(p), axis) for p in pickups]


def presentation_points(suspension) -> tuple:
    """``(projections, midpoints)`` in the reference's element order: left corner, right corner,
    then the shared axle mechanisms."""
    projections: list = []
    midpoints: list = []
    if getattr(suspension, "is_axle", False):
        from .suspensions.axle import ArbTBar, T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY
        for side in (Side.LEFT, Side.RIGHT):
            projections += _corner_projections(suspension.corners[side], lambda p, side=side: PointRef(side, p))
        if isinstance(suspension.anti_roll, ArbTBar):
            midpoints.append(PointMidpoint(T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY))
    else:
        projections += _corner_projections(suspension, lambda p: p)
    return list(dict.fromkeys(projections)), list(dict.fromkeys(midpoints))


def named_point_keys(suspension, positions) -> list:
    """Every physical and projected position name in stable order (presentation.py:252-264)."""
    projections, midpoints = presentation_points(suspension)
    return ([point_key_name(k) for k in positions] + [axis_projection_name(p) for p in projections]
            + [point_midpoint_name(m) for m in midpoints])


def resolve_positions(positions: dict, suspension) -> dict:
    """One solver state -> all named physical and projected positions (presentation.py:295-348).

    Raises ``ValueError`` if a presentation point refers to a missing point or a projection axis
    is degenerate."""
    projections, midpoints = presentation_points(suspension)
    needed = [k for p in projections for k in (p.point, *p.rotation_axis)] + \
             [k for m in midpoints for k in (m.point_a, m.point_b)]
    missing = [k for k in dict.fromkeys(needed) if k not in positions]
    if missing:
        raise ValueError(f"Cannot resolve missing assembly points: {missing!r}")
    named = {point_key_name(k): tuple(float(v) for v in np.asarray(getattr(p, "data", p), dtype=np.float64))
             for k, p in positions.items()}

    def xyz(key) -> np.ndarray:
        return np.asarray(named[point_key_name(key)], dtype=np.float64)

    for projection in projections:
        point, start, end = xyz(projection.point), xyz(projection.rotation_axis[0]), xyz(projection.rotation_axis[1])
        direction = end - start
        length_sq = float(direction @ direction)
        if length_sq <= 0.0:
            raise ValueError(f"Cannot project onto a zero-length rotation axis: {projection.rotation_axis!r}")
        projected = start + float((point - start) @ direction) / length_sq * direction
        named[axis_projection_name(projection)] = tuple(float(v) for v in projected)
    for midpoint in midpoints:
        a, b = xyz(midpoint.point_a), xyz(midpoint.point_b)
        named[point_midpoint_name(midpoint)] = tuple(float(v) for v in a + (b - a) / 2.0)
    return named
