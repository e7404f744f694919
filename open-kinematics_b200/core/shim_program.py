"""Camber-shim pre-solve records for the device (reference
``corner/double_wishbone.py:501-581``, ``suspensions/config/shims.py:284-501``).

A record names the points the split-body shim assembly reads (upper/lower ball joints, upper
wishbone pickups, heading link, optional upright-mounted pushrod + rocker axis), the points the
solved upright rotation and rocker rotation carry, and the default per-instance parameters
(face datums A/B, face normal, design and setup thickness)."""

from __future__ import annotations

from .enums import PointID
from .primitives.point_ref import PointRef

P = PointID


def corner_shim_record(corner, key=lambda pid: pid, label: str = "camber_shim"):
    """Record for one double-wishbone corner, or None when no shim is configured."""
    shim = getattr(corner.config, "camber_shim", None)
    if shim is None or not hasattr(corner, "upright_attachment_points"):
        return None
    present = set(corner.hardpoints)
    coupled = corner.shim_rocker_coupled()
    rocker_points = []
    if coupled:
        rocker_points = [p for p in dict.fromkeys((*corner.actuation.rocker_mounted_point_ids,
                                                   *corner.spring.rocker_mounted_points)) if p in present]
    return {
        "label": label,
        "ubj": key(P.UPPER_WISHBONE_OUTBOARD), "lbj": key(P.LOWER_WISHBONE_OUTBOARD),
        "uw_front": key(P.UPPER_WISHBONE_INBOARD_FRONT), "uw_rear": key(P.UPPER_WISHBONE_INBOARD_REAR),
        "heading_in": key(corner.wheel_heading_link.inboard_point),
        "heading_out": key(corner.wheel_heading_link.outboard_point),
        "rocker": [key(P.ROCKER_AXIS_A), key(P.ROCKER_AXIS_B), key(P.PUSHROD_INBOARD), key(P.PUSHROD_OUTBOARD)]
        if coupled else None,
        "upright_points": [key(p) for p in corner.upright_attachment_points() if p in present],
        "rocker_points": [key(p) for p in rocker_points],
        "params": [*shim.shim_face_point_a.data, *shim.shim_face_point_b.data, *shim.shim_face_normal.data,
                   shim.design_thickness, shim.setup_thickness],
    }


def shim_records(suspension) -> list:
    if getattr(suspension, "is_axle", False):
        out = []
        for side, corner in suspension.corners.items():
            rec = corner_shim_record(corner, lambda pid, side=side: PointRef(side, pid),
                                     f"{side.name.lower()}.camber_shim")
            if rec is not None:
                out.append(rec)
        return out
    rec = corner_shim_record(suspension)
    return [rec] if rec is not None else []
