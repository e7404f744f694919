"""Enumerations shared with the reference's YAML schema.

The member names and integer values are interface constants: ``PointID`` values
fix the sorted free-point (solver column) order, see reference
``src/kinematics/core/enums.py:33-81`` and ``core/state.py:50``.
"""

from enum import IntEnum, StrEnum


class Axis(IntEnum):
    X = 0
    Y = 1
    Z = 2


def _schema_enum(name: str, *values: str) -> type:
    """String enumeration of one schema keyword set: member ``FOO_BAR`` has the YAML value ``foo_bar``."""
    return StrEnum(name, {value.upper(): value for value in values})


# Keyword sets of the geometry / sweep schema (reference enums.py:13-30, :84-170).
TargetPositionMode = _schema_enum("TargetPositionMode", "relative", "absolute")
Units = _schema_enum("Units", "millimeters", "degrees")
ShimType = _schema_enum("ShimType", "outboard_camber")
SuspensionType = _schema_enum("SuspensionType", "double_wishbone", "macpherson")
Scope = _schema_enum("Scope", "corner", "axle")
AxlePosition = _schema_enum("AxlePosition", "front", "rear")
ActuationType = _schema_enum("ActuationType", "direct", "pushrod_rocker")
MountBody = _schema_enum("MountBody", "lower_wishbone", "upright")
CornerSpringType = _schema_enum("CornerSpringType", "none", "coilover", "torsion_bar")
ArbType = _schema_enum("ArbType", "none", "u_bar", "t_bar")
HeaveLinkType = _schema_enum("HeaveLinkType", "none", "rocker_to_rocker")
SteeringType = _schema_enum("SteeringType", "none", "rack")

_POINT_NAMES = """
NOT_ASSIGNED
LOWER_WISHBONE_INBOARD_FRONT LOWER_WISHBONE_INBOARD_REAR LOWER_WISHBONE_OUTBOARD
UPPER_WISHBONE_INBOARD_FRONT UPPER_WISHBONE_INBOARD_REAR UPPER_WISHBONE_OUTBOARD
PUSHROD_INBOARD PUSHROD_OUTBOARD
TRACKROD_INBOARD TRACKROD_OUTBOARD TOE_LINK_INBOARD TOE_LINK_OUTBOARD
AXLE_INBOARD AXLE_OUTBOARD AXLE_MIDPOINT
STRUT_TOP STRUT_BOTTOM
WHEEL_CENTER WHEEL_INBOARD WHEEL_OUTBOARD
CONTACT_PATCH_CENTER
CAMBER_SHIM_FACE_POINT_A CAMBER_SHIM_FACE_POINT_B CAMBER_SHIM_FACE_NORMAL
ROCKER_AXIS_A ROCKER_AXIS_B DROPLINK_ROCKER DROPLINK_U_BAR
ARB_U_BAR_AXIS_A ARB_U_BAR_AXIS_B HEAVE_LINK_ROCKER ARB_T_BAR_PIVOT DROPLINK_T_BAR
""".split()

# Values 0..33 in declaration order (reference enums.py:36-81).
PointID = IntEnum("PointID", {name: i for i, name in enumerate(_POINT_NAMES)})
PointID.__doc__ = "Identifiers for authored and derived suspension points."


