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


class TargetPositionMode(StrEnum):
    RELATIVE = "relative"
    ABSOLUTE = "absolute"


class Units(StrEnum):
    MILLIMETERS = "millimeters"
    DEGREES = "degrees"


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


class ShimType(StrEnum):
    OUTBOARD_CAMBER = "outboard_camber"


class SuspensionType(StrEnum):
    DOUBLE_WISHBONE = "double_wishbone"
    MACPHERSON = "macpherson"


class Scope(StrEnum):
    CORNER = "corner"
    AXLE = "axle"


class AxlePosition(StrEnum):
    FRONT = "front"
    REAR = "rear"


class ActuationType(StrEnum):
    DIRECT = "direct"
    PUSHROD_ROCKER = "pushrod_rocker"


class MountBody(StrEnum):
    LOWER_WISHBONE = "lower_wishbone"
    UPRIGHT = "upright"


class CornerSpringType(StrEnum):
    NONE = "none"
    COILOVER = "coilover"
    TORSION_BAR = "torsion_bar"


class ArbType(StrEnum):
    NONE = "none"
    U_BAR = "u_bar"
    T_BAR = "t_bar"


class HeaveLinkType(StrEnum):
    NONE = "none"
    ROCKER_TO_ROCKER = "rocker_to_rocker"


class SteeringType(StrEnum):
    NONE = "none"
    RACK = "rack"
