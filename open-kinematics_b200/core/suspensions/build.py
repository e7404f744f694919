"""Geometry mapping -> suspension model (reference core/suspensions/build.py:57-389,
core/schema/geometry.py:95-215, core/schema/decoding.py:19-82, registry.py:43-87).

The YAML/mapping schema is the reference's; parsing is hand-rolled (no pydantic
dependency on the hot-path host shim) and rejects unknown keys.
"""

from __future__ import annotations

import numpy as np

from ..enums import (
    ActuationType, ArbType, AxlePosition, CornerSpringType, HeaveLinkType, MountBody, PointID,
    Scope, SuspensionType, Units,
)
from ..primitives.geometry import Point3
from ..primitives.point_ref import Side
from ..schema.config import (
    CamberShimConfig, SteeringConfig, SuspensionConfig, VehicleConfig, WheelConfig,
    _forbid_extra, decode_point,
)
from .axle import ArbNone, ArbTBar, ArbUBar, AxleSuspension, HeaveLink
from .corner import (
    ActuationDirect, ActuationPushrodRocker, CornerSpring, DoubleWishboneSuspension,
    MacPhersonSuspension, _validate_side_sign,
)

P = PointID


def _decode_enum(enum_cls, value, what: str):
    if isinstance(value, enum_cls):
        return value
    try:
        return enum_cls(value)
    except ValueError:
        options = ", ".join(repr(m.value) for m in enum_cls)
        raise ValueError(f"Invalid {what} {value!r}; expected one of {options}") from None


def decode_point_id(name) -> PointID:
    if isinstance(name, PointID):
        return name
    key = str(name)
    if key != key.lower():
        raise ValueError(f"Unknown point '{name}' (point names are lowercase snake_case)")
    try:
        return PointID[key.upper()]
    except KeyError:
        raise ValueError(f"Unknown point '{name}'") from None


def decode_side(value) -> Side:
    if isinstance(value, Side):
        return value
    try:
        return {"left": Side.LEFT, "right": Side.RIGHT, "center": Side.CENTER}[value]
    except KeyError:
        raise ValueError(f"Invalid side {value!r}; expected 'left', 'right' or 'center'") from None


def _decode_hardpoints(m: dict) -> dict:
    return {decode_point_id(k): decode_point(v) for k, v in m.items()}


def _mirror(points: dict) -> dict:
    """Reflect through the vehicle XZ plane (build.py:344-354)."""
    flip = np.array([1.0, -1.0, 1.0])
    return {k: Point3(p.data * flip) for k, p in points.items()}


def _check_dw_combination(actuation: dict, spring: dict) -> None:
    if actuation["type"] is ActuationType.DIRECT and spring["type"] is CornerSpringType.TORSION_BAR:
        raise ValueError("Direct torsion-bar actuation is not implemented yet")


def _decode_actuation(m: dict) -> dict:
    _forbid_extra(m, {"type", "mount"}, "actuation")
    return {"type": _decode_enum(ActuationType, m["type"], "actuation type"),
            "mount": _decode_enum(MountBody, m["mount"], "mount body")}


def _decode_spring(m: dict) -> dict:
    _forbid_extra(m, {"type"}, "spring")
    return {"type": _decode_enum(CornerSpringType, m["type"], "spring type")}


def _build_actuation(spec: dict, external_points: tuple = ()):
    bodies = DoubleWishboneSuspension.MOUNT_BODIES
    if spec["type"] is ActuationType.DIRECT:
        if external_points:
            raise ValueError("Direct actuation does not accept rocker pickups")
        return ActuationDirect(spring_pickup_body=bodies[spec["mount"]])
    return ActuationPushrodRocker(pushrod_outboard_body=bodies[spec["mount"]], external_point_ids=external_points)


_COMMON_KEYS = {"name", "version", "units", "type", "scope"}


def _common(m: dict) -> dict:
    units = _decode_enum(Units, m.get("units", "millimeters"), "units")
    return {"name": str(m.get("name", "unnamed")), "version": str(m.get("version", "0.0.0")), "units": units}


def _build_dw_corner(common, side, config, actuation, spring, hardpoints, external_points=()):
    _check_dw_combination(actuation, spring)
    _validate_side_sign(hardpoints, side)
    return DoubleWishboneSuspension(
        name=common["name"], version=common["version"], side=side,
        hardpoints={k: p.copy() for k, p in hardpoints.items()}, config=config,
        actuation=_build_actuation(actuation, external_points),
        spring=CornerSpring(spring["type"].value),
    )


def _build_mac_corner(common, side, config, hardpoints):
    _validate_side_sign(hardpoints, side)
    if config.camber_shim is not None:
        raise ValueError("Suspension type 'macpherson' does not support outboard camber shims")
    return MacPhersonSuspension(
        name=common["name"], version=common["version"], side=side,
        hardpoints={k: p.copy() for k, p in hardpoints.items()}, config=config,
    )


def _build_corner(m: dict, stype: SuspensionType):
    common = _common(m)
    side = decode_side(m.get("side", "left"))
    if side is Side.CENTER:
        raise ValueError("Corner geometry side must be 'left' or 'right'.")
    config = SuspensionConfig.from_mapping(m["config"])
    hardpoints = _decode_hardpoints(m["hardpoints"])
    if stype is SuspensionType.DOUBLE_WISHBONE:
        _forbid_extra(m, _COMMON_KEYS | {"side", "config", "actuation", "spring", "hardpoints"}, "geometry")
        return _build_dw_corner(common, side, config, _decode_actuation(m["actuation"]),
                                _decode_spring(m["spring"]), hardpoints)
    _forbid_extra(m, _COMMON_KEYS | {"side", "config", "hardpoints"}, "geometry")
    return _build_mac_corner(common, side, config, hardpoints)


def _build_axle(m: dict, stype: SuspensionType):
    _forbid_extra(m, _COMMON_KEYS | {"vehicle_config", "axle_config", "hardpoints"}, "geometry")
    common = _common(m)
    vehicle = VehicleConfig.from_mapping(m["vehicle_config"])
    ac = dict(m["axle_config"])
    is_dw = stype is SuspensionType.DOUBLE_WISHBONE
    allowed = {"axle_position", "steering", "wheel", "anti_roll", "heave_link"}
    if is_dw:
        allowed |= {"actuation", "spring", "left_setup", "right_setup"}
    _forbid_extra(ac, allowed, "axle_config")
    axle_position = _decode_enum(AxlePosition, ac["axle_position"], "axle position")
    steering = SteeringConfig.from_mapping(ac["steering"])
    wheel = WheelConfig.from_mapping(ac["wheel"])
    arb = _decode_enum(ArbType, ac["anti_roll"]["type"], "anti-roll type")
    heave = _decode_enum(HeaveLinkType, ac["heave_link"]["type"], "heave-link type")

    hp = m["hardpoints"]
    _forbid_extra(hp, {"left", "right", "center"}, "hardpoints")
    left = _decode_hardpoints(hp["left"])
    explicit_right = hp.get("right") is not None
    right = _decode_hardpoints(hp["right"]) if explicit_right else _mirror(left)
    center = _decode_hardpoints(hp.get("center") or {})
    side_points = {Side.LEFT: left, Side.RIGHT: right}

    if not is_dw:
        if arb in (ArbType.U_BAR, ArbType.T_BAR):
            raise ValueError("The implemented anti-roll mechanism requires pushrod-rocker actuation, "
                             "which a MacPherson corner does not provide")
        if heave is HeaveLinkType.ROCKER_TO_ROCKER:
            raise ValueError("A rocker-to-rocker heave link requires pushrod-rocker actuation, "
                             "which a MacPherson corner does not provide")
        corners = {
            side: _build_mac_corner(
                {**common, "name": f"{common['name']}_{side.name.lower()}"}, side,
                SuspensionConfig.from_parts(vehicle, steering, wheel, axle_position, None),
                side_points[side])
            for side in (Side.LEFT, Side.RIGHT)
        }
        droplinks: dict = {}
    else:
        actuation = _decode_actuation(ac["actuation"])
        spring = _decode_spring(ac["spring"])
        _check_dw_combination(actuation, spring)
        has_rocker = actuation["type"] is ActuationType.PUSHROD_ROCKER
        if arb in (ArbType.U_BAR, ArbType.T_BAR) and not has_rocker:
            raise ValueError("The implemented anti-roll mechanism requires pushrod-rocker actuation")
        if heave is HeaveLinkType.ROCKER_TO_ROCKER and not has_rocker:
            raise ValueError("A rocker-to-rocker heave link requires pushrod-rocker actuation")

        def setup(key):
            block = ac.get(key)
            if block is None:
                return None
            _forbid_extra(block, {"camber_shim"}, key)
            shim = block.get("camber_shim")
            return {"camber_shim": None if shim is None else CamberShimConfig.from_mapping(shim)}

        left_setup = setup("left_setup") or {"camber_shim": None}
        right_setup = setup("right_setup")
        if right_setup is not None and not explicit_right:
            raise ValueError("axle_config.right_setup requires explicit hardpoints.right")
        if explicit_right and left_setup["camber_shim"] is not None and right_setup is None:
            raise ValueError("Explicit hardpoints.right requires axle_config.right_setup when "
                             "axle_config.left_setup contains side-local setup")
        if right_setup is None:
            shim = left_setup["camber_shim"]
            right_setup = {"camber_shim": None if shim is None else shim.mirrored()}
        setups = {Side.LEFT: left_setup, Side.RIGHT: right_setup}

        external: list = []
        droplinks = {}
        if arb in (ArbType.U_BAR, ArbType.T_BAR):
            external.append(P.DROPLINK_ROCKER)
            arm_id = P.DROPLINK_U_BAR if arb is ArbType.U_BAR else P.DROPLINK_T_BAR
            for side, points in side_points.items():
                if arm_id not in points:
                    raise ValueError(f"{side.name} {arb.value.replace('_', '-')} requires {arm_id.name}")
                droplinks[side] = points.pop(arm_id)
        if heave is HeaveLinkType.ROCKER_TO_ROCKER:
            external.append(P.HEAVE_LINK_ROCKER)
        corners = {
            side: _build_dw_corner(
                {**common, "name": f"{common['name']}_{side.name.lower()}"}, side,
                SuspensionConfig.from_parts(vehicle, steering, wheel, axle_position, setups[side]["camber_shim"]),
                actuation, spring, side_points[side], tuple(external))
            for side in (Side.LEFT, Side.RIGHT)
        }

    if arb is ArbType.NONE:
        if center:
            raise ValueError("Axle without anti-roll hardware does not accept center points")
        anti_roll = ArbNone()
    elif arb is ArbType.U_BAR:
        anti_roll = ArbUBar(center_points=center, droplink_points=droplinks)
    else:
        anti_roll = ArbTBar(center_points=center, droplink_points=droplinks)

    return AxleSuspension(
        type_key=stype, name=common["name"], version=common["version"],
        config=corners[Side.LEFT].config, corners=corners, anti_roll=anti_roll,
        heave_link=HeaveLink(heave.value),
    )


def build_from_mapping(data: dict):
    if not isinstance(data, dict):
        raise ValueError("Geometry must be a mapping")
    if "type" not in data:
        raise ValueError("Geometry mapping requires a 'type'")
    stype = _decode_enum(SuspensionType, data["type"], "suspension type")
    scope = _decode_enum(Scope, data.get("scope", "corner"), "scope")
    if scope is Scope.AXLE:
        return _build_axle(data, stype)
    return _build_corner(data, stype)
