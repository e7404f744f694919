"""Configuration blocks of a geometry file (reference core/schema/config.py:12-159).

Plain frozen dataclasses with explicit ``from_mapping`` parsers; unknown keys are
rejected like the reference's ``extra="forbid"`` models.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..enums import AxlePosition, SteeringType
from ..primitives.constants import EPS_GEOMETRIC, MM_PER_INCH
from ..primitives.geometry import Direction3, Point3


def _forbid_extra(mapping: dict, allowed: set, what: str) -> None:
    extra = set(mapping) - allowed
    if extra:
        raise ValueError(f"{what}: unexpected keys {sorted(extra)}")


def decode_point(value) -> Point3:
    if isinstance(value, Point3):
        return value.copy()
    if isinstance(value, dict):
        if set(value) != {"x", "y", "z"}:
            raise ValueError(f"A point requires exactly keys x, y, z; got {sorted(value)}")
        return Point3([float(value["x"]), float(value["y"]), float(value["z"])])
    return Point3(value)


def decode_direction(value) -> Direction3:
    if isinstance(value, Direction3):
        return value
    if isinstance(value, dict):
        return Direction3([float(value["x"]), float(value["y"]), float(value["z"])])
    return Direction3(value)


@dataclass(frozen=True)
class TireConfig:
    aspect_ratio: float
    section_width: float
    rim_diameter: float

    def __post_init__(self):
        if not 0 <= self.aspect_ratio <= 1:
            raise ValueError(f"aspect_ratio must be in [0, 1], got {self.aspect_ratio}")

    @property
    def sidewall_height(self) -> float:
        return self.aspect_ratio * self.section_width

    @property
    def rim_diameter_mm(self) -> float:
        return self.rim_diameter * MM_PER_INCH

    @property
    def nominal_radius(self) -> float:
        return (self.rim_diameter_mm + 2 * self.sidewall_height) / 2

    @classmethod
    def from_mapping(cls, m: dict) -> "TireConfig":
        _forbid_extra(m, {"aspect_ratio", "section_width", "rim_diameter"}, "tire")
        return cls(float(m["aspect_ratio"]), float(m["section_width"]), float(m["rim_diameter"]))


@dataclass(frozen=True)
class WheelConfig:
    offset: float
    tire: TireConfig

    @classmethod
    def from_mapping(cls, m: dict) -> "WheelConfig":
        _forbid_extra(m, {"offset", "tire"}, "wheel")
        return cls(float(m["offset"]), TireConfig.from_mapping(m["tire"]))


@dataclass(frozen=True)
class CamberShimConfig:
    shim_face_point_a: Point3
    shim_face_point_b: Point3
    shim_face_normal: Direction3
    design_thickness: float
    setup_thickness: float

    def __post_init__(self):
        sep = float(np.linalg.norm(self.shim_face_point_b.data - self.shim_face_point_a.data))
        if sep < EPS_GEOMETRIC:
            raise ValueError("shim_face_point_a and shim_face_point_b must be distinct")

    @classmethod
    def from_mapping(cls, m: dict) -> "CamberShimConfig":
        _forbid_extra(
            m,
            {"shim_face_point_a", "shim_face_point_b", "shim_face_normal", "design_thickness", "setup_thickness"},
            "camber_shim",
        )
        return cls(
            decode_point(m["shim_face_point_a"]),
            decode_point(m["shim_face_point_b"]),
            decode_direction(m["shim_face_normal"]),
            float(m["design_thickness"]),
            float(m["setup_thickness"]),
        )

    def mirrored(self) -> "CamberShimConfig":
        """Reflect through the vehicle XZ plane (reference build.py:360-375)."""
        flip = np.array([1.0, -1.0, 1.0])
        return CamberShimConfig(
            Point3(self.shim_face_point_a.data * flip),
            Point3(self.shim_face_point_b.data * flip),
            Direction3(self.shim_face_normal.data * flip),
            self.design_thickness,
            self.setup_thickness,
        )


@dataclass(frozen=True)
class SteeringConfig:
    type: SteeringType

    @classmethod
    def from_mapping(cls, m: dict) -> "SteeringConfig":
        _forbid_extra(m, {"type"}, "steering")
        return cls(SteeringType(m["type"]))


@dataclass(frozen=True)
class VehicleConfig:
    cg_position: Point3
    wheelbase: float
    front_brake_bias: float | None = None
    driven_axle: AxlePosition | None = None

    def __post_init__(self):
        b = self.front_brake_bias
        if b is not None and not 0.0 <= b <= 1.0:
            raise ValueError(f"front_brake_bias must be in [0, 1], got {b}")

    @classmethod
    def from_mapping(cls, m: dict) -> "VehicleConfig":
        _forbid_extra(m, {"cg_position", "wheelbase", "front_brake_bias", "driven_axle"}, "vehicle_config")
        bias = m.get("front_brake_bias")
        driven = m.get("driven_axle")
        return cls(
            decode_point(m["cg_position"]),
            float(m["wheelbase"]),
            None if bias is None else float(bias),
            None if driven is None else AxlePosition(driven),
        )


@dataclass(frozen=True)
class SuspensionConfig:
    """Complete runtime configuration of one built corner."""

    cg_position: Point3
    wheelbase: float
    steering: SteeringConfig
    wheel: WheelConfig
    front_brake_bias: float | None = None
    driven_axle: AxlePosition | None = None
    axle_position: AxlePosition | None = None
    camber_shim: CamberShimConfig | None = None

    @classmethod
    def from_mapping(cls, m: dict) -> "SuspensionConfig":
        _forbid_extra(
            m,
            {"cg_position", "wheelbase", "front_brake_bias", "driven_axle", "steering", "wheel",
             "axle_position", "camber_shim"},
            "config",
        )
        vehicle = VehicleConfig.from_mapping(
            {k: m[k] for k in ("cg_position", "wheelbase", "front_brake_bias", "driven_axle") if k in m}
        )
        shim = m.get("camber_shim")
        pos = m.get("axle_position")
        return cls(
            vehicle.cg_position,
            vehicle.wheelbase,
            SteeringConfig.from_mapping(m["steering"]),
            WheelConfig.from_mapping(m["wheel"]),
            vehicle.front_brake_bias,
            vehicle.driven_axle,
            None if pos is None else AxlePosition(pos),
            None if shim is None else CamberShimConfig.from_mapping(shim),
        )

    @classmethod
    def from_parts(cls, vehicle: VehicleConfig, steering, wheel, axle_position, camber_shim):
        return cls(
            vehicle.cg_position, vehicle.wheelbase, steering, wheel,
            vehicle.front_brake_bias, vehicle.driven_axle, axle_position, camber_shim,
        )
