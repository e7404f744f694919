"""Sweep targets (reference core/targeting.py:35-186)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Union

import numpy as np

from .enums import Axis, TargetPositionMode
from .primitives.constants import EPS_GEOMETRIC
from .primitives.geometry import Direction3


def _axis(values) -> Direction3:
    arr = np.array(values, dtype=np.float64)
    arr.flags.writeable = False
    return Direction3.from_trusted(arr)


class WorldAxisSystem:
    X = _axis((1.0, 0.0, 0.0))
    Y = _axis((0.0, 1.0, 0.0))
    Z = _axis((0.0, 0.0, 1.0))


class PointTarget(NamedTuple):
    point_id: object
    direction: "PointTargetDirection"
    value: float
    mode: TargetPositionMode = TargetPositionMode.RELATIVE


@dataclass(slots=True, frozen=True)
class PointTargetAxis:
    axis: Axis


@dataclass(slots=True, frozen=True)
class PointTargetVector:
    vector: Direction3


PointTargetDirection = Union[PointTargetAxis, PointTargetVector]


def resolve_target(target: PointTargetDirection) -> Direction3:
    if isinstance(target, PointTargetAxis):
        try:
            return (WorldAxisSystem.X, WorldAxisSystem.Y, WorldAxisSystem.Z)[Axis(target.axis)]
        except (ValueError, IndexError):
            raise ValueError(f"Unsupported axis: {target.axis!r}") from None
    if isinstance(target, PointTargetVector):
        return target.vector
    raise TypeError(f"Unsupported target type: {type(target)!r}")


@dataclass
class SweepConfig:
    """One list of ``PointTarget`` per sweep dimension, all of equal length."""

    target_sweeps: list

    def __post_init__(self):
        lengths = [len(s) for s in self.target_sweeps]
        if len(set(lengths)) > 1:
            raise ValueError(f"All sweep dimensions must have the same length. Got: {lengths}")

    @property
    def n_steps(self) -> int:
        return len(self.target_sweeps[0]) if self.target_sweeps else 0


@dataclass(frozen=True)
class ActuatorDOF:
    name: str
    point_keys: tuple
    direction: Direction3

    def matches(self, target: PointTarget) -> bool:
        if target.point_id not in self.point_keys:
            return False
        d = resolve_target(target.direction)
        return abs(float(np.dot(d.data, self.direction.data))) >= 1.0 - EPS_GEOMETRIC


def validate_sweep_controls(sweep_config: SweepConfig, actuator_dofs: tuple) -> None:
    """Every physical actuator must be targeted exactly once per step
    (reference targeting.py:168-186)."""
    for actuator in actuator_dofs:
        for step in range(sweep_config.n_steps):
            hits = sum(1 for sweep in sweep_config.target_sweeps if actuator.matches(sweep[step]))
            if hits != 1:
                raise ValueError(
                    f"Sweep requires exactly one target for actuator '{actuator.name}' "
                    f"along its motion axis; found {hits} at step {step}."
                )
