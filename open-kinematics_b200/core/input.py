"""Public input facade (reference core/input.py:57-76, core/schema/sweep.py:74-196)."""

from __future__ import annotations

import numpy as np

from .enums import Axis, TargetPositionMode
from .primitives.geometry import Direction3
from .primitives.point_ref import Side
from .schema.config import _forbid_extra
from .suspensions.build import build_from_mapping, decode_point_id, decode_side
from .targeting import PointTarget, PointTargetAxis, PointTargetVector, SweepConfig, validate_sweep_controls


def build_suspension(data: dict):
    """Validate a geometry mapping and build the suspension model."""
    return build_from_mapping(data)


_AXES = {"x": Axis.X, "y": Axis.Y, "z": Axis.Z}
_UNIT = np.eye(3)


def _direction(spec: dict):
    _forbid_extra(spec, {"axis", "vector"}, "direction")
    axis, vector = spec.get("axis"), spec.get("vector")
    if (axis is None) == (vector is None):
        raise ValueError("Specify exactly one of 'axis' or 'vector'")
    if axis is not None:
        if isinstance(axis, Axis):
            return PointTargetAxis(axis)
        if axis not in _AXES:
            raise ValueError(f"Invalid axis {axis!r}; expected 'x', 'y' or 'z'")
        return PointTargetAxis(_AXES[axis])
    v = np.asarray(vector, dtype=np.float64)
    if v.shape != (3,):
        raise ValueError(f"Vector must be 3D, got shape {v.shape}")
    norm = float(np.linalg.norm(v))
    if norm == 0.0:
        raise ValueError("Direction vector cannot be zero")
    v = v / norm
    for k in range(3):
        if np.allclose(v, _UNIT[k]):
            return PointTargetAxis(Axis(k))
    return PointTargetVector(Direction3(v))


def _expand(spec: dict, default_steps) -> list:
    name = spec.get("name") or str(spec["point"])
    if spec.get("values") is not None:
        return [float(v) for v in spec["values"]]
    if spec.get("start") is None or spec.get("stop") is None:
        raise ValueError(f"Target '{name}': must specify either 'values' or both 'start' and 'stop'")
    if default_steps is None:
        raise ValueError(f"Target '{name}': no 'steps' count available (specify at target or file level)")
    return list(np.linspace(float(spec["start"]), float(spec["stop"]), int(default_steps)))


def build_sweep(data: dict, suspension=None) -> SweepConfig:
    """Validate a sweep mapping and expand it to per-step targets."""
    _forbid_extra(data, {"version", "steps", "targets"}, "sweep")
    if int(data.get("version", 1)) != 1:
        raise ValueError(f"Unsupported sweep version: {data.get('version')}")
    steps = data.get("steps")
    sequences = [_expand(t, steps) for t in data["targets"]]
    lengths = {len(s) for s in sequences}
    if len(lengths) > 1:
        raise ValueError(f"All targets must have the same length, got: {sorted(lengths)}")

    dimensions = []
    for spec, values in zip(data["targets"], sequences):
        _forbid_extra(spec, {"point", "direction", "name", "side", "mode", "start", "stop", "values"}, "target")
        point = decode_point_id(spec["point"])
        side = None if spec.get("side") is None else decode_side(spec["side"])
        if side is Side.CENTER:
            raise ValueError("Sweep target side must be 'left' or 'right'.")
        mode = TargetPositionMode(spec.get("mode", "relative"))
        direction = _direction(spec["direction"])
        if suspension is not None:
            key = suspension.resolve_target_key(point, side)
            state = suspension.structure()[0]      # point set only; never needs the setup pose
            if key not in state.positions:
                raise ValueError(f"Sweep target point '{key.name}' is not present in suspension type "
                                 f"'{suspension.reported_type_key()}'.")
            if key not in state.free_points and key not in suspension.derived_spec().functions:
                raise ValueError(f"Sweep target point '{key.name}' is fixed in suspension type "
                                 f"'{suspension.reported_type_key()}'.")
        else:
            if side is not None:
                raise ValueError(f"Sweep target for '{point.name}' specifies a 'side', which requires a "
                                 "suspension context to resolve.")
            key = point
        dimensions.append([PointTarget(key, direction, v, mode) for v in values])
    config = SweepConfig(dimensions)
    if suspension is not None:
        validate_sweep_controls(config, suspension.actuator_dofs())
    return config
