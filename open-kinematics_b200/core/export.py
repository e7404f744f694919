"""Flattening of typed positions at the export boundary (reference core/export.py:11-27)."""

from __future__ import annotations

import numpy as np


def extract_array(position) -> np.ndarray:
    """The raw xyz array of a ``Point3`` / array-like position."""
    return np.asarray(getattr(position, "data", position), dtype=np.float64)


def flatten_positions(positions, output_points) -> dict:
    """``{public point name: (x, y, z)}`` for the selected points, in ``output_points`` order;
    points absent from ``positions`` are skipped."""
    flat = {}
    for point in output_points:
        position = positions.get(point)
        if position is None:
            continue
        raw = extract_array(position)
        flat[point.name.lower()] = (float(raw[0]), float(raw[1]), float(raw[2]))
    return flat
