"""Suspension state container (reference core/state.py:23-175).

``free_points_order = sorted(free_points)`` defines the solver column order:
unknown ``3k+c`` is coordinate ``c`` of the k-th sorted free point.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .primitives.geometry import Point3


@dataclass
class SuspensionState:
    positions: dict
    free_points: set
    free_points_order: list = field(init=False)

    def __post_init__(self) -> None:
        self.free_points_order = sorted(self.free_points)

    @property
    def fixed_points(self) -> set:
        return set(self.positions) - self.free_points

    def get_free_array(self) -> np.ndarray:
        return np.concatenate([self.positions[k].data for k in self.free_points_order])

    def update_from_array(self, array: np.ndarray) -> None:
        n = len(self.free_points_order)
        if array.shape != (3 * n,):
            raise ValueError(f"Array shape {array.shape} doesn't match expected ({3 * n},)")
        rows = array.reshape(n, 3)
        for i, key in enumerate(self.free_points_order):
            self.positions[key] = Point3.from_trusted(rows[i])

    def update_positions(self, new_positions: dict) -> None:
        self.positions = new_positions

    def copy(self) -> "SuspensionState":
        return SuspensionState(
            positions={k: p.copy() for k, p in self.positions.items()},
            free_points=set(self.free_points),
        )

    def get(self, point_id) -> Point3:
        return self.positions[point_id]

    def set(self, point_id, position: Point3) -> None:
        self.positions[point_id] = position.copy()

    def __getitem__(self, point_id) -> Point3:
        return self.positions[point_id]

    def __setitem__(self, point_id, position: Point3) -> None:
        self.positions[point_id] = position.copy()

    def __contains__(self, point_id) -> bool:
        return point_id in self.positions

    def items(self):
        return self.positions.items()

    def keys(self):
        return self.positions.keys()

    def values(self):
        return self.positions.values()
