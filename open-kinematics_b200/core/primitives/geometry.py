"""Minimal 3-vector value types used at the host boundary.

The reference distinguishes ``Point3`` / ``Vector3`` / ``Direction3``
(core/primitives/geometry.py:85,202,387).  The host shim only needs them as
typed carriers of a ``.data`` ndarray, so one small base class backs all three
and the affine rules are expressed by the result type of each operator.
"""

from __future__ import annotations

import numpy as np

from .constants import EPS_GEOMETRIC


def extract_array(x) -> np.ndarray:
    data = getattr(x, "data", x)
    return np.asarray(data, dtype=np.float64)


class _Triple:
    __slots__ = ("data",)
    __array_priority__ = 100.0

    def __init__(self, data) -> None:
        arr = np.array(extract_array(data), dtype=np.float64, copy=True)
        if arr.shape != (3,):
            raise ValueError(f"{type(self).__name__} requires 3 components, got shape {arr.shape}")
        self.data = arr

    @classmethod
    def from_trusted(cls, data: np.ndarray):
        obj = object.__new__(cls)
        obj.data = data
        return obj

    def copy(self):
        return type(self).from_trusted(self.data.copy())

    def __getitem__(self, idx) -> float:
        return float(self.data[int(idx)])

    def __iter__(self):
        return iter(self.data.tolist())

    def __array__(self, dtype=None, copy=None):
        return np.array(self.data, dtype=dtype, copy=True)

    def __eq__(self, other) -> bool:
        return type(other) is type(self) and bool(np.array_equal(self.data, other.data))

    __hash__ = None

    def almost_equals(self, other, tolerance: float = EPS_GEOMETRIC) -> bool:
        return bool(np.allclose(self.data, extract_array(other), atol=tolerance, rtol=0.0))

    def __repr__(self) -> str:
        x, y, z = self.data
        return f"{type(self).__name__}([{x!r}, {y!r}, {z!r}])"


class Vector3(_Triple):
    """Free vector (displacement)."""

    def __add__(self, other):
        if isinstance(other, Point3):
            return Point3.from_trusted(self.data + other.data)
        return Vector3.from_trusted(self.data + extract_array(other))

    __radd__ = __add__

    def __sub__(self, other):
        return Vector3.from_trusted(self.data - extract_array(other))

    def __rsub__(self, other):
        return Vector3.from_trusted(extract_array(other) - self.data)

    def __mul__(self, scalar):
        return Vector3.from_trusted(self.data * float(scalar))

    __rmul__ = __mul__

    def __truediv__(self, scalar):
        return Vector3.from_trusted(self.data / float(scalar))

    def __neg__(self):
        return Vector3.from_trusted(-self.data)

    def dot(self, other) -> float:
        return float(np.dot(self.data, extract_array(other)))

    def cross(self, other) -> "Vector3":
        return Vector3.from_trusted(np.cross(self.data, extract_array(other)))

    def squared_norm(self) -> float:
        return float(np.dot(self.data, self.data))

    def norm(self) -> float:
        return float(np.sqrt(np.dot(self.data, self.data)))

    def normalize(self) -> "Direction3":
        return Direction3(self.data)


class Direction3(_Triple):
    """Unit direction; normalises on construction and rejects ~zero input
    (reference geometry.py:398-416)."""

    def __init__(self, data) -> None:
        super().__init__(data)
        length = float(np.linalg.norm(self.data))
        if length < EPS_GEOMETRIC:
            raise ValueError("Cannot build a direction from a zero-length vector")
        self.data = self.data / length

    def __mul__(self, scalar):
        return Vector3.from_trusted(self.data * float(scalar))

    __rmul__ = __mul__

    def __neg__(self):
        return Direction3.from_trusted(-self.data)

    def dot(self, other) -> float:
        return float(np.dot(self.data, extract_array(other)))

    def cross(self, other) -> Vector3:
        return Vector3.from_trusted(np.cross(self.data, extract_array(other)))

    def vector(self) -> Vector3:
        return Vector3.from_trusted(self.data.copy())


class Point3(_Triple):
    """Position in the world frame."""

    def __sub__(self, other):
        if isinstance(other, Point3):
            return Vector3.from_trusted(self.data - other.data)
        return Point3.from_trusted(self.data - extract_array(other))

    def __add__(self, other):
        return Point3.from_trusted(self.data + extract_array(other))

    __radd__ = __add__


def midpoint(a: Point3, b: Point3) -> Point3:
    return a + (b - a) / 2.0
